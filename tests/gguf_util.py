"""Writes a tiny random-init Llama-style GGUF file with the independent `gguf` python package (the reference
converter's own writer) from a HostModel, so the tests drive our parser with a file we did not produce ourselves."""
import numpy as np


def write_gguf(path, hm, arch="llama", tie_output=False, extra_tensors=None):
    import gguf
    from gguf import GGMLQuantizationType as QT

    cfg = hm.cfg
    w = gguf.GGUFWriter(path, arch)
    w.add_name(cfg.name)
    w.add_block_count(cfg.n_layers)
    w.add_context_length(2048)
    w.add_embedding_length(cfg.hidden)
    w.add_feed_forward_length(cfg.ffn)
    w.add_head_count(cfg.n_heads)
    w.add_head_count_kv(cfg.n_kv_heads)
    w.add_key_length(cfg.head_dim)
    w.add_layer_norm_rms_eps(cfg.eps)
    w.add_rope_freq_base(cfg.rope_theta)
    w.add_vocab_size(cfg.vocab)
    qt = {"Q4_K": QT.Q4_K, "Q6_K": QT.Q6_K, "Q8_0": QT.Q8_0}

    def lin(name, hl):
        w.add_tensor(name, np.ascontiguousarray(hl.data), raw_shape=hl.data.shape, raw_dtype=qt[hl.fmt])

    w.add_tensor("token_embd.weight", hm.embed.astype(np.float16))
    for i, lay in enumerate(hm.layers):
        p = f"blk.{i}."
        for gname, key in (("attn_q", "q"), ("attn_k", "k"), ("attn_v", "v"), ("attn_output", "o"), ("ffn_gate", "gate"),
                           ("ffn_up", "up"), ("ffn_down", "down")):
            lin(p + gname + ".weight", lay[key])
        w.add_tensor(p + "attn_norm.weight", lay["attn_norm"].astype(np.float32))
        w.add_tensor(p + "ffn_norm.weight", lay["mlp_norm"].astype(np.float32))
    w.add_tensor("output_norm.weight", hm.final_norm.astype(np.float32))
    if not tie_output:
        lin("output.weight", hm.lm_head)
    for name, arr in (extra_tensors or {}).items():
        w.add_tensor(name, arr)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
