"""GGUF parsing (CPU): a file written by the independent `gguf` package is parsed by blazr_b200.gguf_loader --
metadata -> config with the reference's keys/defaults (src/loader/gguf.rs:101-200), tensor directory, GGUF -> HF
names, raw block bytes identical to what was written; malformed files raise."""
import os
import struct

import numpy as np
import pytest

from blazr_b200 import decode, gguf_loader
from gguf_util import write_gguf

gguf = pytest.importorskip("gguf")


@pytest.fixture(scope="module")
def tiny_file(tmp_path_factory):
    hm = decode.build_host_model(decode.PRESETS["tiny"], "Q4_K_M", seed=3)
    path = str(tmp_path_factory.mktemp("gguf") / "tiny.gguf")
    write_gguf(path, hm)
    return path, hm


def test_metadata_to_config(tiny_file):
    path, hm = tiny_file
    g = gguf_loader.Gguf.open(path)
    cfg = gguf_loader.config_from_gguf_metadata(g)
    ref = hm.cfg
    assert (cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.ffn, cfg.vocab) == \
        (ref.hidden, ref.n_layers, ref.n_heads, ref.n_kv_heads, ref.head_dim, ref.ffn, ref.vocab)
    assert abs(cfg.eps - ref.eps) < 1e-12 and cfg.rope_theta == ref.rope_theta
    info = gguf_loader.get_gguf_info(path)
    assert info.architecture == "llama" and info.num_layers == ref.n_layers and info.vocab_size == ref.vocab
    assert info.quantization_type in ("Q4_K", "F32") and not info.is_moe
    assert info.file_size_bytes == os.path.getsize(path)


def test_tensor_bytes_and_names(tiny_file):
    path, hm = tiny_file
    g = gguf_loader.Gguf.open(path)
    h2 = gguf_loader.host_model_from_gguf(g)
    assert len(h2.layers) == len(hm.layers)
    for a, b in zip(h2.layers, hm.layers):
        for k in ("q", "k", "v", "o", "gate", "up", "down"):
            assert a[k].fmt == b[k].fmt and (a[k].N, a[k].K) == (b[k].N, b[k].K)
            assert np.array_equal(a[k].data, b[k].data)
        assert np.array_equal(a["attn_norm"], b["attn_norm"])
    assert np.array_equal(h2.lm_head.data, hm.lm_head.data)
    assert np.array_equal(h2.embed, hm.embed)
    assert gguf_loader.hf_name("blk.3.attn_q.weight") == "model.layers.3.self_attn.q_proj.weight"
    assert gguf_loader.hf_name("blk.0.ffn_down.weight") == "model.layers.0.mlp.down_proj.weight"
    assert gguf_loader.hf_name("token_embd.weight") == "model.embed_tokens.weight"
    assert gguf_loader.hf_name("output.weight") == "lm_head.weight"
    ti = g.tensor_info("blk.0.ffn_down.weight")
    assert ti.shape == (hm.cfg.ffn, hm.cfg.hidden) and ti.offset % g.alignment == 0


def test_defaults_match_reference(tmp_path):
    """reference gguf.rs: heads default 32, kv heads default = heads, eps 1e-5, rope base 10000"""
    w = gguf.GGUFWriter(str(tmp_path / "d.gguf"), "llama")
    w.add_block_count(2)
    w.add_embedding_length(4096)
    w.add_vocab_size(100)
    w.add_tensor("dummy", np.zeros(4, dtype=np.float32))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    cfg = gguf_loader.config_from_gguf_metadata(gguf_loader.Gguf.open(str(tmp_path / "d.gguf")))
    assert (cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.eps, cfg.rope_theta) == (32, 32, 128, 1e-5, 10000.0)


def test_malformed_files_raise(tmp_path, tiny_file):
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(b"NOPE" + b"\0" * 64)
    with pytest.raises(gguf_loader.GgufError):
        gguf_loader.Gguf.open(str(bad))
    path, _ = tiny_file
    raw = open(path, "rb").read()
    trunc = tmp_path / "trunc.gguf"
    trunc.write_bytes(raw[:len(raw) // 2])
    with pytest.raises(gguf_loader.GgufError):
        gguf_loader.Gguf.open(str(trunc))
    v9 = tmp_path / "v9.gguf"
    v9.write_bytes(raw[:4] + struct.pack("<I", 9) + raw[8:])
    with pytest.raises(gguf_loader.GgufError):
        gguf_loader.Gguf.open(str(v9))
    g = gguf_loader.Gguf.open(path)
    with pytest.raises(gguf_loader.GgufError):
        g.tensor_info("no.such.tensor")


def test_stacked_expert_tensors(tmp_path):
    """3-D ffn_*_exps tensors (ne = K, N, E) -> per-expert row views; gate and up fused row-wise per expert"""
    from blazr_b200 import synth
    E, hidden, ffn = 4, 256, 128
    tq4, tq8 = synth.GGML["Q4_K"], synth.GGML["Q8_0"]
    gate = [synth.random_ggml(tq4, ffn, hidden, seed=10 + e) for e in range(E)]
    up = [synth.random_ggml(tq4, ffn, hidden, seed=20 + e) for e in range(E)]
    down = [synth.random_ggml(tq8, hidden, ffn, seed=30 + e) for e in range(E)]
    path = str(tmp_path / "moe.gguf")
    w = gguf.GGUFWriter(path, "llama")
    w.add_block_count(1)
    w.add_expert_count(E)
    QT = gguf.GGMLQuantizationType
    for name, mats, qt in (("ffn_gate_exps", gate, QT.Q4_K), ("ffn_up_exps", up, QT.Q4_K), ("ffn_down_exps", down, QT.Q8_0)):
        stacked = np.ascontiguousarray(np.stack(mats, axis=0))          # [E, N, row_bytes]
        w.add_tensor(f"blk.0.{name}.weight", stacked, raw_shape=stacked.shape, raw_dtype=qt)
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    g = gguf_loader.Gguf.open(path)
    assert g.tensor_info("blk.0.ffn_gate_exps.weight").shape == (hidden, ffn, E)
    assert g.tensor_info("blk.0.ffn_down_exps.weight").shape == (ffn, hidden, E)
    ex = gguf_loader.host_experts_from_gguf(g, 0)
    assert len(ex) == E
    for e, (gu_b, gt, dn_b, dt) in enumerate(ex):
        assert (gt, dt) == (tq4, tq8)
        assert np.array_equal(gu_b[:ffn], gate[e]) and np.array_equal(gu_b[ffn:], up[e]) and np.array_equal(dn_b, down[e])
    assert gguf_loader.get_gguf_info(path).is_moe
    assert gguf_loader.hf_name("blk.0.ffn_gate_exps.weight") == "model.layers.0.mlp.experts.gate_proj.weight"


def test_unsupported_architecture_and_stray_tensors_raise(tmp_path):
    """No silent fallback: a qwen2-tagged file (NEOX RoPE, q/k/v biases) is refused, a llama file that carries tensors the dense
    decoder would drop (a bias here) is refused, and a missing embedding_length is an error as in reference gguf.rs:120-123."""
    hm = decode.build_host_model(decode.PRESETS["tiny"], "Q8_0", seed=5)
    p1 = str(tmp_path / "qwen.gguf")
    write_gguf(p1, hm, arch="qwen2")
    with pytest.raises(gguf_loader.GgufError, match="architecture 'qwen2'"):
        gguf_loader.host_model_from_gguf(gguf_loader.Gguf.open(p1))
    p2 = str(tmp_path / "bias.gguf")
    write_gguf(p2, hm, extra_tensors={"blk.0.attn_q.bias": np.zeros(hm.cfg.n_heads * hm.cfg.head_dim, dtype=np.float32)})
    with pytest.raises(gguf_loader.GgufError, match="does not consume"):
        gguf_loader.host_model_from_gguf(gguf_loader.Gguf.open(p2))
    p3 = str(tmp_path / "mistral.gguf")
    write_gguf(p3, hm, arch="mistral")
    assert len(gguf_loader.host_model_from_gguf(gguf_loader.Gguf.open(p3)).layers) == hm.cfg.n_layers
    w = gguf.GGUFWriter(str(tmp_path / "noemb.gguf"), "llama")
    w.add_block_count(2)
    w.add_vocab_size(100)
    w.add_tensor("dummy", np.zeros(4, dtype=np.float32))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    with pytest.raises(gguf_loader.GgufError, match="embedding_length"):
        gguf_loader.config_from_gguf_metadata(gguf_loader.Gguf.open(str(tmp_path / "noemb.gguf")))
