"""CPU, world_size 2 over gloo: the tensor-parallel shard plan (blazr_b200/tp.py) reassembles the 1-rank
result -- column-parallel shards concatenate (all_gather), row-parallel shards sum (all_reduce) -- checked with
the oracle's arithmetic on each rank's packed shard.  (The GPU kernels are covered by the -m gpu tests; this
covers the N>1 host logic.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from blazr_b200 import synth, tp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hidden, nh, nkv, hd, ffn, vocab = 512, 8, 2, 64, 1024, 2048
        pl = tp.plan(hidden, nh, nkv, hd, ffn, vocab, rank, world)
        t = synth.GGML["Q6_K"]
        x = synth.random_act(2, hidden, seed=3)
        # ---- column-parallel: gate rows ----
        wg = synth.random_ggml(t, ffn, hidden, seed=5)
        full = oracle.matmul_ggml_f32(t, wg, ffn, hidden, x)
        r0, r1 = pl.ffn_rows
        mine = oracle.matmul_ggml_f32(t, np.ascontiguousarray(wg[r0:r1]), r1 - r0, hidden, x)
        parts = [torch.zeros((2, ffn // world)) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(mine))
        col_ok = np.array_equal(torch.cat(parts, dim=1).numpy(), full)
        # ---- row-parallel: down_proj K split at 256-block granularity, all-reduce(sum) ----
        wd = synth.random_ggml(t, hidden, ffn, seed=6)
        xa = synth.random_act(2, ffn, seed=4)
        fulld = oracle.matmul_ggml_f32(t, wd, hidden, ffn, xa)
        c0, c1 = pl.down_cols
        bpr = ffn // 256 * 210
        blk = np.ascontiguousarray(wd.reshape(hidden, ffn // 256, 210)[:, c0 // 256:c1 // 256].reshape(hidden, -1))
        part = torch.from_numpy(oracle.matmul_ggml_f32(t, blk, hidden, c1 - c0, np.ascontiguousarray(xa[:, c0:c1])))
        dist.all_reduce(part)
        row_err = float(np.abs(part.numpy() - fulld).max() / np.abs(fulld).max())
        # ---- the plan itself tiles every dimension exactly ----
        spans = [torch.tensor(list(pl.q_rows + pl.kv_rows + pl.ffn_rows + pl.vocab_rows))]
        allsp = [torch.zeros(8, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allsp, spans[0])
        ret[rank] = (col_ok, row_err, [a.tolist() for a in allsp])
    finally:
        dist.destroy_process_group()


def test_tp2_shards_reassemble_over_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        col_ok, row_err, spans = ret[r]
        assert col_ok
        assert row_err < 1e-6
        # ranks tile [0, total) contiguously for q rows, kv rows, ffn rows, vocab rows
        for i, total in zip(range(0, 8, 2), (512, 128, 1024, 2048)):
            assert spans[0][i] == 0 and spans[-1][i + 1] == total
            assert all(spans[k][i + 1] == spans[k + 1][i] for k in range(world - 1))


def test_tp_validation_matches_reference_rule():
    # reference src/engine/tensor_parallel.rs:195-206: 32 heads / 8 kv ok at tp 4, 6 kv heads not
    tp.validate_tp_config(4, 32, 8)
    with pytest.raises(ValueError):
        tp.validate_tp_config(4, 32, 6)
    tp.validate_tp_config(1, 7, 3)
    pl = tp.plan(8192, 64, 8, 128, 28672, 128256, 3, 8)  # Llama-3-70B at TP8
    assert pl.n_heads == 8 and pl.n_kv_heads == 1
    assert pl.q_rows == (3072, 4096) and pl.kv_rows == (384, 512)
    assert pl.down_cols == (10752, 14336) and (pl.down_cols[1] - pl.down_cols[0]) // 256 == 14
    assert pl.o_cols[1] - pl.o_cols[0] == 1024


def _ep_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # expert parallelism: 8 experts over 2 ranks, top-3 routing of 2 tokens; every rank
        # evaluates its local selected experts with the oracle on the replicated hidden state, all-reduce combines
        E, top_k, hidden, ffn = 8, 3, 256, 128
        t = synth.GGML["Q8_0"]
        gate = [synth.random_ggml(t, ffn, hidden, seed=40 + e) for e in range(E)]
        down = [synth.random_ggml(t, hidden, ffn, seed=60 + e) for e in range(E)]
        x = synth.random_act(2, hidden, seed=8)
        sel = torch.tensor([[1, 6, 3], [0, 6, 7]], dtype=torch.int32)
        gw = torch.tensor([[0.5, 0.3, 0.2], [0.6, 0.25, 0.15]], dtype=torch.float32)

        def expert(e, xr):
            g = oracle.matmul_ggml_f32(t, gate[e], ffn, hidden, xr)
            return oracle.matmul_ggml_f32(t, down[e], hidden, ffn, np.maximum(g, 0))

        full = np.zeros((2, hidden), dtype=np.float32)
        for tk in range(2):
            for j in range(top_k):
                full[tk] += float(gw[tk, j]) * expert(int(sel[tk, j]), x[tk:tk + 1])[0]
        e0, e1 = tp.expert_range(E, rank, world)
        ls, lg = tp.ep_local_slots(sel, gw, e0, e1)
        part = np.zeros((2, hidden), dtype=np.float32)
        for tk in range(2):
            for j in range(top_k):
                if int(ls[tk, j]) >= 0:
                    part[tk] += float(lg[tk, j]) * expert(e0 + int(ls[tk, j]), x[tk:tk + 1])[0]
                else:
                    assert float(lg[tk, j]) == 0.0
        pt = torch.from_numpy(part)
        dist.all_reduce(pt)
        ret[rank] = ((e0, e1), float(np.abs(pt.numpy() - full).max() / np.abs(full).max()), int((ls >= 0).sum()))
    finally:
        dist.destroy_process_group()


def test_ep2_local_experts_all_reduce_over_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ep_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0][0] == (0, 4) and ret[1][0] == (4, 8)
    assert ret[0][1] < 1e-6 and ret[1][1] < 1e-6
    assert ret[0][2] + ret[1][2] == 6  # every routed slot is hosted by exactly one rank


@pytest.mark.parametrize("vocab,world", [(32000, 2), (32000, 4), (32000, 8), (128256, 8), (128256, 4), (2048, 2), (102400, 8)])
def test_vocab_shards_are_equal_sized_and_index_the_vocabulary(vocab, world):
    """the lm_head all-gather needs the same element count on every rank (32000 / 8 at 128-row granularity is not even:
    unequal shards hang NCCL): every rank contributes vocab_shard_rows columns, padding stays -inf, and the column index
    of the gathered row is the vocabulary id"""
    vs = tp.vocab_shard_rows(vocab, world)
    assert vs % 128 == 0 and vs * world >= vocab
    rng = np.random.Generator(np.random.PCG64(vocab + world))
    logits = rng.standard_normal(vocab).astype(np.float32)
    gathered = np.full((world, vs), -np.inf, dtype=np.float32)
    covered = 0
    for r in range(world):
        pl = tp.plan(4096, 32, 8, 128, 14336, vocab, r, world)
        v0, v1 = pl.vocab_rows
        assert v0 == covered and v1 > v0 and v1 - v0 <= vs
        gathered[r, :v1 - v0] = logits[v0:v1]
        covered = v1
    assert covered == vocab
    assert int(np.argmax(gathered.reshape(-1))) == int(np.argmax(logits))
