"""Host-side builders of the batched-decode inputs (pure Python, no GPU).

Mirror of reference src/engine/batch_decode.rs:59-131 (`process_decode_batch`): for every decoding sequence the
last token, the KV slot of the new token (`block * block_size + offset`, -1 when the block table is too short) and its
block table; the batch pads block tables with 0 to the longest one and reports the largest sequence length.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np


@dataclass
class DecodeSeqData:
    """reference batch_decode.rs DecodeSeqData"""
    seq_id: int
    last_token: int
    slot: int
    block_table: List[int]
    seq_len: int


@dataclass
class DecodeBatch:
    seq_ids: List[int]
    input_ids: np.ndarray      # int64 [N, 1]
    slot_mapping: np.ndarray   # int32 [N]
    block_table: np.ndarray    # int32 [N, max_num_blocks], shorter tables padded with 0
    max_seq_len: int

    @property
    def position(self) -> int:
        """the `max_seq_len - 1` position argument of forward_with_paged_kv_cache (batch_decode.rs:146)"""
        return self.max_seq_len - 1


def decode_seq_data(seq_id: int, token_history: Sequence[int], blocks: Sequence[int], block_size: int) -> DecodeSeqData:
    """batch_decode.rs:76-98"""
    seq_len = len(token_history)
    last_token = token_history[-1] if seq_len else 0
    token_pos = seq_len - 1
    block_idx, block_offset = divmod(token_pos, block_size) if seq_len else (0, 0)
    slot = int(blocks[block_idx]) * block_size + block_offset if 0 <= block_idx < len(blocks) and seq_len else -1
    return DecodeSeqData(seq_id, int(last_token), int(slot), [int(b) for b in blocks], seq_len)


def build_decode_batch(decode_seqs: Sequence[int], token_histories: Dict[int, Sequence[int]], block_tables: Dict[int, Sequence[int]],
                       block_size: int, serviceable: Optional[Dict[int, bool]] = None) -> Optional[DecodeBatch]:
    """batch_decode.rs:59-131: sequences missing a token history, a block table or a generation config are skipped;
    returns None when nothing is left (the reference returns Ok(()) without a forward pass)."""
    data: List[DecodeSeqData] = []
    for sid in decode_seqs:
        hist = token_histories.get(sid)
        blocks = block_tables.get(sid)
        if hist is None or blocks is None:
            continue
        if serviceable is not None and not serviceable.get(sid, False):
            continue
        data.append(decode_seq_data(sid, hist, blocks, block_size))
    if not data:
        return None
    n = len(data)
    max_blocks = max(len(s.block_table) for s in data)
    bt = np.zeros((n, max_blocks), dtype=np.int32)
    for i, s in enumerate(data):
        bt[i, :len(s.block_table)] = s.block_table
    return DecodeBatch([s.seq_id for s in data], np.asarray([[s.last_token] for s in data], dtype=np.int64),
                       np.asarray([s.slot for s in data], dtype=np.int32), bt, max(s.seq_len for s in data))
