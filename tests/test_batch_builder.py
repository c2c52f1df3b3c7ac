"""Host logic of the batched-decode input builder against the rules of reference src/engine/batch_decode.rs:59-131."""
import numpy as np

from blazr_b200 import batch


def test_slot_is_block_times_size_plus_offset():
    # seq_len 19, block_size 16 -> token_pos 18 -> block 1, offset 2 -> slot = blocks[1] * 16 + 2
    s = batch.decode_seq_data(7, list(range(19)), [5, 9], 16)
    assert (s.last_token, s.slot, s.seq_len, s.block_table) == (18, 9 * 16 + 2, 19, [5, 9])
    # block table too short for the new position -> -1 (batch_decode.rs:88-92)
    assert batch.decode_seq_data(7, list(range(33)), [5, 9], 16).slot == -1
    # first token of a block
    assert batch.decode_seq_data(1, list(range(17)), [3, 4], 16).slot == 4 * 16


def test_batch_padding_and_skips():
    hist = {1: [10, 11, 12], 2: list(range(40)), 3: [1]}
    tables = {1: [2], 2: [0, 6, 7], 4: [9]}
    b = batch.build_decode_batch([1, 2, 3, 4], hist, tables, 16, serviceable={1: True, 2: True, 3: True, 4: True})
    assert b.seq_ids == [1, 2]                       # 3 has no block table, 4 no history
    assert b.input_ids.tolist() == [[12], [39]] and b.input_ids.dtype == np.int64
    assert b.slot_mapping.tolist() == [2 * 16 + 2, 7 * 16 + 7]
    assert b.block_table.tolist() == [[2, 0, 0], [0, 6, 7]]   # padded with 0 to the longest table
    assert b.max_seq_len == 40 and b.position == 39
    assert batch.build_decode_batch([3, 4], hist, tables, 16) is None
    assert batch.build_decode_batch([1], hist, tables, 16, serviceable={1: False}) is None
