"""Host logic of the multi-GPU path (one process per GPU, files sharded, one exchange step).

  lpt_shard         greedy longest-processing-time partition of the walk-ordered file list by bytes; whole
                    files only (chunks never span files).
  stream_suffixes   the global crc_suffix of every context-stream segment, so each rank's CRC partial is
                    position independent and the cacheID is the XOR of the partials.
  merge_tables      what mksnap_allgather_tables does on the device, stated on the host for the gloo tests:
                    header all-gather (rows, CRC partial, counters), padded-row all-gather, sort+unique.
No hashing here: digests come from libmksnap (or, in the CPU tests, from the oracle).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def lpt_shard(lens: Sequence[int], n_ranks: int) -> List[List[int]]:
    """-> per rank, the (ascending) indices of the files it owns."""
    order = sorted(range(len(lens)), key=lambda i: (-int(lens[i]), i))
    load = [0] * n_ranks
    out: List[List[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lens[i])
    return [sorted(x) for x in out]


def imbalance(lens: Sequence[int], shards: List[List[int]]) -> float:
    loads = [sum(int(lens[i]) for i in s) for s in shards]
    return max(loads) / (sum(loads) / len(loads)) if sum(loads) else 1.0


def stream_suffixes(seg_lens: Sequence[int]) -> np.ndarray:
    a = np.asarray(seg_lens, dtype=np.uint64)
    return (np.uint64(a.sum()) - np.cumsum(a)).astype(np.uint64)


HEADER_WORDS = 8  # rows, crc partial, crc_bytes, cdc_bytes, n_chunks, n_files, n_streams, reserved


def merge_tables(headers: np.ndarray, padded_rows: np.ndarray) -> Tuple[np.ndarray, dict]:
    """headers: [R, 8] u64; padded_rows: [R, pad, 32] u8 (rank r valid for headers[r,0] rows).
    -> (sorted unique table [m,32], summed counters with the XOR-combined CRC partial)."""
    R = headers.shape[0]
    rows = [padded_rows[r, : int(headers[r, 0])] for r in range(R)]
    cat = np.concatenate(rows, axis=0) if rows else np.zeros((0, 32), np.uint8)
    if cat.shape[0]:
        view = np.ascontiguousarray(cat).view([("d", "V32")]).reshape(-1)
        # bytewise order == order of the raw 32-byte strings
        keys = [bytes(x) for x in cat]
        uniq = sorted(set(keys))
        table = np.frombuffer(b"".join(uniq), dtype=np.uint8).reshape(-1, 32)
        del view
    else:
        table = np.zeros((0, 32), np.uint8)
    crc = 0
    for r in range(R):
        crc ^= int(headers[r, 1])
    tot = dict(crc_pure=crc, crc_bytes=int(headers[:, 2].sum()), cdc_bytes=int(headers[:, 3].sum()),
               n_chunks=int(headers[:, 4].sum()), n_files=int(headers[:, 5].sum()), n_unique=int(table.shape[0]))
    return table, tot
