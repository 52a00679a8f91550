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


# ----------------------------------------------------------------------------------------------------
# Range-partitioned exchange (mksnap_exchange_tables): what every rank does, stated on the host.
#   rank r owns the digests whose big-endian 64-bit prefix p satisfies  floor(p * R / 2^64) == r
#   1. each rank cuts its sorted-unique table at the R+1 range boundaries and sends slice j to rank j (all-to-all)
#   2. each rank sorts + uniques what it received -> its range of the GLOBAL sorted-unique table (u_r rows)
#   3. all-gather of u_r and of every rank's first <=255 rows ("heads")
#   4. Merkle level 0 (groups of 256 consecutive GLOBAL rows): a group is hashed by the rank that owns its first
#      row; the rows it lacks at the end of its range come from the heads of the ranks that follow
#   5. all-gather of the level-1 digests; upper levels on every rank -> the same root as a single table
# `sha256` is injected so this module itself hashes nothing (tests pass hashlib; the product uses the GPU).
# ----------------------------------------------------------------------------------------------------
HEAD_ROWS = 255


def range_owner(prefix_be64: int, n_ranks: int) -> int:
    return (int(prefix_be64) * n_ranks) >> 64


def range_bounds(table: np.ndarray, n_ranks: int) -> List[int]:
    """table [n,32] sorted: -> R+1 row indices; slice j = rows[b[j]:b[j+1]] goes to rank j."""
    pref = [int.from_bytes(bytes(row[:8]), "big") for row in table]
    owners = [range_owner(p, n_ranks) for p in pref]
    out, i = [], 0
    for j in range(n_ranks + 1):
        while i < len(owners) and owners[i] < j:
            i += 1
        out.append(i)
    return out


def level0_plan(all_u: Sequence[int], rank: int) -> dict:
    """Which global Merkle groups rank `rank` hashes, given every rank's range size.  All integers in rows."""
    U = int(sum(all_u))
    g0 = int(sum(all_u[:rank]))
    u = int(all_u[rank])
    first_mult = (g0 + 255) // 256 * 256
    if U == 0 or first_mult >= g0 + u:
        return dict(U=U, g0=g0, lead=0, full=0, tail_own=0, borrowed=0, groups=0)
    lead = first_mult - g0
    owned = u - lead
    full, tail_own = owned // 256, owned % 256
    borrowed = min(256 - tail_own, U - (g0 + u)) if tail_own else 0
    return dict(U=U, g0=g0, lead=lead, full=full, tail_own=tail_own, borrowed=borrowed, groups=full + (1 if tail_own else 0))


def exchange_tables_model(tables: Sequence[np.ndarray], sha256) -> Tuple[bytes, List[np.ndarray]]:
    """tables[r]: rank r's local sorted-unique table [n_r,32].  -> (global root, per-rank range tables)."""
    R = len(tables)
    bounds = [range_bounds(t, R) for t in tables]
    ranges = []
    for me in range(R):                                            # 1 + 2
        got = [bytes(row) for k in range(R) for row in tables[k][bounds[k][me]:bounds[k][me + 1]]]
        uniq = sorted(set(got))
        ranges.append(np.frombuffer(b"".join(uniq), dtype=np.uint8).reshape(-1, 32))
    all_u = [int(x.shape[0]) for x in ranges]                      # 3
    heads = [x[:HEAD_ROWS] for x in ranges]
    U = sum(all_u)
    if U == 0:
        return sha256(b""), ranges
    l1: List[bytes] = []
    for me in range(R):                                            # 4
        p = level0_plan(all_u, me)
        rows = ranges[me]
        for g in range(p["full"]):
            l1.append(sha256(rows[p["lead"] + 256 * g:p["lead"] + 256 * (g + 1)].tobytes()))
        if p["tail_own"]:
            tail = [rows[p["lead"] + 256 * p["full"]:].tobytes()]
            need, k = p["borrowed"], me + 1
            while need:
                take = min(need, all_u[k])
                tail.append(heads[k][:take].tobytes())
                need -= take
                k += 1
            l1.append(sha256(b"".join(tail)))
    cur = l1                                                       # 5
    while len(cur) > 1:
        cur = [sha256(b"".join(cur[i:i + 256])) for i in range(0, len(cur), 256)]
    return cur[0], ranges
