"""CPU, world_size 2 over gloo: the host side of the N>1 path -- LPT sharding, position-independent CRC
partials, header + padded-row all-gather, merge -- gives exactly the single-rank result.  The per-rank
digests come from the oracle here (no GPU); on the GPU box tests/test_gpu_multi.py does the same with NCCL.
Second half: the range-partitioned exchange (mksnap_exchange_tables) -- every transport step as a gloo collective,
world sizes 2 and 3 -- and its single-process model on awkward distributions."""
import os
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from makisu_b200 import abi, shard
from oracle import lib as olib


def _context(seed=5, n=40):
    rng = np.random.default_rng(seed)
    lens = [int(x) for x in rng.integers(0, 90_000, n)] + [700_000, 0, 10]
    blob = rng.integers(0, 256, 250_000, dtype=np.uint8)
    files = [rng.integers(0, 256, l, dtype=np.uint8) for l in lens]
    files[3] = blob.copy()
    files[17] = blob.copy()  # duplicate content on (likely) different ranks => cross-rank dedup
    lens[3] = lens[17] = blob.size
    names = [b"d%02d/f%03d" % (i % 4, i) for i in range(len(files))]
    return names, files, lens


def _rank_result(rank, world, names, files, lens):
    # CRC stream = name_i || content_i in walk order; every segment knows its global suffix
    seg_lens = []
    for nm, l in zip(names, lens):
        seg_lens += [len(nm), l]
    suf = shard.stream_suffixes(seg_lens)
    mine = shard.lpt_shard(lens, world)[rank]
    L = olib.L()
    crc_pure = 0
    for i in mine:
        for j, data in ((2 * i, names[i]), (2 * i + 1, files[i].tobytes())):
            crc_pure ^= L.mko_crc32_mulmod(olib.crc32_pure(data), L.mko_crc32_xpow8n(int(suf[j])))
    arena = np.concatenate([files[i] for i in mine]) if mine else np.zeros(0, np.uint8)
    offs = np.concatenate([[0], np.cumsum([lens[i] for i in mine])[:-1]]) if mine else []
    t = olib.chunk_table(arena, offs, [lens[i] for i in mine])
    hdr = np.zeros(shard.HEADER_WORDS, dtype=np.int64)
    hdr[:6] = [t["n_unique"], crc_pure, sum(len(names[i]) + lens[i] for i in mine), sum(lens[i] for i in mine),
               t["n_chunks"], len(mine)]
    return hdr, t["table"]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names, files, lens = _context()
    hdr, table = _rank_result(rank, world, names, files, lens)
    hdrs = [torch.zeros(shard.HEADER_WORDS, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(hdrs, torch.from_numpy(hdr))
    H = torch.stack(hdrs).numpy().astype(np.uint64)
    pad = max(1, int(H[:, 0].max()))
    mine = np.zeros((pad, 32), dtype=np.uint8)
    mine[: table.shape[0]] = table
    rows = [torch.zeros((pad, 32), dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(rows, torch.from_numpy(mine))
    merged, tot = shard.merge_tables(H, torch.stack(rows).numpy())
    L = olib.L()
    total_len = tot["crc_bytes"]
    crc = tot["crc_pure"] ^ L.mko_crc32_mulmod(0xFFFFFFFF, L.mko_crc32_xpow8n(total_len)) ^ 0xFFFFFFFF
    q.put((rank, crc, tot, olib.merkle_root(merged).hex(), merged.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_equal_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in ps:
        p.join(30)
        assert p.exitcode == 0
    names, files, lens = _context()
    want_crc = zlib.crc32(b"".join(n + f.tobytes() for n, f in zip(names, files)))
    arena = np.concatenate(files)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    single = olib.chunk_table(arena, offs, lens)
    for rank, crc, tot, root, merged in res:
        assert crc == want_crc
        assert tot["n_chunks"] == single["n_chunks"] and tot["n_files"] == len(files)
        assert tot["n_unique"] == single["n_unique"] < single["n_chunks"]  # the duplicated blob dedups across ranks
        assert merged == single["table"].tobytes() and root == single["root"].hex()


def test_lpt_shard_properties():
    rng = np.random.default_rng(1)
    lens = np.clip((rng.zipf(1.3, 5000) * 10).astype(np.int64), 10, 1 << 30)
    for n in (1, 2, 4, 8):
        sh = shard.lpt_shard(lens, n)
        assert sorted(i for s in sh for i in s) == list(range(len(lens)))
        assert shard.imbalance(lens, sh) < 1.0 + max(lens) / (lens.sum() / n) + 1e-9


# ---- range-partitioned exchange (mksnap_exchange_tables): the host logic over gloo -------------------------------
def _sha(b: bytes) -> bytes:
    import hashlib
    return hashlib.sha256(b).digest()


def _exchange_worker(rank, world, port, q):
    """Every transport step of the exchange as a gloo collective; the per-rank arithmetic is shard.range_bounds /
    shard.level0_plan -- the functions the device code (x_phase1..5 in mksnap.cu) restates."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names, files, lens = _context()
    _, table = _rank_result(rank, world, names, files, lens)
    b = shard.range_bounds(table, world)
    # 1 all-to-all of the slices
    outbox = [table[b[j]:b[j + 1]].tobytes() for j in range(world)]
    boxes = [None] * world
    dist.all_gather_object(boxes, outbox)
    got = sorted({boxes[k][rank][i:i + 32] for k in range(world) for i in range(0, len(boxes[k][rank]), 32)})
    mine = np.frombuffer(b"".join(got), dtype=np.uint8).reshape(-1, 32)                 # 2 my range of the global table
    recs = [None] * world                                                               # 3 rows per range + heads
    dist.all_gather_object(recs, (int(mine.shape[0]), mine[:shard.HEAD_ROWS].tobytes()))
    all_u = [r[0] for r in recs]
    p = shard.level0_plan(all_u, rank)                                                  # 4 level 0 of the groups I own
    l1 = [_sha(mine[p["lead"] + 256 * g:p["lead"] + 256 * (g + 1)].tobytes()) for g in range(p["full"])]
    if p["tail_own"]:
        tail, need, k = [mine[p["lead"] + 256 * p["full"]:].tobytes()], p["borrowed"], rank + 1
        while need:
            take = min(need, all_u[k])
            tail.append(recs[k][1][:32 * take])
            need -= take
            k += 1
        l1.append(_sha(b"".join(tail)))
    assert len(l1) == p["groups"]
    alll1 = [None] * world                                                              # 5 level-1 all-gather, upper levels
    dist.all_gather_object(alll1, l1)
    cur = [d for part in alll1 for d in part]
    if sum(all_u) == 0:
        cur = [_sha(b"")]
    while len(cur) > 1:
        cur = [_sha(b"".join(cur[i:i + 256])) for i in range(0, len(cur), 256)]
    q.put((rank, cur[0].hex(), sum(all_u), mine.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_exchange_ranks_equal_single_rank(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 17 * world) % 1000
    ps = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in ps:
        p.join(30)
        assert p.exitcode == 0
    names, files, lens = _context()
    single = olib.chunk_table(np.concatenate(files), np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
    for rank, root, n_unique, rows in res:
        assert root == single["root"].hex() and n_unique == single["n_unique"]
    assert b"".join(r[3] for r in res) == single["table"].tobytes()      # the ranges, in rank order, are the table


def test_exchange_model_edge_cases():
    """shard.exchange_tables_model against the oracle's Merkle root: empty ranks, ranks below one group, one row in
    total, nothing at all, groups straddling several ranks, all digests on one rank (skew), duplicates across ranks."""
    import ctypes
    rng = np.random.default_rng(1)

    def mk(n, skew=False):
        a = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        if skew:
            a[:, 0] = rng.integers(0, 3, n)
        u = sorted({bytes(r) for r in a})
        return np.frombuffer(b"".join(u), dtype=np.uint8).reshape(-1, 32) if u else np.zeros((0, 32), np.uint8)
    cases = [([1000], False), ([700, 900], False), ([0, 5, 300, 0], False), ([10, 0, 3, 700, 256, 255, 1, 0], False),
             ([600] * 4, True), ([0, 0, 0], False), ([1, 0, 0, 0, 0], False), ([4000] * 8, False), ([256, 0], False),
             ([512, 256, 256], True), ([255, 1], False), ([1] * 7, False)]
    for sizes, skew in cases:
        tabs = [mk(n, skew) for n in sizes]
        if len(tabs) > 1 and tabs[0].shape[0] > 3 and tabs[1].shape[0]:
            u = sorted({bytes(r) for r in tabs[1]} | {bytes(r) for r in tabs[0][:3]})
            tabs[1] = np.frombuffer(b"".join(u), dtype=np.uint8).reshape(-1, 32)
        root, ranges = shard.exchange_tables_model(tabs, _sha)
        union = sorted({bytes(r) for t in tabs for r in t})
        cat = np.frombuffer(b"".join(union), dtype=np.uint8).reshape(-1, 32) if union else np.zeros((0, 32), np.uint8)
        out = (ctypes.c_uint8 * 32)()
        olib.L().mko_merkle_root(np.ascontiguousarray(cat).ctypes.data if union else None, len(union), out)
        assert bytes(out) == root, sizes
        assert b"".join(x.tobytes() for x in ranges) == b"".join(union)
        for r, x in enumerate(ranges):                      # every row sits on the rank that owns its prefix
            assert all(shard.range_owner(int.from_bytes(bytes(row[:8]), "big"), len(tabs)) == r for row in x)


def test_level0_plan_invariants():
    """For any distribution of range sizes: every global Merkle group is owned by exactly one rank, the groups appear in
    rank order, a rank never borrows more than 255 rows and never beyond the end of the table -- the facts x_phase4/5 in
    mksnap.cu rely on (hypothesis-free sweep: all size vectors of length <= 4 over a set of boundary values, plus random)."""
    import itertools
    vals = [0, 1, 2, 255, 256, 257, 511, 512, 513, 1000]
    vecs = [list(v) for n in (1, 2, 3) for v in itertools.product(vals, repeat=n)]
    vecs += [list(v) for v in itertools.product([0, 1, 255, 256, 257, 700], repeat=4)]
    rng = np.random.default_rng(0)
    vecs += [list(map(int, rng.integers(0, 3000, int(rng.integers(1, 9))))) for _ in range(2000)]
    for all_u in vecs:
        U = sum(all_u)
        owned = []
        for r in range(len(all_u)):
            p = shard.level0_plan(all_u, r)
            assert abi.exchange_plan(all_u, r) == p                          # the C++ statement inside libmksnap (x_plan)
            assert p["U"] == U and p["g0"] == sum(all_u[:r])
            assert 0 <= p["borrowed"] <= 255 and p["lead"] <= 255
            if p["groups"]:
                first = (p["g0"] + p["lead"]) // 256
                assert (p["g0"] + p["lead"]) % 256 == 0 and p["lead"] < all_u[r]
                owned += list(range(first, first + p["groups"]))
                end = p["g0"] + all_u[r] + p["borrowed"]
                assert end <= U and (end % 256 == 0 or end == U)             # the tail group is complete or is the last one
                assert p["borrowed"] <= sum(all_u[r + 1:])
            else:
                assert p["full"] == p["tail_own"] == p["borrowed"] == 0
        assert owned == list(range((U + 255) // 256)), all_u


def test_bench_global_sharded_layout_reassembles_the_global_stream():
    """bench.py's strong-scaling layout: every rank's extents carry GLOBAL crc_suffix values, so sorting the union of
    all ranks' extents by suffix (descending) must give back name_0, content_0, name_1, ... of the one global context,
    for any rank count; dup contexts point several files at one shared pool region."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(3)
    names = [b"d%02d/f%04d" % (i % 5, i) for i in range(97)]
    sizes = rng.integers(0, 5000, 97).astype(np.uint64)
    total = int(sizes.sum()) + sum(len(n) for n in names)
    for world in (1, 2, 3, 8):
        rows, seen_files = [], 0
        for rank in range(world):
            lay, info = bench.global_sharded_layout(names, sizes, rank, world)
            assert info["bytes_total"] == int(sizes.sum()) and lay["stream_len"] == total
            e = lay["ext"]
            seen_files += len(e) // 2
            for k in range(len(e)):
                if e["flags"][k] == 1:  # a name: its bytes sit in the meta blob
                    o = int(e["arena_off"][k]) - lay["meta_base"]
                    rows.append((int(e["crc_suffix"][k]), "n", lay["blob"][o:o + int(e["len"][k])]))
                else:
                    assert int(e["arena_off"][k]) % 512 == 0 and int(e["arena_off"][k]) + int(e["len"][k]) <= lay["meta_base"]
                    rows.append((int(e["crc_suffix"][k]), "f", int(e["len"][k])))
        assert seen_files == len(names)
        rows.sort(key=lambda r: -r[0])
        after = total
        for i, nm in enumerate(names):
            s_n, kind_n, val_n = rows[2 * i]
            s_f, kind_f, val_f = rows[2 * i + 1]
            after -= len(nm)
            assert (kind_n, val_n, s_n) == ("n", nm, after)
            after -= int(sizes[i])
            assert (kind_f, val_f, s_f) == ("f", int(sizes[i]), after)
        assert after == 0
    # dup: files share pool regions
    pick = rng.integers(0, 10, 97)
    lay, info = bench.global_sharded_layout(names, np.full(97, 1024, np.uint64), 1, 4, pick, 1024, 10)
    assert lay["fill_bytes"] == 10 * 1024 and set(int(x) // 1024 for x in lay["ext"]["arena_off"][1::2]) <= set(range(10))
