"""CPU, world_size 2 over gloo: the host side of the N>1 path -- LPT sharding, position-independent CRC
partials, header + padded-row all-gather, merge -- gives exactly the single-rank result.  The per-rank
digests come from the oracle here (no GPU); on the GPU box tests/test_gpu_multi.py does the same with NCCL."""
import os
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from makisu_b200 import shard
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
