"""-m gpu, needs >= 2 GPUs (skipped otherwise): two processes, one per GPU, each digests its shard, then
mksnap_allgather_tables (NCCL over NVLink) -- every rank must hold the single-GPU result of the whole context -- and
mksnap_exchange_tables (ncclSend/ncclRecv all-to-all): same root and counters, the table range-partitioned."""
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, idfile, q):
    import time

    from makisu_b200 import shard
    from makisu_b200.abi import Engine
    from tests.test_multirank_cpu import _context
    from tests.util import pack
    from makisu_b200.abi import Extent, MKSNAP_X_CDC, MKSNAP_X_CRC
    names, files, lens = _context()
    eng = Engine(device=rank, device_arena_bytes=64 << 20, max_extents=4096)
    if rank == 0:
        uid = Engine.comm_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idfile + ".tmp", idfile)
    else:
        while not os.path.exists(idfile):
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    eng.comm_init(uid, world, rank)
    seg_lens = []
    for nm, l in zip(names, lens):
        seg_lens += [len(nm), l]
    suf = shard.stream_suffixes(seg_lens)
    mine = shard.lpt_shard(lens, world)[rank]
    segs, meta = [], []
    for i in mine:
        segs += [np.frombuffer(names[i], dtype=np.uint8), files[i]]
        meta += [(2 * i, MKSNAP_X_CRC), (2 * i + 1, MKSNAP_X_CRC | MKSNAP_X_CDC)]
    arena, offs = pack(segs)
    ext = []
    for (j, fl), o, sg in zip(meta, offs, segs):
        e = Extent()
        e.arena_off, e.len, e.crc_suffix, e.flags = o, len(sg), int(suf[j]), fl
        ext.append(e)
    eng.begin()
    eng.device_upload(0, 0, arena)
    eng.device_submit(0, arena.size, ext)
    eng.finish()
    res = eng.allgather_tables()
    gathered = (eng.ctx_crc32(res), res.n_chunks, res.n_unique, res.n_files, bytes(res.root), eng.get_table(res.n_unique).tobytes())
    # the same shard again, merged the scalable way: ncclSend/ncclRecv all-to-all, the table stays range-partitioned
    eng.begin()
    eng.device_submit(0, arena.size, ext)
    eng.finish()
    rx = eng.exchange_tables()
    exchanged = (eng.ctx_crc32(rx), rx.n_chunks, rx.n_unique, rx.n_files, bytes(rx.root), eng.get_table(eng.table_rows()).tobytes())
    q.put((rank, gathered, exchanged))
    eng.close()


@pytest.mark.timeout(300)
def test_two_gpus_equal_single(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import lib as olib
    from tests.test_multirank_cpu import _context
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    idfile = str(tmp_path / "nccl_id")
    ps = [ctx.Process(target=_worker, args=(r, 2, idfile, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    names, files, lens = _context()
    want_crc = zlib.crc32(b"".join(n + f.tobytes() for n, f in zip(names, files)))
    single = olib.chunk_table(np.concatenate(files), np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
    for rank, (crc, n_chunks, n_unique, n_files, root, table), _ in res:
        assert crc == want_crc
        assert (n_chunks, n_unique, n_files) == (single["n_chunks"], single["n_unique"], len(files))
        assert root == single["root"] and table == single["table"].tobytes()
    ranges = {}
    for rank, _, (crc, n_chunks, n_unique, n_files, root, rows) in res:
        assert crc == want_crc and root == single["root"]
        assert (n_chunks, n_unique, n_files) == (single["n_chunks"], single["n_unique"], len(files))
        ranges[rank] = rows
    assert b"".join(ranges[r] for r in sorted(ranges)) == single["table"].tobytes()
    assert all(len(v) for v in ranges.values())
