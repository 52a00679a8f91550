#!/usr/bin/env python
"""bench_fs.py -- the same hot path driven from REAL FILES (not the contract bench; extra evidence).

Builds a synthetic build context on tmpfs, then times
  GPU : libmkhost (C++ filepath.Walk-ordered packer, parallel pread into pinned arenas) + libmksnap
        * cacheID of `COPY . /app/`                       (reference add_copy_step.go:102-122)
        * commit of the layer: TarDigest + chunk table     (reference common.go:67-111)
  CPU : the oracle's file-based restatement of the same two steps, single thread like the reference
        (zlib CRC-32 over 32 KiB reads; hashlib SHA-256 (OpenSSL, SHA-NI) over the tar stream).
and checks that both agree bit for bit.  python bench_fs.py [--files 4096 --file-kib 512 --dir /dev/shm/mkctx]
"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=4096)
    ap.add_argument("--file-kib", type=int, default=512)
    ap.add_argument("--dir", default="/dev/shm/mkctx")
    ap.add_argument("--threads", type=int, default=32)
    a = ap.parse_args()
    from makisu_b200 import host
    from makisu_b200.abi import Engine
    from oracle import ctx_crc, layer_tar as lt

    shutil.rmtree(a.dir, ignore_errors=True)
    ctx = os.path.join(a.dir, "ctx")
    root = os.path.join(a.dir, "root")
    os.makedirs(root)
    rng = np.random.default_rng(1)
    fb = a.file_kib << 10
    for i in range(a.files):
        d = os.path.join(ctx, "d%03d" % (i % 64))
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "f%06d.bin" % i), "wb") as f:
            f.write(rng.integers(0, 256, fb, dtype=np.uint8).tobytes())
    total = a.files * fb
    seed = ctx_crc.from_step_cache_id(ctx_crc.plan_seed(True, False), "scratch")
    arena = 256 << 20  # several arenas in flight: file reads of batch k+1 overlap H2D + kernels of batch k
    out = {"files": a.files, "file_kib": a.file_kib, "bytes": total, "tmpfs": a.dir, "host_threads": a.threads}
    with Engine(device=0, device_arena_bytes=arena, n_host_arenas=4, host_arena_bytes=arena, max_extents=1 << 16,
                max_chunks=total // 4096 + a.files + 1024) as eng:
        host.copy_step_cache_id(eng, seed, "COPY", ". /app/", ctx, ["."])  # warm up (page cache, CUDA)
        t = time.perf_counter()
        crc, slen = host.context_crc32(eng, (seed + "COPY" + ". /app/").encode(), ctx, ["."], a.threads)
        out["gpu_cacheid_s"] = time.perf_counter() - t
        gid = "%x" % crc
        t = time.perf_counter()
        layer = host.commit_copy_ops(eng, root, 1_600_000_000, [host.CopyOperation(["/"], ctx, "/", "/app/")], a.threads)
        out["gpu_commit_s"] = time.perf_counter() - t
        t = time.perf_counter()
        layer2 = host.commit_copy_ops(eng, root, 1_600_000_000, [host.CopyOperation(["/"], ctx, "/", "/app/")], a.threads,
                                      flags=host.MKHOST_NO_TAR_DIGEST)
        out["gpu_commit_no_tardigest_s"] = time.perf_counter() - t
        assert layer2["root"] == layer["root"] and layer2["n_chunks"] == layer["n_chunks"]
    t = time.perf_counter()
    cid = ctx_crc.copy_step_cache_id(seed, "COPY", ". /app/", ctx, ["."])
    out["cpu_cacheid_s"] = time.perf_counter() - t
    t = time.perf_counter()
    fs = lt.MemFS(lambda: 1_600_000_000, root)
    td = lt.tar_digest(fs.add_layer_by_copy_ops([lt.CopyOperation.new(["/"], ctx, "/", "/app/")]))
    out["cpu_commit_s"] = time.perf_counter() - t
    assert gid == cid, (gid, cid)
    assert layer["tar_digest"] == td, (layer["tar_digest"], td)
    out.update(cache_id=gid, tar_digest=td, n_chunks=int(layer["n_chunks"]),
               gpu_cacheid_GiBps=total / 2**30 / out["gpu_cacheid_s"], cpu_cacheid_GiBps=total / 2**30 / out["cpu_cacheid_s"],
               gpu_commit_GiBps=total / 2**30 / out["gpu_commit_s"],
               gpu_commit_no_tardigest_GiBps=total / 2**30 / out["gpu_commit_no_tardigest_s"], cpu_commit_GiBps=total / 2**30 / out["cpu_commit_s"],
               note="GPU commit includes the serial TarDigest of ONE stream (latency-bound, ~36 MB/s) plus CDC + chunk SHA-256; "
                    "CPU commit is tar + SHA-256 only (no gzip).  Digests agree bit for bit.")
    print(json.dumps(out))
    shutil.rmtree(a.dir, ignore_errors=True)


if __name__ == "__main__":
    main()
