"""-m gpu.  Arena-fed host paths and the world-of-one NCCL exchange.  (Round 1 wrote these after its GPU budget was
spent and marked them non-strict xfail; all three passed on the driver's B200 run (GPUTEST_r01: 3 xpassed), so the
markers are gone and failures here are failures.)

  * the range-partitioned exchange over a 1-rank NCCL communicator (header / record / level-1 all-gathers of one
    rank; the rank's own slice is a device copy); the same code runs with NCCL at 2, 4 and 8 ranks
    (tests/test_gpu_multi.py, bench.py self-check) and with the in-process transport at 1..8 ranks
    (tests/test_gpu_exchange.py)
  * mkhost_memfs_commit_copy_ops(..., MKHOST_MATERIALIZE): the COPY step's file copy (CopyOperation.Execute,
    lib/snapshot/copy_op.go:82-147) fed from the arena the layer is packed in
  * mkhost_memfs_update_from_tar(..., MKHOST_UNTAR): the base layer untarred from the arena while it is digested."""
import os
import stat

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NOW = 1_600_000_000


def test_nccl_exchange_world_of_one():
    from makisu_b200.abi import Engine, Extent, MKSNAP_X_CDC, MKSNAP_X_CRC
    fb = 1 << 20

    def ext(i):
        e = Extent()
        e.arena_off, e.len, e.crc_suffix, e.flags = i * fb, fb, (199 - i) * fb, MKSNAP_X_CDC | MKSNAP_X_CRC
        return e
    with Engine(device=0, device_arena_bytes=256 << 20, max_extents=4096) as e:
        e.comm_init(Engine.comm_unique_id(), 1, 0)
        e.begin()
        e.synth_fill(0, 0, 200 * fb, 99)
        e.device_submit(0, 200 * fb, [ext(i) for i in range(200)])
        r = e.finish()
        table = e.get_table(r.n_unique).copy()
        x = e.exchange_tables()
        assert (x.n_chunks, x.n_unique, bytes(x.root), e.ctx_crc32(x)) == (r.n_chunks, r.n_unique, bytes(r.root), e.ctx_crc32(r))
        assert e.table_rows() == r.n_unique
        np.testing.assert_array_equal(e.get_table(e.table_rows()), table)


def _tree(root):
    out = {}
    for d, dirs, files in os.walk(root):
        for n in sorted(dirs + files):
            p = os.path.join(d, n)
            st = os.lstat(p)
            rel = os.path.relpath(p, root)
            if stat.S_ISLNK(st.st_mode):
                out[rel] = ("l", os.readlink(p))
            elif stat.S_ISDIR(st.st_mode):
                out[rel] = ("d", stat.S_IMODE(st.st_mode), st.st_uid, st.st_gid)
            else:
                out[rel] = ("f", stat.S_IMODE(st.st_mode), st.st_uid, st.st_gid, open(p, "rb").read())
    return out


@pytest.mark.skipif(os.geteuid() != 0, reason="chown needs root")
def test_commit_copy_ops_materializes_from_the_arena(tmp_path):
    from makisu_b200 import host
    from makisu_b200.abi import Engine
    from oracle import copier as oc
    from oracle import layer_tar as lt
    rng = np.random.default_rng(3)
    ctx = tmp_path / "ctx"
    (ctx / "app" / "sub").mkdir(parents=True)
    for rel, n, mode in [("app/a.bin", 700_000, 0o644), ("app/sub/b.bin", 3_000_000, 0o755), ("app/empty", 0, 0o600),
                         ("conf.txt", 900, 0o640)]:
        p = ctx / rel
        p.write_bytes(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        os.chmod(p, mode)
        os.utime(p, (1_500_000_000, 1_500_000_000))
    os.symlink("a.bin", ctx / "app" / "link")
    for d, _, _ in os.walk(ctx):
        os.utime(d, (1_500_000_000, 1_500_000_000))
    roots = {k: tmp_path / k for k in ("fused", "plain", "oracle")}
    for r in roots.values():
        r.mkdir()
        os.chmod(r, 0o755)
    ops = [host.CopyOperation(["/app"], str(ctx), "/", "/srv/app/", 5, 6), host.CopyOperation(["/conf.txt"], str(ctx), "/", "/etc/conf.txt", 7, 8)]
    with Engine(device=0, device_arena_bytes=16 << 20, n_host_arenas=2, host_arena_bytes=16 << 20, max_extents=1 << 12) as eng:
        fused = host.MemFS(str(roots["fused"])).commit_copy_ops(
            eng, NOW, ops, flags=host.MKHOST_MATERIALIZE | host.MKHOST_MATERIALIZE_CHOWN)
        plain = host.MemFS(str(roots["plain"])).commit_copy_ops(eng, NOW, ops)
    assert fused == plain                                           # the layer, its digest and chunk table do not change
    assert _tree(str(roots["plain"])) == {}                         # without the flag nothing is written
    # what CopyOperation.Execute (--chown) would have put on disk, by the oracle's Copier
    oc.execute_copy_op(str(ctx), ["/app"], str(roots["oracle"]) + "/srv/app/", 5, 6, True, False, False, [])
    oc.execute_copy_op(str(ctx), ["/conf.txt"], str(roots["oracle"]) + "/etc/conf.txt", 7, 8, True, False, False, [])
    got, want = _tree(str(roots["fused"])), _tree(str(roots["oracle"]))
    assert got == want
    assert got["srv/app/sub/b.bin"][1:4] == (0o755, 5, 6) and got["etc/conf.txt"][1:4] == (0o640, 7, 8)
    assert got["srv/app/link"] == ("l", "a.bin") and got["srv/app"] == ("d", 0o755, 5, 6)
    # and the layer equals the oracle's for the same ops
    fs = lt.MemFS(lambda: NOW, str(roots["oracle"]))
    entries = fs.add_layer_by_copy_ops([lt.CopyOperation.new(["/app"], str(ctx), "/", "/srv/app/", uid=5, gid=6),
                                        lt.CopyOperation.new(["/conf.txt"], str(ctx), "/", "/etc/conf.txt", uid=7, gid=8)])
    assert fused["tar_digest"] == lt.tar_digest(entries)


@pytest.mark.skipif(os.geteuid() != 0, reason="chown needs root")
def test_update_from_tar_untars_from_the_arena(tmp_path):
    """UpdateFromTarReader(untar=true) with the GPU in the loop: members are written under the root from the arena
    that is being digested; disk state == the oracle's untar, DiffID == SHA-256 of the blob.  (The untar logic itself
    is verified on the CPU: tests/test_host_tar_ingest_cpu.py.)"""
    import hashlib
    import io
    import tarfile
    from makisu_b200 import host
    from makisu_b200.abi import Engine
    from oracle import layer_tar as lt
    rng = np.random.default_rng(8)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w", format=tarfile.PAX_FORMAT) as tf:
        def add(name, type_=tarfile.REGTYPE, data=b"", link="", mode=0o644):
            ti = tarfile.TarInfo(name)
            ti.type, ti.mode, ti.mtime, ti.linkname, ti.uid, ti.gid = type_, mode, 1_500_000_000, link, 7, 8
            ti.size = len(data) if type_ == tarfile.REGTYPE else 0
            tf.addfile(ti, io.BytesIO(data) if ti.size else None)
        add("usr/", tarfile.DIRTYPE, mode=0o755)
        add("usr/lib/", tarfile.DIRTYPE, mode=0o2755)
        for i in range(6):
            add(f"usr/lib/lib{i}.so", data=rng.integers(0, 256, int(rng.integers(1, 900_000)), dtype=np.uint8).tobytes(), mode=0o755)
        add("usr/lib/alias.so", tarfile.LNKTYPE, link="usr/lib/lib0.so", mode=0o755)
        add("lib", tarfile.SYMTYPE, link="usr/lib", mode=0o777)
        add("etc/", tarfile.DIRTYPE, mode=0o755)
        add("etc/empty", data=b"")
        add("etc/.wh.stale", data=b"")
    data = buf.getvalue()
    (tmp_path / "base.tar").write_bytes(data)
    disks, layers = [], []
    for impl in ("oracle", "gpu"):
        root = tmp_path / impl
        (root / "etc").mkdir(parents=True)
        (root / "etc" / "stale").write_bytes(b"to be whited out")
        os.chmod(root / "etc", 0o700)                              # exists: updated in place to the header's 0755
        for d in (root / "etc", root):
            os.utime(d, (1_400_000_000, 1_400_000_000))
        if impl == "oracle":
            layers.append(len(lt.MemFS(lambda: NOW, str(root)).update_from_tar(data, untar=True)))
        else:
            with Engine(device=0, device_arena_bytes=2 << 20, n_host_arenas=2, host_arena_bytes=2 << 20, max_extents=1 << 10) as eng:
                with open(tmp_path / "base.tar", "rb") as f:
                    got = host.MemFS(str(root)).update_from_tar(eng, NOW, f.fileno(), flags=host.MKHOST_UNTAR)
            assert got["tar_digest"] == "sha256:" + hashlib.sha256(data).hexdigest()
            layers.append(got["n_entries"])
        tree = {}
        for d, dirs, files in os.walk(root):
            for n in sorted(dirs + files):
                p = os.path.join(d, n)
                st = os.lstat(p)
                rel = os.path.relpath(p, root)
                if stat.S_ISLNK(st.st_mode):
                    tree[rel] = ("l", os.readlink(p), st.st_uid, st.st_gid)
                elif stat.S_ISDIR(st.st_mode):
                    tree[rel] = ("d", stat.S_IMODE(st.st_mode), st.st_uid, st.st_gid, st.st_mtime_ns)
                else:
                    tree[rel] = ("f", stat.S_IMODE(st.st_mode), st.st_uid, st.st_gid, st.st_mtime_ns, st.st_nlink, open(p, "rb").read())
        disks.append(tree)
    assert layers[0] == layers[1]
    assert disks[0] == disks[1]
    assert "etc/stale" not in disks[1] and disks[1]["etc"][:4] == ("d", 0o755, 7, 8) and disks[1]["usr/lib/alias.so"][5] == 2


def test_failed_packers_give_their_arena_back(tmp_path):
    """A packer that throws between mksnap_arena_acquire and mksnap_arena_submit (damaged archive) must not cost the
    handle an arena: n_host_arenas + 2 failed ingests on one engine, then a good one (mksnap.cu:641)."""
    import hashlib
    import threading
    from makisu_b200 import host
    from makisu_b200.abi import Engine
    from oracle import layer_tar as lt
    from tests.mock_engine.scenarios import _base_tar
    data = _base_tar(np.random.default_rng(5))
    victim = [m for m in lt.read_tar(data) if m.hdr.typeflag == lt.TYPE_REG and m.data_len][3]
    bad = bytearray(data)
    bad[victim.data_off - 512 + 150] ^= 0x55
    root = tmp_path / "r"
    root.mkdir()
    with Engine(device=0, device_arena_bytes=4 << 20, n_host_arenas=2, host_arena_bytes=1 << 20, max_extents=4096) as eng:
        for _ in range(4):
            r, w = os.pipe()
            t = threading.Thread(target=lambda: (os.write(w, bytes(bad)), os.close(w)))
            t.start()
            with pytest.raises(host.HostError):
                host.MemFS(str(root)).update_from_tar(eng, NOW, r)
            os.close(r)
            t.join()
        good = tmp_path / "good.tar"
        good.write_bytes(data)
        with open(good, "rb") as f:
            ok = host.MemFS(str(root)).update_from_tar(eng, NOW, f.fileno())
        assert ok["tar_digest"] == "sha256:" + hashlib.sha256(data).hexdigest()
        # explicit release through the C-ABI
        ptr, cap, aid = (eng.begin(), eng.arena_acquire())[1]
        eng.arena_release(aid)
        with pytest.raises(Exception):
            eng.arena_release(aid)


class _RealEngineFactory:
    """scenario driver for tests/mock_engine/scenarios.py with real engines (the GPU twin of the CPU mock runs)"""

    def __init__(self):
        self.made = []

    def __call__(self, host_arena_bytes, n_host_arenas=2, max_extents=1 << 12):
        from makisu_b200.abi import Engine
        e = Engine(device=0, device_arena_bytes=host_arena_bytes, n_host_arenas=n_host_arenas,
                   host_arena_bytes=host_arena_bytes, max_extents=max_extents)
        self.made.append(e)
        return e

    def close(self):
        for e in self.made:
            e.close()


def test_incremental_cache_id_on_the_gpu(tmp_path):
    """mkhost_context_crc32_cached on a B200: K0 keeps pure(extent) per CRC extent (k_crc32_extents), unchanged files are
    folded on the host (mksnap_crc_add) -- second build sends no file bytes, edits re-send one file, value == zlib."""
    from makisu_b200 import host
    from oracle import ctx_crc
    from tests.mock_engine import scenarios
    c = scenarios._ctx(str(tmp_path))
    seed = ctx_crc.from_step_cache_id(ctx_crc.plan_seed(True, False), "scratch")
    prefix = (seed + "COPY" + ". /app/").encode()
    f = _RealEngineFactory()
    try:
        eng = f(1 << 20)
        cache = host.CrcCache()
        want = int(ctx_crc.copy_step_cache_id(seed, "COPY", ". /app/", c, ["."]), 16)
        crc, slen, st = cache.context_crc32(eng, prefix, c, ["."])
        assert crc == want and st["files_reused"] == 0 and st["bytes_sent"] == st["bytes_total"] > 0
        h2d0 = eng.stats().h2d_bytes
        crc2, _, st2 = cache.context_crc32(eng, prefix, c, ["."])
        assert crc2 == want and st2["bytes_sent"] == 0 and st2["files_reused"] == st["files_total"]
        assert eng.stats().h2d_bytes - h2d0 < st["bytes_total"] // 100      # < 1 % of the bytes move on the second build
        with open(os.path.join(c, "d2", "f001.bin"), "r+b") as fh:
            fh.write(b"changed!")
        crc3, _, st3 = cache.context_crc32(eng, prefix, c, ["."])
        assert crc3 == int(ctx_crc.copy_step_cache_id(seed, "COPY", ". /app/", c, ["."]), 16) != want
        assert st3["files_reused"] == st["files_total"] - 1
        assert host.context_crc32(eng, prefix, c, ["."])[0] == crc3
    finally:
        f.close()


def test_ingest_member_larger_than_the_arena(tmp_path):
    """mkhost_memfs_update_from_tar on the GPU with members of 3.2 MB / 1.1 MB through 1 MiB pinned arenas: bodies in
    pieces (MKSNAP_X_MORE / MKSNAP_X_CONT + continued streams) give the DiffID, chunk table and per-file digests of the
    undivided archive (same scenario as the CPU mock: tests/mock_engine/scenarios.py)."""
    from tests.mock_engine import scenarios
    f = _RealEngineFactory()
    try:
        orig = f.__call__

        def make(host_arena_bytes, **kw):
            e = orig(host_arena_bytes, **kw)
            e.submits = lambda: 0
            return e
        scenarios.ingest_member_larger_than_the_arena(make, str(tmp_path))
    finally:
        f.close()
