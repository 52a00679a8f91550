"""-m gpu: the C++ host side driving the GPU through the C-ABI, against the oracle's reference restatement:
cacheID of a COPY step (add_copy_step.go:102-122) and TarDigest of the committed layer (common.go:67-111)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NOW = 1_600_000_000


def _mk(root, rel, data=b"", mode=0o644, mtime=1_500_000_000):
    p = os.path.join(root, rel)
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "wb") as f:
        f.write(data)
    os.chmod(p, mode)
    os.utime(p, (mtime, mtime))


@pytest.fixture(scope="module")
def eng():
    from makisu_b200.abi import Engine
    e = Engine(device=0, device_arena_bytes=64 << 20, n_host_arenas=2, host_arena_bytes=8 << 20, max_extents=1 << 14)
    yield e
    e.close()


@pytest.fixture(scope="module")
def ctx(tmp_path_factory):
    c = str(tmp_path_factory.mktemp("ctx"))
    rng = np.random.default_rng(42)
    _mk(c, "Dockerfile", b"FROM scratch\nCOPY . /app/\n")
    for d in range(4):
        for i in range(12):
            n = int(rng.integers(0, 600_000))
            _mk(c, f"d{d}/f{i:03d}.bin", rng.integers(0, 256, n, dtype=np.uint8).tobytes())
    _mk(c, "big.bin", rng.integers(0, 256, 11_000_000, dtype=np.uint8).tobytes())  # larger than one 8 MiB arena
    _mk(c, "empty", b"")
    _mk(c, "zeros", bytes(300_000))
    os.symlink("big.bin", os.path.join(c, "link"))
    for d, _, _ in os.walk(c):
        os.utime(d, (1_500_000_000, 1_500_000_000))
    return c


def test_copy_step_cache_id_matches_reference_restatement(eng, ctx):
    from makisu_b200 import host
    from oracle import ctx_crc
    seed = ctx_crc.from_step_cache_id(ctx_crc.plan_seed(True, False), "scratch")
    for args, paths in [(". /app/", ["."]), ("d1 d2 big.bin /x/", ["d1", "d2", "big.bin"]), ("d*/f00?.bin /y/", ["d*/f00?.bin"])]:
        want = ctx_crc.copy_step_cache_id(seed, "COPY", args, ctx, paths)
        got = host.copy_step_cache_id(eng, seed, "COPY", args, ctx, paths)
        assert got == want, (args, got, want)
    # TestCopyStepSetCacheID relations (copy_step_test.go:51-169)
    a = host.copy_step_cache_id(eng, seed, "COPY", ". /app/", ctx, ["."])
    assert a == host.copy_step_cache_id(eng, seed, "COPY", ". /app/", ctx, ["."])
    assert a != host.copy_step_cache_id(eng, seed + "x", "COPY", ". /app/", ctx, ["."])
    assert a != host.copy_step_cache_id(eng, seed, "COPY", ". /app2/", ctx, ["."])


def test_commit_layer_tar_digest_and_chunk_table(ctx, tmp_path):
    from makisu_b200 import host
    from makisu_b200.abi import Engine
    from oracle import layer_tar as lt
    from oracle import lib as olib
    root = tmp_path / "root"
    root.mkdir()
    os.chmod(root, 0o755)
    with Engine(device=0, device_arena_bytes=64 << 20, n_host_arenas=1, host_arena_bytes=64 << 20, max_extents=1 << 14) as eng:
        got = host.commit_copy_ops(eng, str(root), NOW, [host.CopyOperation(["/"], ctx, "/", "/app/", 3, 4)])
    # same layer through 12 MiB arenas: the tar stream continues across submits (SHA-256 midstate on the device)
    with Engine(device=0, device_arena_bytes=12 << 20, n_host_arenas=2, host_arena_bytes=12 << 20, max_extents=1 << 14) as eng:
        got2 = host.commit_copy_ops(eng, str(root), NOW, [host.CopyOperation(["/"], ctx, "/", "/app/", 3, 4)])
    assert got2 == got
    # and the emitted tar: byte-identical to the reference-order stream, digest == TarDigest, readable by tar
    import hashlib
    import tarfile
    tar_path = tmp_path / "layer.tar"
    with Engine(device=0, device_arena_bytes=12 << 20, n_host_arenas=2, host_arena_bytes=12 << 20, max_extents=1 << 14) as eng:
        with open(tar_path, "wb") as f:
            got3 = host.commit_copy_ops(eng, str(root), NOW, [host.CopyOperation(["/"], ctx, "/", "/app/", 3, 4)],
                                        tar_fd=f.fileno())
    assert got3 == got
    data = open(tar_path, "rb").read()
    assert "sha256:" + hashlib.sha256(data).hexdigest() == got["tar_digest"] and len(data) == got["tar_bytes"]
    names = [m.name for m in tarfile.open(tar_path).getmembers()]
    assert names[0] == "app" and "app/big.bin" in names and "app/link" in names
    fs = lt.MemFS(lambda: NOW, str(root))
    entries = fs.add_layer_by_copy_ops([lt.CopyOperation.new(["/"], ctx, "/", "/app/", uid=3, gid=4)])
    blob = b"".join(lt.layer_tar_chunks(entries))
    assert got["n_entries"] == len(entries) and got["tar_bytes"] == len(blob)
    assert got["tar_digest"] == lt.tar_digest(entries)
    # chunk table over the regular files of the layer, in tar order
    arena = np.frombuffer(blob, dtype=np.uint8)
    offs, lens, pos = [], [], 0
    for e in entries:
        pos += len(lt.entry_header_bytes(e))
        if e.hdr.typeflag == lt.TYPE_REG and e.hdr.size:
            offs.append(pos)
            lens.append(e.hdr.size)
            pos += (e.hdr.size + 511) // 512 * 512
    want = olib.chunk_table(arena, offs, lens)
    assert got["n_chunks"] == want["n_chunks"] and got["n_unique"] == want["n_unique"] and got["root"] == want["root"]


def test_memfs_scan_layers_with_whiteouts(eng, tmp_path):
    """AddLayerByScan committed on the GPU: first layer, an empty second layer (Go's 1024-zero-byte tar), then a
    layer of whiteouts -- TarDigest equals the oracle's stream each time (mem_fs_test.go:1038)."""
    import shutil
    from makisu_b200 import host
    from oracle import layer_tar as lt
    root = tmp_path / "r"
    rng = np.random.default_rng(3)
    _mk(str(root), "d/keep.bin", rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes())
    _mk(str(root), "d/gone.bin", rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes())
    _mk(str(root), "d/sub/x.bin", b"x" * 5000)
    for d, _, _ in os.walk(root):
        os.utime(d, (1_500_000_000, 1_500_000_000))
    o = lt.MemFS(lambda: NOW, str(root))
    h = host.MemFS(str(root))
    a = h.commit_scan(eng, NOW)
    assert a["tar_digest"] == lt.tar_digest(o.add_layer_by_scan()) and a["n_entries"] == 5
    b = h.commit_scan(eng, NOW)
    assert b["n_entries"] == 0 and b["tar_bytes"] == 1024
    assert b["tar_digest"] == lt.tar_digest(o.add_layer_by_scan()) == "sha256:5f70bf18a086007016e948b04aed3b82103a36bea41755b6cddfaf10ace3c6ef"
    os.remove(root / "d" / "gone.bin")
    shutil.rmtree(root / "d" / "sub")
    os.utime(root / "d", (NOW, NOW))
    c = h.commit_scan(eng, NOW)
    assert c["tar_digest"] == lt.tar_digest(o.add_layer_by_scan()) and c["n_entries"] == 3 and c["n_chunks"] == 0
    h.close()


def test_add_layers_equal_copy_vs_scan_on_gpu(eng, tmp_path):
    """TestAddLayersEqual (mem_fs_test.go:1118): the same tree committed via copy ops and via scan gives the same
    tar bytes after the first header (the synthesized ancestor differs: clk.Now(), no trailing slash)."""
    import hashlib
    from makisu_b200 import host
    src = tmp_path / "src"
    _mk(str(src), "test1/test2/f.bin", bytes(range(256)) * 2000, mtime=NOW - 50)
    os.symlink("f.bin", src / "test1" / "test2" / "lnk")
    for d in ("test1/test2", "test1", "."):
        os.utime(src / d, (NOW - 40, NOW - 40))
    root1 = tmp_path / "root1"
    root1.mkdir()
    t1, t2 = tmp_path / "a.tar", tmp_path / "b.tar"
    with open(t1, "wb") as f:
        r1 = host.MemFS(str(root1)).commit_copy_ops(eng, NOW, [host.CopyOperation(["/test1"], str(src), "/", "/test1/",
                                                                                  os.getuid(), os.getgid())], tar_fd=f.fileno())
    with open(t2, "wb") as f:
        r2 = host.MemFS(str(src)).commit_scan(eng, NOW, tar_fd=f.fileno())
    b1, b2 = open(t1, "rb").read(), open(t2, "rb").read()
    assert b1[512:] == b2[512:] and b1[:5] == b"test1" and b2[:6] == b"test1/"
    assert r1["root"] == r2["root"] and r1["n_chunks"] == r2["n_chunks"] > 0
    assert r1["tar_digest"] == "sha256:" + hashlib.sha256(b1).hexdigest()
    assert r2["tar_digest"] == "sha256:" + hashlib.sha256(b2).hexdigest()


@pytest.mark.parametrize("arena_mib", [64, 3])
def test_update_from_tar_verifies_blob_and_chunks_files(ctx, tmp_path, arena_mib):
    """UpdateFromTarReader(untar=false) (mem_fs.go:165-255) with the GPU in the loop: the stream read from an fd is
    digested (DiffID of the pulled blob, digest.go:42-50) and its regular files are chunked; the MemFS merge equals
    the oracle's.  3 MiB arenas: the stream spans many submits (SHA-256 midstate parked on the device), members never
    straddle an arena, and tarfile's 10 KiB record padding after the end marker still belongs to the blob."""
    import hashlib
    import io
    import tarfile
    from makisu_b200 import host
    from makisu_b200.abi import Engine
    from oracle import layer_tar as lt
    from oracle import lib as olib
    rng = np.random.default_rng(5)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w", format=tarfile.PAX_FORMAT) as tf:
        def add(name, type_=tarfile.REGTYPE, data=b"", link="", mode=0o644):
            ti = tarfile.TarInfo(name)
            ti.type, ti.mode, ti.mtime, ti.linkname, ti.uid, ti.gid = type_, mode, 1_500_000_000, link, 7, 8
            ti.size = len(data) if type_ == tarfile.REGTYPE else 0
            tf.addfile(ti, io.BytesIO(data) if ti.size else None)
        add("usr/", tarfile.DIRTYPE, mode=0o755)
        add("usr/lib/", tarfile.DIRTYPE, mode=0o755)
        shared = rng.integers(0, 256, 900_000, dtype=np.uint8).tobytes()
        for i in range(14):
            n = int(rng.integers(1, 700_000))
            add(f"usr/lib/lib{i:02d}.so", data=rng.integers(0, 256, n, dtype=np.uint8).tobytes(), mode=0o755)
        add("usr/lib/copy_a.bin", data=shared)
        add("usr/lib/" + "d" * 150 + "/copy_b.bin", data=shared)            # PAX path record + duplicate content
        add("usr/lib/alias.so", tarfile.LNKTYPE, link="usr/lib/lib00.so")
        add("lib", tarfile.SYMTYPE, link="usr/lib", mode=0o777)
        add("etc/empty", data=b"")
        add("etc/.wh.removed", data=b"")
        add("dev/null", tarfile.CHRTYPE)
        add("zeros", data=bytes(400_000))
    data = buf.getvalue()
    assert len(data) % 10240 == 0
    root = tmp_path / "root"
    root.mkdir()
    o = lt.MemFS(lambda: NOW, str(root))
    want_layer = o.update_from_tar(data)
    members = [m for m in lt.read_tar(data) if m.hdr.typeflag == lt.TYPE_REG and m.data_len]
    want = olib.chunk_table(np.frombuffer(data, dtype=np.uint8), [m.data_off for m in members], [m.data_len for m in members])
    tar_path = tmp_path / "base.tar"
    tar_path.write_bytes(data)
    h = host.MemFS(str(root))
    with Engine(device=0, device_arena_bytes=arena_mib << 20, n_host_arenas=2, host_arena_bytes=arena_mib << 20,
                max_extents=1 << 12) as eng:
        with open(tar_path, "rb") as f:
            got = h.update_from_tar(eng, NOW, f.fileno())
        assert got["tar_digest"] == "sha256:" + hashlib.sha256(data).hexdigest()
        assert got["tar_bytes"] == len(data) and got["n_entries"] == len(want_layer)
        assert got["n_chunks"] == want["n_chunks"] and got["n_unique"] == want["n_unique"] and got["root"] == want["root"]
        assert got["n_unique"] < got["n_chunks"]                                   # copy_a / copy_b dedup
        # the tree now holds the base layer: the same blob again merges only what the reference would (the whiteout
        # and its re-added parent), and still verifies
        with open(tar_path, "rb") as f:
            again = h.update_from_tar(eng, NOW, f.fileno())
        want_again = o.update_from_tar(data)
        assert [e.dst for e in want_again] == ["/etc", "/etc/.wh.removed"]
        assert again["n_entries"] == len(want_again) and again["tar_digest"] == got["tar_digest"] and again["root"] == got["root"]
        # through a pipe (gunzip | ingest), digest left to the caller
        r, w = os.pipe()
        import threading
        t = threading.Thread(target=lambda: (os.write(w, data), os.close(w)))
        t.start()
        piped = host.MemFS(str(root)).update_from_tar(eng, NOW, r, flags=host.MKHOST_NO_TAR_DIGEST)
        t.join()
        os.close(r)
        assert piped["root"] == got["root"] and piped["n_entries"] == got["n_entries"]
        assert piped["tar_digest"] == "sha256:" + "00" * 32
        # a member larger than the arena travels in pieces (MKSNAP_X_MORE / MKSNAP_X_CONT): a 4 MiB body of zeros through 3 MiB
        # arenas is one long run of identical max-size chunks; only MKHOST_UNTAR (contiguous body needed) still refuses it
        if arena_mib == 3:
            big = io.BytesIO()
            with tarfile.open(fileobj=big, mode="w") as tf:
                ti = tarfile.TarInfo("huge")
                ti.size = 4 << 20
                tf.addfile(ti, io.BytesIO(bytes(4 << 20)))
            (tmp_path / "big.tar").write_bytes(big.getvalue())
            with open(tmp_path / "big.tar", "rb") as f:
                hg = host.MemFS(str(root)).update_from_tar(eng, NOW, f.fileno())
            assert hg["tar_digest"] == "sha256:" + hashlib.sha256(big.getvalue()).hexdigest()
            assert hg["n_chunks"] == (4 << 20) // 131072 and hg["n_unique"] == 1
            with open(tmp_path / "big.tar", "rb") as f, pytest.raises(host.HostError) as ei:
                host.MemFS(str(root)).update_from_tar(eng, NOW, f.fileno(), flags=host.MKHOST_UNTAR)
            assert "exceeds the arena" in str(ei.value)
    h.close()


def test_content_aware_scan_catches_same_second_same_size_edits(tmp_path):
    """SURVEY section 8f-3.  The reference's scan trusts mtime (1 s resolution) + size, which is why it sync()s and sleeps a
    second first (mem_fs.go:291-311).  With MKHOST_SCAN_CONTENT files the metadata calls unchanged are re-hashed
    on the GPU (one SHA-256 stream per file) and compared with the digest remembered at commit / ingest time.
    Flag clear => the reference's behaviour, checked too."""
    import io
    import tarfile
    from makisu_b200 import host
    from makisu_b200.abi import Engine
    from oracle import layer_tar as lt
    rng = np.random.default_rng(11)
    root = tmp_path / "root"
    T = 1_500_000_000
    _mk(root, "d/a.txt", b"A" * 5000)
    _mk(root, "d/b.bin", rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes())
    _mk(root, "d/big.bin", rng.integers(0, 256, 5_000_000, dtype=np.uint8).tobytes())   # larger than one 2 MiB arena
    _mk(root, "e/c.txt", b"c")
    _mk(root, "e/empty", b"")
    for d, _, _ in os.walk(root):
        os.utime(d, (T, T))
    o = lt.MemFS(lambda: NOW, str(root))
    h = host.MemFS(str(root))
    with Engine(device=0, device_arena_bytes=8 << 20, n_host_arenas=2, host_arena_bytes=8 << 20, max_extents=1 << 12) as eng:
        l1 = o.add_layer_by_scan()
        o.remember_content(l1)
        g1 = h.commit_scan(eng, NOW, flags=host.MKHOST_FILE_DIGESTS)
        assert g1["tar_digest"] == lt.tar_digest(l1) and g1["n_entries"] == len(l1) == 7

        def edit(rel, off):
            p = root / rel
            st = os.lstat(p)
            with open(p, "r+b") as f:
                f.seek(off)
                b = f.read(1)
                f.seek(off)
                f.write(bytes([b[0] ^ 0x5A]))
            os.utime(p, ns=(st.st_atime_ns, st.st_mtime_ns))

        edit("d/b.bin", 123_456)
        edit("d/big.bin", 4_999_999)                                # last byte of a file that spans arenas
        # the reference's scan: metadata identical => nothing to commit (the edit is missed)
        assert o.add_layer_by_scan() == []
        assert h.commit_scan(eng, NOW)["n_entries"] == 0
    # the digest pass streams through 2 MiB arenas: big.bin continues across submits
    with Engine(device=0, device_arena_bytes=2 << 20, n_host_arenas=2, host_arena_bytes=2 << 20, max_extents=1 << 12) as small, \
            Engine(device=0, device_arena_bytes=8 << 20, n_host_arenas=2, host_arena_bytes=8 << 20, max_extents=1 << 12) as eng:
        l3 = o.add_layer_by_scan(content_aware=True)
        o.remember_content(l3)
        assert [e.dst for e in l3] == ["/d", "/d/b.bin", "/d/big.bin"]
        g3 = h.commit_scan(eng, NOW, flags=host.MKHOST_SCAN_CONTENT)
        assert g3["n_entries"] == 3 and g3["tar_digest"] == lt.tar_digest(l3)
        # nothing changed since: both empty, through the small-arena engine this time (exercises stream continuation)
        assert o.add_layer_by_scan(content_aware=True) == []
        assert h.commit_scan(small, NOW, flags=host.MKHOST_SCAN_CONTENT)["n_entries"] == 0
        edit("d/big.bin", 0)
        l5 = o.add_layer_by_scan(content_aware=True)
        assert [e.dst for e in l5] == ["/d", "/d/big.bin"]
        g5 = h.commit_scan(small, NOW, flags=host.MKHOST_SCAN_CONTENT)   # big.bin travels in pieces (MKSNAP_X_MORE / _CONT)
        assert g5["n_entries"] == 2 and g5["tar_digest"] == lt.tar_digest(l5)
    h.close()

    # base layer ingested from a tar (digests remembered per member), files "untarred" with the same metadata, then
    # one of them edited in place
    root2 = tmp_path / "root2"
    body_a, body_b = rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes(), b"config=1\n" * 100
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w", format=tarfile.PAX_FORMAT) as tf:
        for name, typ, data in [("opt/", tarfile.DIRTYPE, b""), ("opt/app.bin", tarfile.REGTYPE, body_a),
                                ("opt/app.conf", tarfile.REGTYPE, body_b)]:
            ti = tarfile.TarInfo(name)
            ti.type, ti.mode, ti.mtime, ti.uid, ti.gid = typ, (0o755 if typ == tarfile.DIRTYPE else 0o644), T, os.getuid(), os.getgid()
            ti.size = len(data)
            tf.addfile(ti, io.BytesIO(data) if data else None)
    data = buf.getvalue()
    _mk(root2, "opt/app.bin", body_a, mtime=T)
    _mk(root2, "opt/app.conf", body_b, mtime=T)
    os.chmod(root2 / "opt", 0o755)
    os.utime(root2 / "opt", (T, T))
    (tmp_path / "base.tar").write_bytes(data)
    o2 = lt.MemFS(lambda: NOW, str(root2))
    h2 = host.MemFS(str(root2))
    with Engine(device=0, device_arena_bytes=8 << 20, n_host_arenas=2, host_arena_bytes=8 << 20, max_extents=1 << 12) as eng:
        assert len(o2.update_from_tar(data, remember=True)) == 3
        with open(tmp_path / "base.tar", "rb") as f:
            got = h2.update_from_tar(eng, NOW, f.fileno(), flags=host.MKHOST_FILE_DIGESTS)
        import hashlib
        assert got["n_entries"] == 3 and got["tar_digest"] == "sha256:" + hashlib.sha256(data).hexdigest()
        # the digests of the regular members are remembered in the tree
        assert h2.file_digest("/opt/app.bin") == hashlib.sha256(body_a).digest() == o2.tree.children["opt"].children["app.bin"].mf.digest
        assert h2.file_digest("/opt/app.conf") == hashlib.sha256(body_b).digest()
        assert h2.file_digest("/opt") is None and h2.file_digest("/nope") is None
        # Scanning a root other than "/" after an ingest: the reference records AbsPath(hdr.Name) as the node's src
        # (mem_fs.go:224), so isOnDisk (mem_fs.go:49-57) looks for /opt/app.bin on the HOST, whites the child out and
        # the walk re-adds it from disk.  Both implementations reproduce that; the re-added files carry fresh digests.
        l = o2.add_layer_by_scan(content_aware=True)
        o2.remember_content(l)
        g = h2.commit_scan(eng, NOW, flags=host.MKHOST_SCAN_CONTENT)
        assert g["n_entries"] == len(l) and g["tar_digest"] == lt.tar_digest(l)
        st = os.lstat(root2 / "opt/app.conf")
        with open(root2 / "opt/app.conf", "r+b") as f:
            f.write(b"config=2")
        os.utime(root2 / "opt/app.conf", ns=(st.st_atime_ns, st.st_mtime_ns))
        assert o2.add_layer_by_scan() == [] and h2.commit_scan(eng, NOW)["n_entries"] == 0          # missed by metadata
        l = o2.add_layer_by_scan(content_aware=True)
        assert [e.dst for e in l] == ["/opt", "/opt/app.conf"]
        g = h2.commit_scan(eng, NOW, flags=host.MKHOST_SCAN_CONTENT)
        assert g["n_entries"] == 2 and g["tar_digest"] == lt.tar_digest(l)
    h2.close()


def test_commit_copy_of_symlinked_sources(eng, tmp_path):
    """COPY linkdir /dst, COPY link.txt /dst2/ (mem_fs.go:380-384: sources go through evalSymlinks before the walk):
    the committed layer holds the link TARGETS' content; TarDigest equals the oracle's stream and SHA-256 of the
    emitted tar, and the archive lists the resolved names."""
    import hashlib
    import tarfile
    from makisu_b200 import host
    from oracle import layer_tar as lt
    c = str(tmp_path / "ctx")
    rng = np.random.default_rng(9)
    _mk(c, "real/a.bin", rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes())
    _mk(c, "real/sub/b.bin", rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes())
    _mk(c, "file.txt", b"F" * 3000)
    os.symlink("real", os.path.join(c, "linkdir"))
    os.symlink("file.txt", os.path.join(c, "link.txt"))
    for d, _, _ in os.walk(c):
        os.utime(d, (1_500_000_000, 1_500_000_000))
    root = tmp_path / "root"
    root.mkdir()
    ops_h = [host.CopyOperation(["/linkdir"], c, "/", "/dst", 1, 2), host.CopyOperation(["/link.txt"], c, "/", "/dst2/", 1, 2)]
    ops_o = [lt.CopyOperation.new(["/linkdir"], c, "/", "/dst", uid=1, gid=2),
             lt.CopyOperation.new(["/link.txt"], c, "/", "/dst2/", uid=1, gid=2)]
    tar_path = tmp_path / "layer.tar"
    with open(tar_path, "wb") as f:
        got = host.commit_copy_ops(eng, str(root), NOW, ops_h, tar_fd=f.fileno())
    entries = lt.MemFS(lambda: NOW, str(root)).add_layer_by_copy_ops(ops_o)
    assert got["tar_digest"] == lt.tar_digest(entries)
    assert got["tar_digest"] == "sha256:" + hashlib.sha256(open(tar_path, "rb").read()).hexdigest()
    names = [m.name for m in tarfile.open(tar_path).getmembers()]
    assert names == ["dst", "dst/a.bin", "dst/sub", "dst/sub/b.bin", "dst2", "dst2/file.txt"]


def test_batch_of_layers_shares_one_session(ctx, tmp_path):
    """mkhost_memfs_commit_layers on the GPU: 6 consecutive COPY layers in ONE session (every arena carries a piece of
    every open layer; 6 SHA-256 chains advance together in K4) give the per-layer TarDigest and tar bytes of 6
    sequential commits (oracle), over 16 MiB arenas (several submits, stream continuation) and in one big arena."""
    import hashlib
    from makisu_b200 import host
    from makisu_b200.abi import Engine
    from oracle import layer_tar as lt
    root = tmp_path / "root"
    root.mkdir()
    os.chmod(root, 0o755)
    specs = [(["/d0"], "/app/d0/"), (["/d1", "/d2"], "/app/"), (["/zeros", "/empty"], "/app/"), (["/d3"], "/srv/"),
             (["/d0"], "/app/d0/"), (["/link"], "/app/")]
    o = lt.MemFS(lambda: NOW, str(root))
    want = []
    for srcs, dst in specs:
        entries = o.add_layer_by_copy_ops([lt.CopyOperation.new(srcs, ctx, "/", dst, uid=5, gid=6)])
        want.append((lt.tar_digest(entries), b"".join(lt.layer_tar_chunks(entries))))
    for arena in (16 << 20, 64 << 20):                             # COPY link resolves to big.bin (11 MB): one entry > its share
        with Engine(device=0, device_arena_bytes=arena, n_host_arenas=2, host_arena_bytes=arena, max_extents=1 << 12) as eng:
            h = host.MemFS(str(root))
            paths = [tmp_path / ("l%d_%d.tar" % (i, arena)) for i in range(len(specs))]
            files = [open(p, "wb") for p in paths]
            got = h.commit_layers(eng, NOW, [[host.CopyOperation(s, ctx, "/", d, 5, 6)] for s, d in specs],
                                  tar_fds=[f.fileno() for f in files])
            for f in files:
                f.close()
            h.close()
        for i, (dig, blob) in enumerate(want):
            assert got[i]["tar_digest"] == dig == "sha256:" + hashlib.sha256(paths[i].read_bytes()).hexdigest(), (i, arena)
            assert paths[i].read_bytes() == blob


def test_commit_layer_with_a_file_larger_than_the_arena(ctx, tmp_path):
    """big.bin (11 MB) through 4 MiB pinned arenas: the packer splits the tar member (MKSNAP_X_MORE / MKSNAP_X_CONT), the
    engine carries the open chunk, the TarDigest stream and the per-file digest across submits -- same TarDigest, tar
    bytes, chunk table and file digest as the oracle over the undivided file (lib/tario/write.go:45 copies any size)."""
    import hashlib
    from makisu_b200 import host
    from makisu_b200.abi import Engine
    from oracle import layer_tar as lt
    from oracle import lib as olib
    root = tmp_path / "root"
    root.mkdir()
    os.chmod(root, 0o755)
    entries = lt.MemFS(lambda: NOW, str(root)).add_layer_by_copy_ops([lt.CopyOperation.new(["/"], ctx, "/", "/app/", uid=3, gid=4)])
    blob = b"".join(lt.layer_tar_chunks(entries))
    arena = np.frombuffer(blob, dtype=np.uint8)
    offs, lens, pos = [], [], 0
    for e in entries:
        pos += len(lt.entry_header_bytes(e))
        if e.hdr.typeflag == lt.TYPE_REG and e.hdr.size:
            offs.append(pos)
            lens.append(e.hdr.size)
            pos += (e.hdr.size + 511) // 512 * 512
    want = olib.chunk_table(arena, offs, lens)
    tar_path = tmp_path / "layer.tar"
    with Engine(device=0, device_arena_bytes=4 << 20, n_host_arenas=2, host_arena_bytes=4 << 20, max_extents=1 << 12) as eng:
        h = host.MemFS(str(root))
        with open(tar_path, "wb") as f:
            got = h.commit_copy_ops(eng, NOW, [host.CopyOperation(["/"], ctx, "/", "/app/", 3, 4)], tar_fd=f.fileno(),
                                    flags=host.MKHOST_FILE_DIGESTS)
        assert tar_path.read_bytes() == blob
        assert got["tar_digest"] == lt.tar_digest(entries) == "sha256:" + hashlib.sha256(blob).hexdigest()
        assert (got["n_chunks"], got["n_unique"], got["root"]) == (want["n_chunks"], want["n_unique"], want["root"])
        assert h.file_digest("/app/big.bin") == hashlib.sha256(open(os.path.join(ctx, "big.bin"), "rb").read()).digest()
        h.close()
