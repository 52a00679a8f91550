"""CPU: the oracle's MemFS / tar-stream restatement, exercised the way the reference's own tests do.

Mirrors reference lib/snapshot/mem_fs_test.go: TestCreateLayerByCopy (:688), TestAddLayersEqual (:1118),
TestAddLayerByScanWhiteout (:1038); lib/docker/image/digest_test.go:37-62 (empty tar); SURVEY.md Appendix A.
"""
import hashlib
import io
import os
import tarfile

import pytest

from oracle import ctx_crc, layer_tar as lt

NOW = 1_600_000_000


def _mk(root, rel, data=b"", mode=0o644, mtime=1_500_000_000):
    p = os.path.join(root, rel)
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "wb") as f:
        f.write(data)
    os.chmod(p, mode)
    os.utime(p, (mtime, mtime))
    return p


def _ctx(tmp_path):
    c = tmp_path / "ctx"
    c.mkdir()
    _mk(c, "Dockerfile", b"FROM scratch\nCOPY . /app/\n")
    _mk(c, "a/b.txt", b"hello b\n" * 100)
    _mk(c, "c.txt", b"c" * 1000)
    for d in ("a", "."):
        os.utime(os.path.join(c, d), (1_500_000_000, 1_500_000_000))
    os.chmod(c / "a", 0o755)
    return str(c)


def _tar_bytes(entries):
    return b"".join(lt.layer_tar_chunks(entries))


def test_empty_layer_is_go_empty_tar(tmp_path):
    root = tmp_path / "root"
    root.mkdir()
    fs = lt.MemFS(lambda: NOW, str(root))
    assert lt.tar_digest([]) == "sha256:5f70bf18a086007016e948b04aed3b82103a36bea41755b6cddfaf10ace3c6ef"
    assert fs.add_layer_by_copy_ops([]) == []


def test_appendix_a_worked_example(tmp_path):
    """COPY . /app/ on an empty MemFS: entry order, the slash-less synthesized 'app', forced uid/gid."""
    ctx = _ctx(tmp_path)
    root = tmp_path / "root"
    root.mkdir()
    os.chmod(root, 0o711)
    fs = lt.MemFS(lambda: NOW, str(root))
    op = lt.CopyOperation.new(["/"], ctx, "/", "/app/", uid=7, gid=8)
    entries = fs.add_layer_by_copy_ops([op])
    assert [e.dst for e in entries] == ["/app", "/app/Dockerfile", "/app/a", "/app/a/b.txt", "/app/c.txt"]
    assert [e.hdr.name for e in entries] == ["app", "app/Dockerfile", "app/a/", "app/a/b.txt", "app/c.txt"]
    assert entries[0].hdr.mode == 0o711 and entries[0].hdr.mtime_ns == NOW * 10**9 and entries[0].hdr.typeflag == b"5"
    assert all((e.hdr.uid, e.hdr.gid) == (7, 8) for e in entries)
    blob = _tar_bytes(entries)
    assert len(blob) == 5 * 512 + 512 + 1024 + 1024 + 1024  # hdrs + Dockerfile(27->512) + b.txt(800->1024) + c.txt(1000->1024) + trailer
    assert blob[-1024:] == b"\0" * 1024
    tf = tarfile.open(fileobj=io.BytesIO(blob))
    names = [m.name for m in tf.getmembers()]
    assert names == ["app", "app/Dockerfile", "app/a", "app/a/b.txt", "app/c.txt"]
    assert tf.extractfile("app/a/b.txt").read() == b"hello b\n" * 100
    assert tf.getmember("app/c.txt").mtime == 1_500_000_000 and tf.getmember("app").mtime == NOW
    assert lt.tar_digest(entries) == "sha256:" + hashlib.sha256(blob).hexdigest()
    # CRC walk order differs from tar order (add_copy_step.go:205-237 vs mem_layer.go:232-244)
    segs = list(ctx_crc.context_segments(ctx, ["."]))
    paths = [s.data.decode() for s in segs if s.kind == "bytes"]
    assert paths == [".", "Dockerfile", "a", "a/b.txt", "c.txt"]


@pytest.mark.parametrize("srcs,dst,want", [
    (["/c.txt"], "/target/file2", ["/target", "/target/file2"]),                       # file -> file
    (["/c.txt"], "/target/dir/", ["/target", "/target/dir", "/target/dir/c.txt"]),     # file -> dir/
    (["/a"], "/target/dir2", ["/target", "/target/dir2", "/target/dir2/b.txt"]),       # dir -> dir
    (["/a", "/c.txt"], "/t/", ["/t", "/t/b.txt", "/t/c.txt"]),                          # many -> dir/
])
def test_create_layer_by_copy_shapes(tmp_path, srcs, dst, want):
    ctx = _ctx(tmp_path)
    root = tmp_path / "root"
    root.mkdir()
    fs = lt.MemFS(lambda: NOW, str(root))
    entries = fs.add_layer_by_copy_ops([lt.CopyOperation.new(srcs, ctx, "/", dst)])
    assert [e.dst for e in entries] == want
    tarfile.open(fileobj=io.BytesIO(_tar_bytes(entries))).getmembers()


def test_copy_multiple_sources_needs_dir_dst(tmp_path):
    with pytest.raises(ValueError):
        lt.CopyOperation.new(["/a", "/b"], str(tmp_path), "/", "/x")


def test_add_layers_equal_copy_vs_scan(tmp_path):
    """TestAddLayersEqual: a layer built by copy ops and one found by scanning the same tree on disk must be
    byte-identical tarballs (same order, same header rules)."""
    src = tmp_path / "src"
    _mk(src, "test1/test2/f.txt", b"x" * 3000, mtime=NOW - 50)
    os.symlink("f.txt", src / "test1" / "test2" / "lnk")
    for d in ("test1/test2", "test1", "."):
        os.utime(src / d, (NOW - 40, NOW - 40))
    # route 1: copy ops into an empty MemFS rooted elsewhere
    root1 = tmp_path / "root1"
    root1.mkdir()
    os.utime(root1, (NOW - 40, NOW - 40))
    fs1 = lt.MemFS(lambda: NOW, str(root1))
    e1 = fs1.add_layer_by_copy_ops([lt.CopyOperation.new(["/test1"], str(src), "/", "/test1/",
                                                         uid=os.getuid(), gid=os.getgid())])
    # route 2: scan a root that physically holds the same tree
    fs2 = lt.MemFS(lambda: NOW, str(src))
    e2 = fs2.add_layer_by_scan()
    assert [e.dst for e in e1] == [e.dst for e in e2] == ["/test1", "/test1/test2", "/test1/test2/f.txt", "/test1/test2/lnk"]
    b1, b2 = _tar_bytes(e1), _tar_bytes(e2)
    # the synthesized ancestor of route 1 carries clk.Now() and no trailing slash (Appendix A quirk); the scanned
    # one carries the on-disk mtime and a slash: compare everything after the first header
    assert b1[512:] == b2[512:]
    assert b1[:100].rstrip(b"\0") == b"test1" and b2[:100].rstrip(b"\0") == b"test1/"


def test_scan_whiteout_and_second_scan_is_empty(tmp_path):
    root = tmp_path / "r"
    _mk(root, "d/keep.txt", b"k")
    _mk(root, "d/gone.txt", b"g")
    fs = lt.MemFS(lambda: NOW, str(root))
    first = fs.add_layer_by_scan()
    assert [e.dst for e in first] == ["/d", "/d/gone.txt", "/d/keep.txt"]
    assert fs.add_layer_by_scan() == []  # metadata-only diff: nothing changed
    os.remove(root / "d" / "gone.txt")
    os.utime(root / "d", (NOW, NOW))
    second = fs.add_layer_by_scan()
    wh = [e for e in second if e.whiteout]
    assert len(wh) == 1 and wh[0].hdr.name == "d/.wh.gone.txt" and wh[0].hdr.size == 0
    blob = _tar_bytes(second)
    assert [m.name for m in tarfile.open(fileobj=io.BytesIO(blob)).getmembers()] == ["d", "d/.wh.gone.txt"]


def test_different_content_same_size_is_similar(tmp_path):
    """lib/tario/compare_test.go:577 DifferentContentButSameSizeConsideredSimilar."""
    root = tmp_path / "r"
    p = _mk(root, "f", b"aaaa")
    fs = lt.MemFS(lambda: NOW, str(root))
    assert len(fs.add_layer_by_scan()) == 1
    with open(p, "wb") as f:
        f.write(b"bbbb")
    os.utime(p, (1_500_000_000, 1_500_000_000))
    assert fs.add_layer_by_scan() == []
