"""CPU: the C++ host side (libmkhost) against the oracle's restatement of the same reference code --
stream order, entry order, header bytes.  No GPU, no hashing."""
import base64
import gzip
import json
import os

import pytest

from makisu_b200 import abi, host
from oracle import ctx_crc, layer_tar as lt
from tests.test_oracle_golden import _parse_block

HERE = os.path.dirname(os.path.abspath(__file__))
NOW = 1_600_000_000


def test_c_abi_exports_every_declared_symbol():
    lib = abi.load()
    import re
    decl = set(re.findall(r"\b(mksnap_[a-z0-9_]+)\s*\(", open(os.path.join(HERE, "..", "include", "mksnap.h")).read()))
    decl -= {"mksnap_t"}
    bound = {n for n, _, _ in abi.SYMBOLS}
    assert decl == bound, (decl ^ bound)
    for n in decl:
        assert hasattr(lib, n)
    hl = host.load()
    hdecl = set(re.findall(r"\b(mkhost_[a-z0-9_]+)\s*\(", open(os.path.join(HERE, "..", "include", "mkhost.h")).read()))
    assert hdecl == {n for n, _, _ in host.SYMBOLS}
    for n in hdecl:
        assert hasattr(hl, n)


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(abi.MksnapError) as ei:
        abi.Engine(device=0, device_arena_bytes=1 << 20, max_extents=16)
    assert "no CPU fallback" in str(ei.value) or "E_CUDA" in str(ei.value)


def test_tar_header_encoder_matches_oracle_and_go_fixture():
    fix = json.load(open(f"{HERE}/golden/reference_fixtures.json"))
    blob = gzip.decompress(base64.b64decode(fix["busybox_headers_gz_b64"]))
    for i in range(fix["busybox_n_headers"]):
        blk = blob[512 * i:512 * (i + 1)]
        h = _parse_block(blk)
        got = host.encode_tar_header(h.name, h.mode, h.uid, h.gid, h.size, h.mtime_ns, h.typeflag, h.linkname)
        assert got == blk, i
    cases = [
        lt.Header(name="/lead/slash.txt", mode=0o644, size=5, mtime_ns=1_500_000_000_999_999_999),
        lt.Header(name="a" * 60 + "/" + "b" * 60, mode=0o755, size=0, typeflag=b"5", mtime_ns=7 * 10**9),
        lt.Header(name="x/" + "y" * 101, mode=0o600, size=1),
        lt.Header(name="d/é.txt", mode=0o644, size=0),
        lt.Header(name="big", mode=0o644, size=9 << 30),
        lt.Header(name="u", mode=0o644, uid=1 << 22, gid=1 << 23),
        lt.Header(name="l", mode=0o777, typeflag=b"2", linkname="t" * 120),
        lt.Header(name="app", mode=0o711, typeflag=b"5", mtime_ns=NOW * 10**9),
    ]
    for h in cases:
        want = lt.encode_header(lt.replace(h, name=h.name.lstrip("/"), mtime_ns=(h.mtime_ns // 10**9) * 10**9))
        got = host.encode_tar_header(h.name, h.mode, h.uid, h.gid, h.size, h.mtime_ns, h.typeflag, h.linkname)
        assert got == want, h


def _mk(root, rel, data=b"", mode=0o644, mtime=1_500_000_000):
    p = os.path.join(root, rel)
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "wb") as f:
        f.write(data)
    os.chmod(p, mode)
    os.utime(p, (mtime, mtime))
    return p


@pytest.fixture
def ctx(tmp_path):
    c = tmp_path / "ctx"
    _mk(c, "Dockerfile", b"FROM scratch\nCOPY . /app/\n")
    _mk(c, "a/b.txt", b"hello b\n" * 100)
    _mk(c, "a/B.txt", b"upper sorts first")
    _mk(c, "a/zz/deep/file.bin", os.urandom(3000))
    _mk(c, "c.txt", b"c" * 1000, mode=0o600)
    _mk(c, "empty", b"")
    _mk(c, ".hidden", b"h")
    os.symlink("c.txt", c / "link")
    os.mkfifo(c / "a" / "fifo")  # special file: skipped
    for d, _, _ in os.walk(c):
        os.utime(d, (1_500_000_000, 1_500_000_000))
    return str(c)


@pytest.mark.parametrize("paths", [["."], ["a", "c.txt"], ["*.txt"], ["a/*", "link"], ["nope"]])
def test_context_stream_order_matches_oracle(ctx, paths):
    want = []
    if paths == ["nope"]:
        with pytest.raises(host.HostError):
            host.describe_context_stream(ctx, paths)
        with pytest.raises(OSError):
            list(ctx_crc.context_segments(ctx, paths))
        return
    for s in ctx_crc.context_segments(ctx, paths):
        want.append(("P " + os.fsdecode(s.data)) if s.kind == "bytes" else f"F {s.size} {s.path}")
    got = host.describe_context_stream(ctx, paths)
    # the oracle does not label link targets separately: normalise
    got = [("P " + g[2:]) if g.startswith("L ") else g for g in got]
    assert got == want


def _desc_from_oracle(entries):
    out = []
    for e in entries:
        h = e.hdr
        out.append("%s %o %d %d %d %d %s %s %s" % (h.typeflag.decode(), h.mode, h.uid, h.gid, h.size, h.mtime_ns // 10**9,
                                                e.dst, h.name, e.src))
    return out


@pytest.mark.parametrize("srcs,dst", [(["/"], "/app/"), (["/c.txt"], "/target/file2"), (["/c.txt"], "/t/d/"),
                                      (["/a"], "/target/dir2"), (["/a", "/c.txt"], "/t/"), (["/a/zz"], "rel/dir/")])
def test_copy_op_layer_matches_oracle(ctx, tmp_path, srcs, dst):
    root = tmp_path / "root"
    root.mkdir()
    os.chmod(root, 0o711)
    fs = lt.MemFS(lambda: NOW, str(root))
    want = _desc_from_oracle(fs.add_layer_by_copy_ops([lt.CopyOperation.new(srcs, ctx, "/work", dst, uid=7, gid=8)]))
    got = host.describe_layer(str(root), NOW, [host.CopyOperation(srcs, ctx, "/work", dst, 7, 8)])
    assert got == want


def test_two_copy_ops_share_ancestors(ctx, tmp_path):
    root = tmp_path / "root"
    root.mkdir()
    fs = lt.MemFS(lambda: NOW, str(root))
    ops_o = [lt.CopyOperation.new(["/a"], ctx, "/", "/srv/x/"), lt.CopyOperation.new(["/c.txt"], ctx, "/", "/srv/y/")]
    ops_h = [host.CopyOperation(["/a"], ctx, "/", "/srv/x/"), host.CopyOperation(["/c.txt"], ctx, "/", "/srv/y/")]
    assert host.describe_layer(str(root), NOW, ops_h) == _desc_from_oracle(fs.add_layer_by_copy_ops(ops_o))


def test_copy_multiple_sources_needs_dir_dst(ctx, tmp_path):
    with pytest.raises(host.HostError) as ei:
        host.describe_layer(str(tmp_path), NOW, [host.CopyOperation(["/a", "/c.txt"], ctx, "/", "/x")])
    assert "destination must end with" in str(ei.value)


def _desc(entries):
    return _desc_from_oracle(entries)


def test_memfs_sequence_scan_whiteout_copy_matches_oracle(tmp_path):
    """mem_fs_test.go: TestCreateLayerByScan (:572), TestAddLayerByScanWhiteout (:1038), metadata-only diff
    (compare_test.go:577) -- the same sequence of layers on the C++ MemFS and on the oracle's."""
    root = tmp_path / "r"
    _mk(root, "d/keep.txt", b"k")
    _mk(root, "d/gone.txt", b"g")
    _mk(root, "d/sub/x.bin", b"x" * 5000)
    _mk(root, "skip/me.txt", b"no")
    _mk(root, ".wh..wh.aufs", b"meta")
    os.symlink("keep.txt", root / "d" / "lnk")
    os.mkfifo(root / "d" / "fifo")
    for d, _, _ in os.walk(root):
        os.utime(d, (1_500_000_000, 1_500_000_000))
    bl = [str(root / "skip")]
    o = lt.MemFS(lambda: NOW, str(root), blacklist=bl)
    h = host.MemFS(str(root), bl)
    first = h.describe_scan(NOW)
    assert first == _desc(o.add_layer_by_scan())
    assert [l.split(" ")[6] for l in first] == ["/d", "/d/gone.txt", "/d/keep.txt", "/d/lnk", "/d/sub", "/d/sub/x.bin"]
    assert h.describe_scan(NOW) == _desc(o.add_layer_by_scan()) == []          # nothing changed
    with open(root / "d" / "keep.txt", "wb") as f:                             # same size, same mtime => similar
        f.write(b"K")
    os.utime(root / "d" / "keep.txt", (1_500_000_000, 1_500_000_000))
    assert h.describe_scan(NOW) == _desc(o.add_layer_by_scan()) == []
    os.remove(root / "d" / "gone.txt")                                         # => whiteout + touched parent
    import shutil
    shutil.rmtree(root / "d" / "sub")
    os.utime(root / "d", (NOW, NOW))
    third = h.describe_scan(NOW)
    assert third == _desc(o.add_layer_by_scan())
    assert [l.split(" ")[7] for l in third] == ["d/", "d/.wh.gone.txt", "d/.wh.sub"]
    # a copy op on top of the scanned tree: ancestors that exist are re-added, new ones synthesized
    ctx = tmp_path / "ctx"
    _mk(ctx, "n/new.txt", b"new")
    for d, _, _ in os.walk(ctx):
        os.utime(d, (1_500_000_000, 1_500_000_000))
    got = h.describe_copy_ops(NOW + 5, [host.CopyOperation(["/n"], str(ctx), "/", "/d/deep/er/", 1, 2)])
    want = _desc(lt.MemFS.add_layer_by_copy_ops(o, [lt.CopyOperation.new(["/n"], str(ctx), "/", "/d/deep/er/", uid=1, gid=2)]))
    # oracle clock is fixed at NOW; only the synthesized ancestors' mtime differs
    assert [l.split(" ")[6:] for l in got] == [l.split(" ")[6:] for l in want]
    assert [l.split(" ")[6] for l in got] == ["/d", "/d/deep", "/d/deep/er", "/d/deep/er/new.txt"]
    # same copy again: nothing new except the re-added ancestors (reference behaviour of addAncestors)
    again = h.describe_copy_ops(NOW + 5, [host.CopyOperation(["/n"], str(ctx), "/", "/d/deep/er/", 1, 2)])
    want2 = _desc(lt.MemFS.add_layer_by_copy_ops(o, [lt.CopyOperation.new(["/n"], str(ctx), "/", "/d/deep/er/", uid=1, gid=2)]))
    assert [l.split(" ")[6] for l in again] == [l.split(" ")[6] for l in want2]
    h.close()


def test_cache_entry_wire_format_matches_reference_and_old_readers():
    """lib/cache/cache_manager.go:34-35,239-252: key prefix, "<tarHex>,<gzipHex>", MAKISU_CACHE_EMPTY, SplitN(…, 2).
    The chunk-table root lives under its own key so an unmodified reader parses the layer entry exactly as before."""
    from oracle import cache_entry as ce
    t, g = "ab" * 32, "cd" * 32
    assert host.cache_key("7d618f69") == ce.cache_key("7d618f69") == "makisu_builder_cache_7d618f69"
    assert host.cache_key("7d618f69", True) == ce.cache_key("7d618f69", True) == "makisu_builder_cache_7d618f69_chunks"
    assert host.cache_entry_create(t, g) == ce.create_entry(t, g) == t + "," + g
    assert host.cache_entry_create(None) == ce.create_entry(None) == "MAKISU_CACHE_EMPTY"
    assert host.cache_entry_parse(t + "," + g) == ce.parse_entry(t + "," + g) == ("sha256:" + t, "sha256:" + g)
    # an old reader keeps everything after the first comma (SplitN 2): this is why the root is NOT a third field
    assert host.cache_entry_parse("a,b,c") == ce.parse_entry("a,b,c") == ("sha256:a", "sha256:b,c")
    for bad in ["", "MAKISU_CACHE_EMPTY", "nocomma"]:
        with pytest.raises(host.HostError) as ei:
            host.cache_entry_parse(bad)
        with pytest.raises(ValueError) as eo:
            ce.parse_entry(bad)
        assert str(ei.value) == str(eo.value) == "parse redis entry: " + bad
    root = bytes(range(32))
    e = host.cache_chunk_entry_create(root, 572245)
    assert e == ce.create_chunk_entry(root, 572245) and host.cache_chunk_entry_parse(e) == ce.parse_chunk_entry(e) == (root, 572245)
    for bad in ["", "zz" * 32 + ",1", "ab" * 32, "ab" * 32 + ",", "ab" * 32 + ",x", "AB" * 32 + ",1"]:
        with pytest.raises(host.HostError):
            host.cache_chunk_entry_parse(bad)
        with pytest.raises(ValueError):
            ce.parse_chunk_entry(bad)
