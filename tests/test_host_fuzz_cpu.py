"""CPU: randomised agreement between the C++ host side and the oracle (both restate the same reference code
independently: one in C++, one in Python) -- tar header bytes over fuzzed fields, and stream / layer order over
random directory trees."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from makisu_b200 import host
from oracle import ctx_crc, layer_tar as lt

name_chars = st.sampled_from(list("abcXYZ019._-+ ~") + ["é", "ß", "日"])
segment = st.text(name_chars, min_size=1, max_size=60).filter(lambda s: s not in (".", ".."))
path = st.lists(segment, min_size=1, max_size=6).map("/".join)


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(name=path, link=st.one_of(st.just(""), path), mode=st.integers(0, 0o7777), uid=st.integers(0, 1 << 34),
       gid=st.integers(0, 1 << 22), size=st.integers(0, 1 << 36), mtime=st.integers(0, (1 << 33) - 1),
       frac=st.integers(0, 999_999_999), kind=st.sampled_from([b"0", b"5", b"2", b"1"]), lead=st.booleans())
def test_tar_header_fuzz(name, link, mode, uid, gid, size, mtime, frac, kind, lead):
    if kind == b"5":
        name += "/"
    if kind != b"2" and kind != b"1":
        link = ""
    if kind != b"0":
        size = 0
    h = lt.Header(name=name, mode=mode, uid=uid, gid=gid, size=size, mtime_ns=mtime * 10**9 + frac, typeflag=kind,
                  linkname=link)
    want = lt.encode_header(lt.replace(h, mtime_ns=mtime * 10**9))  # write.go:61 truncation, then the writer
    got = host.encode_tar_header(("/" if lead else "") + name, mode, uid, gid, size, h.mtime_ns, kind, link)
    assert got == want


def _random_tree(root, rng, depth=0):
    n = int(rng.integers(1, 7))
    for _ in range(n):
        nm = "".join(rng.choice(list("abAB01._-"), size=int(rng.integers(1, 9))))
        if nm in (".", "..") or os.path.lexists(os.path.join(root, nm)):
            continue
        p = os.path.join(root, nm)
        r = rng.random()
        if r < 0.3 and depth < 3:
            os.mkdir(p)
            _random_tree(p, rng, depth + 1)
        elif r < 0.4:
            os.symlink("some/target" if rng.random() < 0.5 else "../up/target", p)
        else:
            with open(p, "wb") as f:
                f.write(os.urandom(int(rng.integers(0, 3000))))
            os.chmod(p, int(rng.choice([0o644, 0o600, 0o755, 0o4755, 0o1777])))
        os.utime(p, (1_400_000_000 + int(rng.integers(0, 10**8)),) * 2, follow_symlinks=False)


@pytest.mark.parametrize("seed", range(12))
def test_random_trees_stream_and_layer_order(tmp_path, seed):
    rng = np.random.default_rng(seed)
    ctx = tmp_path / "ctx"
    ctx.mkdir()
    _random_tree(str(ctx), rng)
    for d, _, _ in os.walk(ctx):
        os.utime(d, (1_450_000_000, 1_450_000_000))
    want = []
    for s in ctx_crc.context_segments(str(ctx), ["."]):
        want.append(("P " + os.fsdecode(s.data)) if s.kind == "bytes" else f"F {s.size} {s.path}")
    got = [("P " + g[2:]) if g.startswith("L ") else g for g in host.describe_context_stream(str(ctx), ["."])]
    assert got == want
    root = tmp_path / "root"
    root.mkdir()
    fs = lt.MemFS(lambda: 1_600_000_000, str(root))
    entries = fs.add_layer_by_copy_ops([lt.CopyOperation.new(["/"], str(ctx), "/", "/opt/app/", uid=5, gid=6)])
    o = ["%s %o %d %d %d %d %s %s %s" % (e.hdr.typeflag.decode(), e.hdr.mode, e.hdr.uid, e.hdr.gid, e.hdr.size,
                                         e.hdr.mtime_ns // 10**9, e.dst, e.hdr.name, e.src) for e in entries]
    assert host.describe_layer(str(root), 1_600_000_000, [host.CopyOperation(["/"], str(ctx), "/", "/opt/app/", 5, 6)]) == o
    # and the header bytes of every entry
    for e in entries:
        hb = lt.entry_header_bytes(e)
        assert host.encode_tar_header(e.hdr.name, e.hdr.mode, e.hdr.uid, e.hdr.gid, e.hdr.size, e.hdr.mtime_ns,
                                      e.hdr.typeflag, e.hdr.linkname) == hb


def test_absolute_symlink_outside_root_is_an_error_in_both(tmp_path):
    """mem_layer.go:176-181: an absolute link target must lie under the MemFS root (TrimRoot fails otherwise)."""
    ctx = tmp_path / "ctx"
    ctx.mkdir()
    os.symlink("/abs/elsewhere", ctx / "l")
    root = tmp_path / "root"
    root.mkdir()
    with pytest.raises(ValueError):
        lt.MemFS(lambda: 0, str(root)).add_layer_by_copy_ops([lt.CopyOperation.new(["/"], str(ctx), "/", "/x/")])
    with pytest.raises(host.HostError) as ei:
        host.describe_layer(str(root), 0, [host.CopyOperation(["/"], str(ctx), "/", "/x/")])
    assert "trim" in str(ei.value)


# ---- tar ingest (UpdateFromTarReader): random archives in three formats, then random byte damage -----------------
def _random_tar(rng, fmt):
    import io
    import tarfile
    buf = io.BytesIO()
    names = set()
    with tarfile.open(fileobj=buf, mode="w", format=fmt) as tf:
        dirs = [""]
        for _ in range(int(rng.integers(1, 25))):
            base = "".join(rng.choice(list("abcXY01._-é"), size=int(rng.integers(1, 40 if fmt != tarfile.USTAR_FORMAT else 12))))
            if base in (".", "..") or base.startswith(".wh."):
                continue
            parent = dirs[int(rng.integers(0, len(dirs)))]
            name = (parent + "/" if parent else "") + base
            if name in names:
                continue
            names.add(name)
            ti = tarfile.TarInfo(name)
            ti.mtime = int(rng.integers(0, 2**31))
            ti.uid, ti.gid = int(rng.integers(0, 70000)), int(rng.integers(0, 70000))
            ti.mode = int(rng.choice([0o644, 0o755, 0o4755, 0o600, 0o100644, 0o40755]))
            r = rng.random()
            data = b""
            if r < 0.25:
                ti.type, ti.name = tarfile.DIRTYPE, name + "/"
                dirs.append(name)
            elif r < 0.35:
                ti.type, ti.linkname = tarfile.SYMTYPE, "../" * int(rng.integers(0, 3)) + "t" * int(rng.integers(1, 120 if fmt != tarfile.USTAR_FORMAT else 60))
            elif r < 0.45 and names:
                ti.type, ti.linkname = tarfile.LNKTYPE, sorted(names)[int(rng.integers(0, len(names)))]
            elif r < 0.5:
                ti.type = tarfile.FIFOTYPE
            else:
                data = bytes(rng.integers(0, 256, int(rng.integers(0, 3000)), dtype=np.uint8))
                ti.size = len(data)
            try:
                tf.addfile(ti, io.BytesIO(data) if data else None)
            except ValueError:      # USTAR cannot express this member (name too long / non-ASCII): leave it out
                names.discard(name)
    return buf.getvalue()


def _ingest_both(root, data):
    import tempfile
    o_err = h_err = None
    want = got = None
    try:
        # joined text, not lines: a damaged archive can put a newline into a name (e.g. PAX records read as a GNU long name)
        # (the C side prints mode as an unsigned 64-bit octal and truncates the seconds toward zero)
        want = "".join(f"{os.fsdecode(e.hdr.typeflag)} {e.hdr.mode & (2**64 - 1):o} {e.hdr.uid} {e.hdr.gid} {e.hdr.size} "
                       f"{int(e.hdr.mtime_ns / 10**9) if abs(e.hdr.mtime_ns) < 2**62 else (abs(e.hdr.mtime_ns) // 10**9) * (1 if e.hdr.mtime_ns > 0 else -1)} "
                       f"{e.dst} {e.hdr.name} {e.src}\n" for e in lt.MemFS(lambda: 1_600_000_000, root).update_from_tar(data))
        want = want.replace("\n", "")
    except ValueError as e:
        o_err = e
    h = host.MemFS(root)
    with tempfile.TemporaryFile() as f:
        f.write(data)
        f.seek(0)
        try:
            got = "".join(h.describe_update_from_tar(1_600_000_000, f.fileno()))
        except host.HostError as e:
            h_err = e
    h.close()
    return want, got, o_err, h_err


@pytest.mark.parametrize("seed", range(30))
def test_random_tars_ingest_like_the_oracle_and_tarfile(tmp_path, seed):
    import io
    import tarfile
    rng = np.random.default_rng(1000 + seed)
    fmt = [tarfile.USTAR_FORMAT, tarfile.PAX_FORMAT, tarfile.GNU_FORMAT][seed % 3]
    data = _random_tar(rng, fmt)
    ms = lt.read_tar(data)
    infos = tarfile.open(fileobj=io.BytesIO(data)).getmembers()
    assert [m.hdr.name.rstrip("/") for m in ms] == [i.name.rstrip("/") for i in infos]
    assert [(m.hdr.typeflag, m.hdr.uid, m.hdr.gid, m.hdr.mtime_ns // 10**9, m.hdr.linkname) for m in ms] == \
        [(i.type, i.uid, i.gid, int(i.mtime), i.linkname) for i in infos]
    assert [m.data_len for m in ms if m.hdr.typeflag == b"0"] == [i.size for i in infos if i.isreg()]
    root = str(tmp_path)
    want, got, o_err, h_err = _ingest_both(root, data)
    assert o_err is None and h_err is None and got == want
    # damage: flip one byte somewhere in the first members, or cut the stream -- both must agree on accept/reject, and
    # on the result when they accept (a flipped body byte is not the parser's business)
    for k in range(6):
        bad = bytearray(data)
        if k % 2 == 0:
            pos = int(rng.integers(0, min(len(bad), 4096)))
            bad[pos] ^= 1 << int(rng.integers(0, 8))
        else:
            bad = bad[:int(rng.integers(1, len(bad)))]
        want, got, o_err, h_err = _ingest_both(root, bytes(bad))
        assert (o_err is None) == (h_err is None), (k, o_err, h_err)
        if o_err is None:
            assert got == want


@pytest.mark.parametrize("seed", range(24))
def test_random_header_fields_with_repaired_checksum(tmp_path, seed):
    """Overwrite one byte of one header field with an awkward value and REPAIR the checksum, so the damage reaches the
    field parsers (octal vs base-256 numerics, typeflags turning members into PAX / GNU long-name records, magic and
    prefix changes).  Both implementations must agree on accept/reject and on the merged layer."""
    import io
    import tarfile
    rng = np.random.default_rng(9000 + seed)
    data = _random_tar(rng, [tarfile.USTAR_FORMAT, tarfile.PAX_FORMAT, tarfile.GNU_FORMAT][seed % 3])
    offs = [i.offset for i in tarfile.open(fileobj=io.BytesIO(data)).getmembers()]
    if not offs:
        return
    fields = [(0, 100), (100, 108), (108, 116), (116, 124), (124, 136), (136, 148), (156, 157), (157, 257), (257, 265),
              (329, 345), (345, 500)]
    accepted = 0
    for _ in range(20):
        bad = bytearray(data)
        off = offs[int(rng.integers(0, len(offs)))]
        lo, hi = fields[int(rng.integers(0, len(fields)))]
        bad[off + int(rng.integers(lo, hi))] = int(rng.choice([0, 32, 48, 55, 56, 57, 0x80, 0xFF, ord("x"), ord("L"), ord("1"),
                                                               ord("/"), int(rng.integers(0, 256))]))
        blk = bytearray(bad[off:off + 512])
        blk[148:156] = b" " * 8
        blk[148:156] = b"%06o\0 " % sum(blk)
        bad[off:off + 512] = blk
        want, got, o_err, h_err = _ingest_both(str(tmp_path), bytes(bad))
        assert (o_err is None) == (h_err is None), (o_err, h_err)
        if o_err is None:
            assert got == want
            accepted += 1
    assert accepted > 0


# ---- stateful: random sequences of MemFS operations on both implementations ---------------------------------------
@pytest.mark.skipif(os.geteuid() != 0, reason="chown needs root")
@pytest.mark.parametrize("seed", range(14))
def test_random_memfs_operation_sequences(tmp_path, seed):
    """ingest (untar or not) / edit the disk / scan / copy ops, in random order, on the oracle's MemFS and the C++ one
    (each on its own root built by the same seeded steps).  Every step must return the same layer, or fail on both."""
    import shutil
    import tarfile
    import tempfile
    NOW = 1_600_000_000
    T = 1_450_000_000
    logs = []
    for impl in ("oracle", "cpp"):
        rng = np.random.default_rng(31000 + seed)
        root, ctx = tmp_path / impl / "root", tmp_path / impl / "ctx"
        root.mkdir(parents=True)
        ctx.mkdir()
        _random_tree(str(ctx), rng)
        for d, dirs, files in os.walk(ctx):
            for n in dirs + files:
                os.utime(os.path.join(d, n), (T, T), follow_symlinks=False)
        os.utime(ctx, (T, T))
        os.utime(root, (T, T))
        fs = lt.MemFS(lambda: NOW, str(root)) if impl == "oracle" else host.MemFS(str(root))
        log = []
        for step in range(7):
            op = ["ingest", "untar", "edit", "scan", "copy", "scan"][int(rng.integers(0, 6))]
            try:
                if op in ("ingest", "untar"):
                    data = _random_tar(rng, [tarfile.USTAR_FORMAT, tarfile.PAX_FORMAT, tarfile.GNU_FORMAT][int(rng.integers(0, 3))])
                    if impl == "oracle":
                        layer = fs.update_from_tar(data, untar=(op == "untar"))
                        out = [e.dst for e in layer]
                    else:
                        with tempfile.TemporaryFile() as f:
                            f.write(data)
                            f.seek(0)
                            out = [l.split(" ")[6] for l in fs.describe_update_from_tar(NOW, f.fileno(), host.MKHOST_UNTAR if op == "untar" else 0)]
                elif op == "edit":
                    names = sorted(os.listdir(root))
                    k = rng.random()
                    if names and k < 0.4:
                        victim = root / names[int(rng.integers(0, len(names)))]
                        if victim.is_dir() and not victim.is_symlink():
                            shutil.rmtree(victim)
                        else:
                            victim.unlink()
                    else:
                        p = root / ("new%d" % step)
                        p.write_bytes(bytes(rng.integers(0, 256, int(rng.integers(0, 2000)), dtype=np.uint8)))
                        os.chmod(p, 0o640)
                        os.utime(p, (T + step, T + step))
                    os.utime(root, (T + step, T + step))
                    out = ["edited"]
                elif op == "scan":
                    if impl == "oracle":
                        out = [e.dst + ("!" if e.whiteout else "") for e in fs.add_layer_by_scan()]
                    else:
                        out = []
                        for l in fs.describe_scan(NOW):
                            parts = l.split(" ")
                            out.append(parts[6] + ("!" if os.path.basename(parts[7].rstrip("/")).startswith(".wh.") else ""))
                else:
                    dst = ["/app/", "/srv/x/y/", "/new%d" % step][int(rng.integers(0, 3))]
                    srcs = ["/"] if dst.endswith("/") else None
                    if srcs is None:
                        files = [n for n in sorted(os.listdir(ctx)) if (ctx / n).is_file() and not (ctx / n).is_symlink()]
                        if not files:
                            log.append((op, "skip"))
                            continue
                        srcs = ["/" + files[0]]
                    if impl == "oracle":
                        out = [e.dst for e in fs.add_layer_by_copy_ops([lt.CopyOperation.new(srcs, str(ctx), "/", dst, uid=2, gid=3)])]
                    else:
                        out = [l.split(" ")[6] for l in fs.describe_copy_ops(NOW, [host.CopyOperation(srcs, str(ctx), "/", dst, 2, 3)])]
                log.append((op, out))
            except (OSError, ValueError, host.HostError):
                log.append((op, "err"))
            # an aborted untar leaves wall-clock mtimes on directories: pin them so that the next scan cannot depend on
            # whether the two runs crossed a second boundary at different steps
            for d, dirs, _ in os.walk(root):
                for n in dirs:
                    if not os.path.islink(os.path.join(d, n)):
                        os.utime(os.path.join(d, n), (T + 100 + step, T + 100 + step))
        if impl == "cpp":
            fs.close()
        logs.append(log)
    assert logs[0] == logs[1]
