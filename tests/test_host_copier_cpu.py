"""CPU: fileio.Copier / CopyOperation.Execute / evalSymlinks -- the C++ host side (libmkhost) against the oracle's
restatement, on the scenarios of the reference's own tests (lib/fileio/copy_test.go, lib/snapshot/copy_op_test.go:73-)
and a few more.  Both implementations run the same operation on identical source trees; the resulting destination
trees must be identical (type, mode, owner, link target, content).  The deferred mode (regular files written from
buffers after the traversal -- the path the arena packer takes) must give the same tree as the direct one.
Needs root for chown (like the reference's tests); skipped otherwise."""
import os
import shutil
import stat

import numpy as np
import pytest

from makisu_b200 import host
from oracle import copier as oc

pytestmark = pytest.mark.skipif(os.geteuid() != 0, reason="chown needs root")


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
    st = os.lstat(root)
    out["."] = ("d", stat.S_IMODE(st.st_mode), st.st_uid, st.st_gid)
    return out


def _mk(root, rel, data=b"", mode=0o644, uid=0, gid=0):
    p = os.path.join(root, rel)
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "wb") as f:
        f.write(data)
    os.chown(p, uid, gid)
    os.chmod(p, mode)
    return p


def _run_all(tmp_path, build_src, build_dst, srcs, dst, work_dir="/", uid=0, gid=0, mode=0, blacklist=()):
    """Same op through oracle / C++ direct / C++ deferred, each on its own copy of the scenario."""
    trees = []
    for impl in ("oracle", "cpp", "cpp_deferred"):
        base = tmp_path / impl
        src_root, dst_root = base / "src", base / "dst"
        src_root.mkdir(parents=True)
        dst_root.mkdir(parents=True)
        os.chmod(dst_root, 0o755)
        build_src(str(src_root))
        if build_dst:
            build_dst(str(dst_root))
        d = dst.replace("$DST", str(dst_root))
        wd = work_dir.replace("$DST", str(dst_root))
        bl = [b.replace("$SRC", str(src_root)) for b in blacklist]
        if impl == "oracle":
            rd = d
            if not os.path.isabs(rd):
                rd = os.path.normpath(os.path.join(wd, rd)) + ("/" if d.endswith("/") or d in (".", "..") else "")
            oc.execute_copy_op(str(src_root), srcs, rd, uid, gid, bool(mode & host.MKHOST_COPY_CHOWN),
                               bool(mode & host.MKHOST_COPY_INTERNAL), bool(mode & host.MKHOST_COPY_PRESERVE_OWNER), bl)
        else:
            m = mode | (host.MKHOST_COPY_DEFERRED if impl == "cpp_deferred" else 0)
            host.copy_op_execute(host.CopyOperation(srcs, str(src_root), wd, d, uid, gid), m, bl)
        trees.append(_tree(str(dst_root)))
    assert trees[0] == trees[1], "C++ differs from the oracle"
    assert trees[1] == trees[2], "deferred mode differs from the direct one"
    return trees[0]


def test_copy_file_scenarios_of_copy_test_go(tmp_path):
    # TestCopyFileTargetNotExist / TargetEmpty / TargetOverwrite / SetSpecialBit / DanglingSymlink (copy_test.go:41-166)
    def src(r):
        _mk(r, "test.txt", b"Testing COPY", 0o777 | stat.S_ISUID, uid=7, gid=8)

    def dst(r):
        _mk(r, "existing.txt", b"Test target file one, longer than the source", 0o600, uid=5, gid=6)
        _mk(r, "wasfile", b"x")

    t = _run_all(tmp_path / "a", src, dst, ["/test.txt"], "$DST/new/dir/out.txt")
    assert t["new/dir/out.txt"] == ("f", 0o777 | stat.S_ISUID, 0, 0, b"Testing COPY")      # from context: owner root, mode kept
    assert t["new"] == ("d", 0o755, 0, 0) and t["new/dir"] == ("d", 0o755, 0, 0)
    t = _run_all(tmp_path / "b", src, dst, ["/test.txt"], "$DST/existing.txt")
    assert t["existing.txt"] == ("f", 0o777 | stat.S_ISUID, 0, 0, b"Testing COPY")          # truncated, overwritten
    t = _run_all(tmp_path / "d", src, dst, ["/test.txt"], "$DST/sub/", uid=11, gid=12, mode=host.MKHOST_COPY_CHOWN)
    assert t["sub"] == ("d", 0o755, 11, 12) and t["sub/test.txt"][1:4] == (0o777 | stat.S_ISUID, 11, 12)
    t = _run_all(tmp_path / "e", src, dst, ["/test.txt"], "$DST/keep.txt", mode=host.MKHOST_COPY_INTERNAL)
    assert t["keep.txt"][1:4] == (0o777 | stat.S_ISUID, 7, 8)                               # --from: owners preserved


def test_copy_directory_scenarios(tmp_path):
    # TestCopyDirectoryTargetNotExist / TargetExists / IncludingSymlink (copy_test.go:168-305) + owner rules (copy.go:38-66)
    def src(r):
        _mk(r, "one/f1", b"Test source file one", 0o640, uid=3, gid=4)
        _mk(r, "two/.keep", b"", 0o600)
        _mk(r, "f2", b"Test source file two", 0o755)
        os.chmod(os.path.join(r, "one"), 0o750 | stat.S_ISGID)
        os.chown(os.path.join(r, "one"), 9, 10)
        os.symlink("one", os.path.join(r, "link"))
        os.symlink("/nonexistent", os.path.join(r, "dangling"))                             # copied as-is (TestCopyFileDanglingSymlink)
        os.mkfifo(os.path.join(r, "fifo"))                                                  # special: ignored

    def dst(r):
        _mk(r, "t/target_only", b"Test target file one", 0o600, uid=1, gid=2)
        os.chmod(os.path.join(r, "t"), 0o700)
        os.chown(os.path.join(r, "t"), 21, 22)
        _mk(r, "t/one/f1", b"old content that is longer", 0o400)
        _mk(r, "t/dangling", b"a file where the source has a link")

    t = _run_all(tmp_path / "a", src, None, ["/"], "$DST/fresh/")
    assert t["fresh"] == ("d", 0o755, 0, 0) and t["fresh/one"] == ("d", 0o750 | stat.S_ISGID, 0, 0)
    assert t["fresh/one/f1"] == ("f", 0o640, 0, 0, b"Test source file one") and t["fresh/link"] == ("l", "one")
    assert "fresh/fifo" not in t and t["fresh/dangling"] == ("l", "/nonexistent")
    t = _run_all(tmp_path / "b", src, dst, ["/"], "$DST/t")
    assert t["t"] == ("d", 0o700, 21, 22)                                                   # existing target dir: untouched
    assert t["t/target_only"][4] == b"Test target file one" and t["t/one/f1"] == ("f", 0o640, 0, 0, b"Test source file one")
    assert t["t/dangling"] == ("l", "/nonexistent")                                         # the existing file is replaced by the link
    t = _run_all(tmp_path / "c", src, dst, ["/"], "$DST/t", uid=30, gid=31, mode=host.MKHOST_COPY_CHOWN)
    assert t["t"] == ("d", 0o700, 21, 22) and t["t/one"][2:] == (30, 31) and t["t/f2"][2:4] == (30, 31)
    t = _run_all(tmp_path / "d", src, None, ["/"], "$DST/arch/", mode=host.MKHOST_COPY_INTERNAL | host.MKHOST_COPY_PRESERVE_OWNER)
    assert t["arch/one"] == ("d", 0o750 | stat.S_ISGID, 9, 10) and t["arch/one/f1"][2:4] == (3, 4)
    t = _run_all(tmp_path / "e", src, None, ["/"], "$DST/bl/", blacklist=["$SRC/one", "$SRC/f2"])
    assert "bl/one" not in t and "bl/f2" not in t and "bl/two/.keep" in t
    t = _run_all(tmp_path / "f", src, None, ["/"], "$DST/bl/", blacklist=["$SRC/one"], mode=host.MKHOST_COPY_INTERNAL)
    assert "bl/one/f1" in t                                                                 # --from: no blacklist (copy_op.go:94-98)


def test_execute_copy_operation_shapes(tmp_path):
    # copy_op_test.go:73-: file -> file, file -> relative file, files -> dir, dir -> dir, relative dst against workdir
    def src(r):
        _mk(r, "test.txt", b"hello", 0o777)
        _mk(r, "test2.txt", b"hello2", 0o777)
        _mk(r, "dir/sub/x", b"x")

    t = _run_all(tmp_path / "a", src, None, ["/test.txt"], "$DST/test2/test.txt", uid=1, gid=1, mode=host.MKHOST_COPY_CHOWN)
    assert t["test2/test.txt"][4] == b"hello"
    t = _run_all(tmp_path / "b", src, None, ["/test.txt"], "test2/test.txt", work_dir="$DST")
    assert t["test2/test.txt"][4] == b"hello"
    t = _run_all(tmp_path / "c", src, None, ["/test.txt", "/test2.txt"], "test2/", work_dir="$DST")
    assert t["test2/test.txt"][4] == b"hello" and t["test2/test2.txt"][4] == b"hello2"
    t = _run_all(tmp_path / "d", src, None, ["/dir"], "$DST/out")
    assert t["out/sub/x"][4] == b"x"
    t = _run_all(tmp_path / "e", src, None, ["/dir", "/test.txt"], "$DST/mix/")
    assert t["mix/sub/x"][4] == b"x" and t["mix/test.txt"][4] == b"hello"                   # a dir source spills its CONTENTS


def test_eval_symlinks_and_errors(tmp_path):
    def src(r):
        _mk(r, "real/data.txt", b"D")
        os.symlink("real", os.path.join(r, "alias"))                                        # relative dir link
        os.symlink(os.path.join(r, "real", "data.txt"), os.path.join(r, "abs_in_root"))     # absolute, inside the root
        os.symlink("/etc/hostname", os.path.join(r, "escape"))                              # absolute, outside

    t = _run_all(tmp_path / "a", src, None, ["/alias/data.txt"], "$DST/o/")
    assert t["o/data.txt"][4] == b"D"
    t = _run_all(tmp_path / "b", src, None, ["/abs_in_root"], "$DST/o/")
    assert t["o/data.txt"][4] == b"D"                                                       # resolved to the target file
    t = _run_all(tmp_path / "c", src, None, ["/alias"], "$DST/o2/")
    assert t["o2/data.txt"][4] == b"D"                                                      # link to a dir: the dir is copied
    for bad, msg, omsg in [(["/escape"], "link points outside of root", "link points outside of root"),
                           (["/missing"], "lstat", "No such file")]:
        base = tmp_path / ("bad" + bad[0].strip("/"))
        (base / "src").mkdir(parents=True)
        (base / "dst").mkdir()
        src(str(base / "src"))
        with pytest.raises(OSError) as eo:
            oc.execute_copy_op(str(base / "src"), bad, str(base / "dst") + "/", 0, 0, False, False, False, [])
        with pytest.raises(host.HostError) as eh:
            host.copy_op_execute(host.CopyOperation(bad, str(base / "src"), "/", str(base / "dst") + "/"))
        assert omsg in str(eo.value) and msg in str(eh.value)
    with pytest.raises(host.HostError) as e:
        host.copy_op_execute(host.CopyOperation(["/x"], str(tmp_path), "/", str(tmp_path) + "/"),
                             host.MKHOST_COPY_CHOWN | host.MKHOST_COPY_PRESERVE_OWNER)
    assert "both chown and archive" in str(e.value)


def test_infinite_loop_guard(tmp_path):
    # TestCopyDirectoryInfiniteLoop (copy_test.go:307-): the target is a child of the source
    for impl in ("oracle", "cpp"):
        s = tmp_path / impl / "src"
        (s / "sub").mkdir(parents=True)
        (s / "target").mkdir()
        _mk(str(s), "sub/f1", b"one")
        _mk(str(s), "f2", b"two")
        if impl == "oracle":
            oc.execute_copy_op(str(s), ["/"], str(s / "target"), 0, 0, False, False, False, [])
        else:
            host.copy_op_execute(host.CopyOperation(["/"], str(s), "/", str(s / "target")))
    a, b = _tree(str(tmp_path / "oracle" / "src" / "target")), _tree(str(tmp_path / "cpp" / "src" / "target"))
    assert a == b and a["sub/f1"][4] == b"one" and "target" not in a
    shutil.rmtree(tmp_path)


def _random_tree(root, rng, depth=0, fifos=True):
    for _ in range(int(rng.integers(1, 7))):
        nm = "".join(rng.choice(list("abAB01._-"), size=int(rng.integers(1, 9))))
        p = os.path.join(root, nm)
        if nm in (".", "..") or os.path.lexists(p):
            continue
        r = rng.random()
        if r < 0.3 and depth < 3:
            os.mkdir(p)
            os.chmod(p, int(rng.choice([0o755, 0o700, 0o2775, 0o1777])))
            os.chown(p, int(rng.integers(0, 5)), int(rng.integers(0, 5)))
            _random_tree(p, rng, depth + 1, fifos)
        elif r < 0.42:
            os.symlink("some/target" if rng.random() < 0.5 else "../up", p)
        elif r < 0.47 and fifos:
            os.mkfifo(p)
        else:
            with open(p, "wb") as f:
                f.write(bytes(rng.integers(0, 256, int(rng.integers(0, 3000)), dtype=np.uint8)))
            os.chown(p, int(rng.integers(0, 5)), int(rng.integers(0, 5)))
            os.chmod(p, int(rng.choice([0o644, 0o600, 0o755, 0o4755, 0o2711])))


@pytest.mark.parametrize("seed", range(16))
def test_random_trees_onto_random_destinations(tmp_path, seed):
    """Random source tree copied onto a random, partly colliding destination tree with a random owner mode: the three
    runs (oracle, C++, C++ deferred) must leave identical trees, or fail alike (a directory landing on a file, ...)."""
    mode = [0, host.MKHOST_COPY_CHOWN, host.MKHOST_COPY_INTERNAL, host.MKHOST_COPY_INTERNAL | host.MKHOST_COPY_PRESERVE_OWNER][seed % 4]
    trees, errs = [], []
    for impl in ("oracle", "cpp", "cpp_deferred"):
        rng = np.random.default_rng(700 + seed)                        # identical scenario for every implementation
        base = tmp_path / impl
        src, dst = base / "src", base / "dst"
        src.mkdir(parents=True)
        dst.mkdir()
        _random_tree(str(src), rng)
        _random_tree(str(dst), rng, fifos=False)                        # same name alphabet => collisions happen (no FIFOs: opening one blocks)
        try:
            if impl == "oracle":
                oc.execute_copy_op(str(src), ["/"], str(dst) + "/", 3, 4, bool(mode & host.MKHOST_COPY_CHOWN),
                                   bool(mode & host.MKHOST_COPY_INTERNAL), bool(mode & host.MKHOST_COPY_PRESERVE_OWNER), [])
            else:
                m = mode | (host.MKHOST_COPY_DEFERRED if impl == "cpp_deferred" else 0)
                host.copy_op_execute(host.CopyOperation(["/"], str(src), "/", str(dst) + "/", 3, 4), m)
            errs.append(None)
        except (OSError, host.HostError) as e:
            errs.append(e)
        trees.append(_tree(str(dst)))
    assert [e is None for e in errs] == [errs[0] is None] * 3, errs
    if errs[0] is None:
        assert trees[0] == trees[1] == trees[2]
    else:
        assert trees[0] == trees[1]                                     # both stop at the same entry (same traversal order)
