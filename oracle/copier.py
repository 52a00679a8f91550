"""ORACLE (test infrastructure only): the file copy a COPY/ADD step performs when it modifies the file system.

Restates
  reference lib/fileio/copy.go:30-400        (Copier: CopyFile, CopyDir, copyFile, copyRegularFile, copySymlink,
                                              copyDirContents, copyDir, mkdirAll and the owner rules in its header)
  reference lib/snapshot/copy_op.go:82-147   (CopyOperation.Execute: which Copier a step gets)
  reference lib/snapshot/utils.go:249-324    (evalSymlinks / walkLink / walkLinks)
Pure file I/O, no arithmetic.  Pinned by the scenarios of the reference's own tests (lib/fileio/copy_test.go,
lib/snapshot/copy_op_test.go:73-), which tests/test_host_copier_cpu.py replays; beyond those "parity unpinned".
"""
from __future__ import annotations

import os
import posixpath
import stat
from dataclasses import dataclass
from typing import List, Optional

from .ctx_crc import is_special_file
from .layer_tar import abs_path, go_clean, is_descendant_of_any, rel_path


@dataclass
class Owner:
    uid: int
    gid: int
    overwrite: bool


class Copier:
    def __init__(self, blacklist: List[str], dst_dir_owner: Optional[Owner] = None,
                 dst_file_and_children_owner: Optional[Owner] = None):
        self.blacklist = list(blacklist)
        self.dst_dir_owner = dst_dir_owner
        self.children_owner = dst_file_and_children_owner

    def _blacklisted(self, p: str) -> bool:
        return is_descendant_of_any(p, self.blacklist)

    # copy.go:122-131
    def copy_file(self, source: str, target: str) -> None:
        self._mkdir_all(posixpath.dirname(go_clean(target)) or "/")
        self._copy_file(source, target)

    # copy.go:142-156
    def copy_dir(self, source: str, target: str) -> None:
        if self._blacklisted(source):
            return
        self._mkdir_all(target)
        self._copy_dir_contents(source, target, target)

    # copy.go:163-193
    def _copy_file(self, src: str, dst: str) -> None:
        st = os.lstat(src)
        if self._blacklisted(src):
            pass  # the reference only logs here and carries on (no return in that branch)
        elif is_special_file(st):
            return
        if stat.S_ISLNK(st.st_mode):
            self._copy_symlink(src, dst)
            return
        if os.path.lexists(dst):
            os.chmod(dst, 0o777)
        self._copy_regular(st, src, dst)

    # copy.go:195-230
    def _copy_regular(self, st: os.stat_result, src: str, dst: str) -> None:
        with open(src, "rb") as r:
            fd = os.open(dst, os.O_WRONLY | os.O_CREAT, 0o777)
            try:
                os.truncate(dst, 0)
                while True:
                    buf = r.read(1 << 20)
                    if not buf:
                        break
                    os.write(fd, buf)
            finally:
                os.close(fd)
        uid, gid = st.st_uid, st.st_gid
        if self.children_owner is not None and self.children_owner.overwrite:
            uid, gid = self.children_owner.uid, self.children_owner.gid
        os.chown(dst, uid, gid)
        os.chmod(dst, stat.S_IMODE(st.st_mode))

    # copy.go:232-249
    def _copy_symlink(self, src: str, dst: str) -> None:
        if os.path.lexists(dst):
            os.remove(dst)
        os.symlink(os.readlink(src), dst)

    # copy.go:252-283
    def _copy_dir_contents(self, src: str, dst: str, orig_dst: str) -> None:
        for name in sorted(os.listdir(src), key=os.fsencode):  # ioutil.ReadDir sorts by name
            cur_src = posixpath.join(src, name)
            if self._blacklisted(cur_src):
                continue
            if cur_src == orig_dst:
                continue  # silently break the infinite loop
            cur_dst = posixpath.join(dst, name)
            if stat.S_ISDIR(os.lstat(cur_src).st_mode):
                self._copy_dir(cur_src, cur_dst)
                self._copy_dir_contents(cur_src, cur_dst, orig_dst)
            else:
                self._copy_file(cur_src, cur_dst)

    # copy.go:286-329
    def _copy_dir(self, src: str, dst: str) -> None:
        st = os.lstat(src)
        if not stat.S_ISDIR(st.st_mode):
            raise OSError("source %s is not a directory" % src)
        if self._blacklisted(src):
            return
        if not os.path.lexists(dst):
            os.mkdir(dst, stat.S_IMODE(st.st_mode))
        elif not stat.S_ISDIR(os.lstat(dst).st_mode):
            raise OSError("dst is not a directory")
        os.chmod(dst, stat.S_IMODE(st.st_mode))
        uid, gid = st.st_uid, st.st_gid
        if self.children_owner is not None and self.children_owner.overwrite:
            uid, gid = self.children_owner.uid, self.children_owner.gid
        os.chown(dst, uid, gid)

    # copy.go:334-399
    def _mkdir_all(self, dst: str) -> None:
        if dst == "":
            raise OSError("empty dst directory")
        a = go_clean(dst)
        if not a.startswith("/"):
            a = go_clean(posixpath.join(os.getcwd(), a))
        parts = a.split("/")
        parts[0] = "/"
        prev = ""
        for d in parts[:-1]:
            cur = posixpath.join(prev, d)
            if not os.path.lexists(cur):
                os.mkdir(cur, 0o755)      # subject to the umask, like os.Mkdir in Go
                os.chown(cur, 0, 0)
            prev = cur
        if not os.path.lexists(a):
            os.mkdir(a, 0o755)
            if self.dst_dir_owner is not None:
                os.chown(a, self.dst_dir_owner.uid, self.dst_dir_owner.gid)
            else:
                os.chown(a, 0, 0)
        elif self.dst_dir_owner is not None and self.dst_dir_owner.overwrite:
            os.chown(a, self.dst_dir_owner.uid, self.dst_dir_owner.gid)


# ---- lib/snapshot/utils.go:249-324 ---------------------------------------------------------------
def _walk_link(path: str, root: str, walked: List[int]):
    if walked[0] > 255:
        raise OSError("eval symlinks: too many links")
    full = go_clean(root + "/" + path)  # filepath.Join(root, path): an absolute `path` stays UNDER root
    st = os.lstat(full)
    if not stat.S_ISLNK(st.st_mode):
        return path, False
    new = os.readlink(full)
    if not new.startswith(root) and new.startswith("/"):
        raise OSError("link points outside of root: %s -> %s" % (full, new))
    walked[0] += 1
    if new.startswith(root):
        new = new[len(root):]
    return new, True


def _go_split(path: str):
    i = path.rfind("/")
    return path[:i + 1], path[i + 1:]


def _walk_links(path: str, root: str, walked: List[int]) -> str:
    d, f = _go_split(path)
    if d == "":
        return _walk_link(f, root, walked)[0]
    if f == "":
        if d.rstrip("/") == root.rstrip("/"):
            return d
        return _walk_links(d[:-1], root, walked)
    newdir = _walk_links(d, root, walked)
    newpath, islink = _walk_link(go_clean(newdir + "/" + f) if newdir else f, root, walked)
    if not islink:
        return newpath
    if newpath.startswith("/"):
        return newpath
    return go_clean(newdir + "/" + newpath) if newdir else go_clean(newpath)


def eval_symlinks(p: str, src_root: str) -> str:
    if p == "":
        return p
    walked = [0]
    while True:
        i = walked[0]
        new = _walk_links(p, src_root, walked)
        if i == walked[0]:
            return abs_path(new)
        p = new


# ---- lib/snapshot/copy_op.go:82-147 --------------------------------------------------------------
def execute_copy_op(src_root: str, srcs: List[str], dst: str, uid: int, gid: int, chown: bool, internal: bool,
                    preserve_owner: bool, blacklist: List[str]) -> None:
    """dst: already resolved against the working directory (NewCopyOperation), dir format kept."""
    for src in srcs:
        src = eval_symlinks(rel_path(src), src_root)
        src = go_clean(posixpath.join(src_root, src.lstrip("/")) if src else src_root)
        st = os.lstat(src)
        bl = [] if internal else blacklist
        if chown:
            c = Copier(bl, Owner(uid, gid, False), Owner(uid, gid, True))
        elif not internal:
            c = Copier(bl, Owner(0, 0, False), Owner(0, 0, True))
        elif preserve_owner:
            c = Copier(bl, Owner(st.st_uid, st.st_gid, False))
        else:
            c = Copier(bl)
        if stat.S_ISDIR(st.st_mode):
            c.copy_dir(src, dst)
        elif dst.endswith("/") or dst in (".", ".."):
            c.copy_file(src, go_clean(posixpath.join(dst, posixpath.basename(src))))
        else:
            c.copy_file(src, dst)
