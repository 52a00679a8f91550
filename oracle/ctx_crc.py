"""ORACLE (test infrastructure only): the build-context fingerprint ("cacheID").

Restates, in reference order, what feeds crc32.NewIEEE() in
  reference lib/builder/step/add_copy_step.go:102-122  (SetCacheID)
  reference lib/builder/step/add_copy_step.go:153-184  (calculateContextChecksum, resolveFromPaths)
  reference lib/builder/step/add_copy_step.go:194-238  (checksumPathContents)
  reference lib/builder/step/base_step.go:62-67        (baseStep.SetCacheID)
  reference lib/builder/step/from_step.go:79-83        (fromStep.SetCacheID)
  reference lib/builder/build_plan.go:96-97            (plan seed)
and Go 1.14 path/filepath.Walk / Glob / Match / Rel semantics (stdlib, not vendored).

Parity: the reference has NO golden cacheID anywhere (copy_step_test.go:51-169 is
relational only) => "parity unpinned"; the arithmetic (CRC-32/IEEE) is pinned by
zlib.crc32 and the check value 0xcbf43926, the byte ORDER only by this restatement.
"""
from __future__ import annotations

import os
import re
import stat
import zlib
from dataclasses import dataclass
from typing import Callable, Iterator, List, Optional, Tuple

BUILD_HASH_DEFAULT = "master-unreleased"  # reference lib/utils/constants.go:23-27


class SkipDir(Exception):
    pass


def _sorted_names(path: str) -> List[str]:
    # Go: readDirNames + sort.Strings => bytewise order of the names
    return sorted(os.listdir(path), key=os.fsencode)


def go_walk(root: str, fn: Callable[[str, os.stat_result], Optional[str]]) -> None:
    """path/filepath.Walk (go1.14): pre-order, lexical order, Lstat (no symlink following).
    fn returns None, or "skipdir" (filepath.SkipDir)."""

    def walk(path: str, st: os.stat_result) -> Optional[str]:
        if not stat.S_ISDIR(st.st_mode):
            return fn(path, st)
        names = _sorted_names(path)
        r = fn(path, st)
        if r is not None:
            return r
        for name in names:
            filename = os.path.join(path, name)
            fst = os.lstat(filename)
            r = walk(filename, fst)
            if r is not None:
                if not stat.S_ISDIR(fst.st_mode) or r != "skipdir":
                    return r
        return None

    st = os.lstat(root)
    r = walk(root, st)
    if r == "skipdir":
        return


def is_special_file(st: os.stat_result) -> bool:
    """reference lib/utils/utils.go:161-163"""
    m = st.st_mode
    return stat.S_ISCHR(m) or stat.S_ISBLK(m) or stat.S_ISFIFO(m) or stat.S_ISSOCK(m)


# ---- filepath.Match / Glob (go1.14) -------------------------------------------------
def _match_to_regex(pat: str) -> re.Pattern:
    i, out = 0, []
    while i < len(pat):
        c = pat[i]
        if c == "*":
            out.append("[^/]*")
        elif c == "?":
            out.append("[^/]")
        elif c == "[":
            j = i + 1
            neg = j < len(pat) and pat[j] == "^"
            if neg:
                j += 1
            cls = []
            while j < len(pat) and pat[j] != "]":
                if pat[j] == "\\" and j + 1 < len(pat):
                    j += 1
                cls.append(re.escape(pat[j]) if pat[j] != "-" else "-")
                j += 1
            out.append("[" + ("^" if neg else "") + "".join(cls) + "]")
            i = j
        elif c == "\\" and i + 1 < len(pat):
            i += 1
            out.append(re.escape(pat[i]))
        else:
            out.append(re.escape(c))
        i += 1
    return re.compile("^" + "".join(out) + "$", re.S)


def _has_meta(p: str) -> bool:
    return any(ch in p for ch in "*?[\\")


def go_glob(pattern: str) -> List[str]:
    if not _has_meta(pattern):
        return [pattern] if os.path.lexists(pattern) else []
    d, f = os.path.split(pattern)
    d = d or "."
    if d == pattern:  # prevent infinite recursion
        return []
    dirs = [d] if not _has_meta(d) else go_glob(d)
    rx = _match_to_regex(f)
    out: List[str] = []
    for dd in dirs:
        if not os.path.isdir(dd):
            continue
        for n in _sorted_names(dd):
            if rx.match(n):
                out.append(os.path.join(dd, n))
    return out


# ---- the stream ---------------------------------------------------------------------
@dataclass
class Segment:
    """One piece of the CRC stream: literal bytes, or the content of a file on disk."""
    kind: str            # "bytes" | "file"
    data: bytes = b""
    path: str = ""
    size: int = 0


def resolve_from_paths(context_dir: str, from_paths: List[str]) -> List[str]:
    """add_copy_step.go:171-184"""
    sources: List[str] = []
    for source in from_paths:
        source = os.path.normpath(os.path.join(context_dir, source))  # filepath.Join cleans
        matches = go_glob(source)
        sources.extend(matches if matches else [source])
    return sources


def context_segments(context_dir: str, from_paths: List[str]) -> Iterator[Segment]:
    """add_copy_step.go:153-169,194-238: relpath, then link target / content, no separators."""
    segs: List[Segment] = []

    def visit(path: str, st: os.stat_result) -> Optional[str]:
        if is_special_file(st):
            return "skipdir" if stat.S_ISDIR(st.st_mode) else None
        rel = os.path.relpath(path, context_dir)  # filepath.Rel
        segs.append(Segment("bytes", os.fsencode(rel)))
        if stat.S_ISDIR(st.st_mode):
            return None
        if stat.S_ISLNK(st.st_mode):
            segs.append(Segment("bytes", os.fsencode(os.readlink(path))))
            return None
        segs.append(Segment("file", path=path, size=st.st_size))
        return None

    for source in resolve_from_paths(context_dir, from_paths):
        go_walk(source, visit)
    return iter(segs)


def crc_hex(crc: int) -> str:
    return "%x" % crc  # fmt.Sprintf("%x", Sum32()): lower case, no zero padding


def copy_step_cache_id(seed: str, directive: str, args: str, context_dir: str, from_paths: List[str],
                       from_stage: str = "") -> str:
    """addCopyStep.SetCacheID (add_copy_step.go:102-122)."""
    crc = zlib.crc32((seed + directive + args).encode())
    if not from_stage:
        for seg in context_segments(context_dir, from_paths):
            if seg.kind == "bytes":
                crc = zlib.crc32(seg.data, crc)
            else:
                with open(seg.path, "rb") as fh:
                    while True:
                        buf = fh.read(32 * 1024)  # io.Copy buffer
                        if not buf:
                            break
                        crc = zlib.crc32(buf, crc)
    return crc_hex(crc)


def base_step_cache_id(seed: str, directive: str, args: str, commit: bool) -> str:
    """baseStep.SetCacheID (base_step.go:62-67): seed+directive+args+fmt("%v", commit)."""
    return crc_hex(zlib.crc32((seed + directive + args + ("true" if commit else "false")).encode()))


def from_step_cache_id(seed: str, image: str) -> str:
    """fromStep.SetCacheID (from_step.go:79-83): crc32(seed + "FROM" + image)."""
    return crc_hex(zlib.crc32((seed + "FROM" + image).encode()))


def plan_seed(force_commit: bool, allow_modify_fs: bool, build_hash: str = BUILD_HASH_DEFAULT) -> str:
    """build_plan.go:96-97: crc32(BuildHash + fmt.Sprintf("%v", *opts)), opts printed as &{force modify}."""
    opts = "&{%s %s}" % ("true" if force_commit else "false", "true" if allow_modify_fs else "false")
    return crc_hex(zlib.crc32((build_hash + opts).encode()))
