"""ORACLE (test infrastructure only): the uncompressed layer tar stream whose
SHA-256 is DigestPair.TarDigest.

Restates
  reference lib/snapshot/mem_fs.go:69-83,276-289,353-433,440-569  (MemFS, AddLayerByCopyOps, addToLayer,
            commitLayer, maybeAddToLayer, isUpdated, addAncestors)
  reference lib/snapshot/mem_fs.go:165-255,574-716                (UpdateFromTarReader, untarOneItem and helpers)
  reference lib/tario/apply.go:26-49                              (ApplyHeader)
  reference lib/snapshot/mem_layer.go:83-88,127-132,152-244       (commit, whiteout, createHeader, addHeader, rangeFiles)
  reference lib/snapshot/copy_op.go:45-80,149-174                 (NewCopyOperation, resolveDestination)
  reference lib/snapshot/utils.go:37-75                           (shouldSkip, walk)
  reference lib/tario/write.go:28-68                              (WriteEntry, WriteHeader)
  reference lib/tario/compare.go:24-120                           (IsSimilarHeader)
  reference lib/pathutils/path.go:24-68
and the header writer of Go 1.14 archive/tar (stdlib, not vendored): FileInfoHeader, Writer.WriteHeader,
allowedFormats, writeUSTARHeader, writePAXHeader, splitUSTARPath, formatOctal/formatString, Close.

Parity: the USTAR field formats are pinned by re-encoding every header of the Go-written fixture
reference testdata/files/busybox/393c.../layer.tar byte for byte (tests/test_oracle_golden.py); the
empty-archive trailer by image.DigestEmptyTar (lib/docker/image/const_darwin.go:18).  PAX output and
makisu's entry ORDER / header VALUES have no golden bytes in the reference ("parity unpinned" beyond the
relational TestAddLayersEqual, mem_fs_test.go:1118).
"""
from __future__ import annotations

import hashlib
import os
import posixpath
import stat
from dataclasses import dataclass, field, replace
from typing import Callable, Dict, Iterator, List, Optional, Tuple

from .ctx_crc import go_walk, is_special_file

TYPE_REG, TYPE_LINK, TYPE_SYMLINK, TYPE_CHAR, TYPE_BLOCK, TYPE_DIR, TYPE_FIFO = b"0", b"1", b"2", b"3", b"4", b"5", b"6"
TYPE_XHEADER = b"x"
WHITEOUT_PREFIX = ".wh."
WHITEOUT_META_PREFIX = ".wh..wh."

C_ISUID, C_ISGID, C_ISVTX = 0o4000, 0o2000, 0o1000
# os.FileMode bits (go1.14 os/types.go)
GO_MODE_DIR, GO_MODE_SYMLINK, GO_MODE_DEVICE, GO_MODE_NAMED_PIPE, GO_MODE_SOCKET = 1 << 31, 1 << 27, 1 << 26, 1 << 25, 1 << 24
GO_MODE_SETUID, GO_MODE_SETGID, GO_MODE_CHAR_DEVICE, GO_MODE_STICKY = 1 << 23, 1 << 22, 1 << 21, 1 << 20


def go_chmod_bits(file_mode: int) -> int:
    """os.Chmod / os.Mkdir / os.OpenFile take an os.FileMode; syscallMode keeps the permission bits and maps
    ModeSetuid / ModeSetgid / ModeSticky back to 04000 / 02000 / 01000 (go1.14 os/file_posix.go)."""
    m = file_mode & 0o777
    if file_mode & GO_MODE_SETUID:
        m |= C_ISUID
    if file_mode & GO_MODE_SETGID:
        m |= C_ISGID
    if file_mode & GO_MODE_STICKY:
        m |= C_ISVTX
    return m


# ---- pathutils (lib/pathutils/path.go) ----------------------------------------------
def go_clean(p: str) -> str:
    """path.Clean (go1.14 path/path.go) for the inputs this module sees."""
    if p == "":
        return "."
    rooted = p.startswith("/")
    out: List[str] = []
    for e in p.split("/"):
        if e in ("", "."):
            continue
        if e == "..":
            if out and out[-1] != "..":
                out.pop()
            elif not rooted:
                out.append("..")
            continue
        out.append(e)
    s = "/".join(out)
    return "/" + s if rooted else (s or ".")


def abs_path(p: str) -> str:
    """pathutils.AbsPath: path.Join("/", strings.TrimRight(p, "/")) -- Join cleans the result."""
    return go_clean("/" + p.rstrip("/"))


def rel_path(p: str) -> str:
    return p.lstrip("/")


def split_path(p: str) -> List[str]:
    t = p.strip("/")
    return t.split("/") if t else []


def trim_root(p: str, root: str) -> str:
    if not p.startswith(root):
        raise ValueError(f"failed to trim root prefix {root} from path {p}")
    return abs_path(p[len(root):])


def is_descendant_of_any(p: str, ancestors: List[str]) -> bool:
    p = abs_path(p)
    for a in ancestors:
        a = abs_path(a)
        if p == a or a == "/" or (posixpath.dirname(p) + "/").startswith(a + "/"):
            return True
    return False


# ---- tar header ---------------------------------------------------------------------
@dataclass
class Header:
    name: str = ""
    mode: int = 0
    uid: int = 0
    gid: int = 0
    size: int = 0
    mtime_ns: int = 0           # ModTime in integer nanoseconds since the epoch
    typeflag: bytes = TYPE_REG
    linkname: str = ""
    uname: str = ""
    gname: str = ""
    devmajor: int = 0
    devminor: int = 0

    def file_mode_bits(self) -> int:
        """Header.FileInfo().Mode() (go1.14 archive/tar common.go headerFileInfo.Mode) as an os.FileMode integer:
        permission bits, setuid/setgid/sticky, and the type bits implied by BOTH the c_IS* bits a foreign writer
        left in Mode and the typeflag."""
        m = self.mode & 0xFFFFFFFF          # os.FileMode(h.Mode): uint32
        mode = m & 0o777
        if m & C_ISUID:
            mode |= GO_MODE_SETUID
        if m & C_ISGID:
            mode |= GO_MODE_SETGID
        if m & C_ISVTX:
            mode |= GO_MODE_STICKY
        mode |= {0o40000: GO_MODE_DIR, 0o10000: GO_MODE_NAMED_PIPE, 0o120000: GO_MODE_SYMLINK, 0o60000: GO_MODE_DEVICE,
                 0o20000: GO_MODE_DEVICE | GO_MODE_CHAR_DEVICE, 0o140000: GO_MODE_SOCKET}.get(m & ~0o7777 & 0xFFFFFFFF, 0)
        mode |= {TYPE_SYMLINK: GO_MODE_SYMLINK, TYPE_CHAR: GO_MODE_DEVICE | GO_MODE_CHAR_DEVICE, TYPE_BLOCK: GO_MODE_DEVICE,
                 TYPE_DIR: GO_MODE_DIR, TYPE_FIFO: GO_MODE_NAMED_PIPE}.get(self.typeflag, 0)
        return mode

    def is_special(self) -> bool:
        """utils.IsSpecialFile(hdr.FileInfo()) (lib/utils/utils.go:161-163)."""
        return bool(self.file_mode_bits() & (GO_MODE_CHAR_DEVICE | GO_MODE_DEVICE | GO_MODE_NAMED_PIPE | GO_MODE_SOCKET))


def file_info_header(st: os.stat_result, link: str = "") -> Header:
    """tar.FileInfoHeader (go1.14) + stat_unix.go; Name is set by the caller."""
    m = st.st_mode
    h = Header(mode=stat.S_IMODE(m) & 0o777, mtime_ns=st.st_mtime_ns)
    if stat.S_ISREG(m):
        h.typeflag, h.size = TYPE_REG, st.st_size
    elif stat.S_ISDIR(m):
        h.typeflag = TYPE_DIR
    elif stat.S_ISLNK(m):
        h.typeflag, h.linkname = TYPE_SYMLINK, link
    elif stat.S_ISCHR(m):
        h.typeflag = TYPE_CHAR
    elif stat.S_ISBLK(m):
        h.typeflag = TYPE_BLOCK
    elif stat.S_ISFIFO(m):
        h.typeflag = TYPE_FIFO
    else:
        raise ValueError("archive/tar: sockets not supported")
    if m & stat.S_ISUID:
        h.mode |= C_ISUID
    if m & stat.S_ISGID:
        h.mode |= C_ISGID
    if m & stat.S_ISVTX:
        h.mode |= C_ISVTX
    h.uid, h.gid = st.st_uid, st.st_gid
    if h.typeflag in (TYPE_CHAR, TYPE_BLOCK):
        h.devmajor, h.devminor = os.major(st.st_rdev), os.minor(st.st_rdev)
    return h


def _is_ascii(s: bytes) -> bool:
    return all(c < 0x80 for c in s)


def _to_ascii(s: bytes) -> bytes:
    return bytes(c for c in s if c < 0x80)


def split_ustar_path(name: bytes) -> Optional[Tuple[bytes, bytes]]:
    length = len(name)
    if length <= 100 or not _is_ascii(name):
        return None
    if length > 155 + 1:
        length = 155 + 1
    elif name[length - 1:length] == b"/":
        length -= 1
    i = name[:length].rfind(b"/")
    nlen = len(name) - i - 1
    plen = i
    if i <= 0 or nlen > 100 or nlen == 0 or plen > 155:
        return None
    return name[:i], name[i + 1:]


def _fmt_string(block: bytearray, off: int, size: int, s: bytes) -> None:
    n = min(len(s), size)
    block[off:off + n] = s[:n]
    if len(s) < size:
        block[off + len(s)] = 0
    if len(s) > size and block[off + size - 1] == 0x2F:
        k = len(s[:size].rstrip(b"/"))
        block[off + k] = 0


def _fits_octal(n: int, x: int) -> bool:
    return 0 <= x < (1 << ((n - 1) * 3))


def _fmt_octal(block: bytearray, off: int, size: int, x: int) -> None:
    if not _fits_octal(size, x):
        x = 0
    s = ("%o" % x).encode()
    pad = size - len(s) - 1
    if pad > 0:
        s = b"0" * pad + s
    _fmt_string(block, off, size, s)


def _finish_block(block: bytearray) -> bytes:
    block[257:263] = b"ustar\x00"
    block[263:265] = b"00"
    block[148:156] = b" " * 8
    chk = sum(block)
    _fmt_octal(block, 148, 7, chk)
    block[155] = 0x20
    return bytes(block)


def _template_v7plus(h: Header, name: bytes, linkname: bytes, ascii_only: bool) -> bytearray:
    b = bytearray(512)
    f = (lambda s: _to_ascii(s)) if ascii_only else (lambda s: s)
    b[156:157] = h.typeflag
    _fmt_string(b, 0, 100, f(name))
    _fmt_string(b, 157, 100, f(linkname))
    _fmt_octal(b, 100, 8, h.mode)
    _fmt_octal(b, 108, 8, h.uid)
    _fmt_octal(b, 116, 8, h.gid)
    _fmt_octal(b, 124, 12, h.size)
    _fmt_octal(b, 136, 12, h.mtime_ns // 10**9)
    _fmt_string(b, 265, 32, f(os.fsencode(h.uname)))
    _fmt_string(b, 297, 32, f(os.fsencode(h.gname)))
    _fmt_octal(b, 329, 8, h.devmajor)
    _fmt_octal(b, 337, 8, h.devminor)
    return b


def _pax_record(k: str, v: bytes) -> bytes:
    size = len(k) + len(v) + 3
    size += len(str(size))
    rec = str(size).encode() + b" " + k.encode() + b"=" + v + b"\n"
    if len(rec) != size:
        size = len(rec)
        rec = str(size).encode() + b" " + k.encode() + b"=" + v + b"\n"
    return rec


def encode_header(h: Header) -> bytes:
    """Writer.WriteHeader for Format == FormatUnknown: returns the header block(s) (USTAR, or PAX
    extended header + data + main header).  mtime is rounded to the second (no-op after makisu's
    truncation, write.go:61)."""
    name, linkname = os.fsencode(h.name), os.fsencode(h.linkname)
    h = replace(h, mtime_ns=((h.mtime_ns + 5 * 10**8) // 10**9) * 10**9)  # ModTime.Round(time.Second)
    pax: Dict[str, bytes] = {}
    ustar_ok = True

    def verify_string(s: bytes, size: int, key: Optional[str]):
        nonlocal ustar_ok
        too_long = len(s) > size
        if not _is_ascii(s) or too_long:
            if not (key == "path" and split_ustar_path(s) is not None):
                ustar_ok = False
            if key is None:
                raise ValueError("archive/tar: header field cannot be encoded")
            pax[key] = s

    def verify_numeric(n: int, size: int, key: Optional[str]):
        nonlocal ustar_ok
        if not _fits_octal(size, n):
            ustar_ok = False
            if key is None:
                raise ValueError("archive/tar: header field too long")
            pax[key] = str(n).encode()

    verify_string(name, 100, "path")
    verify_string(linkname, 100, "linkpath")
    verify_string(os.fsencode(h.uname), 32, "uname")
    verify_string(os.fsencode(h.gname), 32, "gname")
    verify_numeric(h.mode, 8, None)
    verify_numeric(h.uid, 8, "uid")
    verify_numeric(h.gid, 8, "gid")
    verify_numeric(h.size, 12, "size")
    verify_numeric(h.devmajor, 8, None)
    verify_numeric(h.devminor, 8, None)
    verify_numeric(h.mtime_ns // 10**9, 12, "mtime")
    if h.typeflag in (TYPE_REG, TYPE_CHAR, TYPE_BLOCK, TYPE_FIFO) and h.name.endswith("/"):
        raise ValueError("archive/tar: filename may not have trailing slash")

    if ustar_ok:
        prefix = b""
        sp = split_ustar_path(name)
        if sp is not None:
            prefix, name = sp
        b = _template_v7plus(h, name, linkname, ascii_only=False)
        _fmt_string(b, 345, 155, prefix)
        return _finish_block(b)

    out = b""
    if pax:
        data = b"".join(_pax_record(k, pax[k]) for k in sorted(pax))
        d, f = posixpath.split(h.name)
        xname = _to_ascii(os.fsencode(posixpath.join(d, "PaxHeaders.0", f)))[:100].rstrip(b"/")
        xb = bytearray(512)
        xb[156:157] = TYPE_XHEADER
        _fmt_string(xb, 0, 100, xname)
        _fmt_octal(xb, 100, 8, 0)
        _fmt_octal(xb, 108, 8, 0)
        _fmt_octal(xb, 116, 8, 0)
        _fmt_octal(xb, 124, 12, len(data))
        _fmt_octal(xb, 136, 12, 0)
        out += _finish_block(xb) + data + b"\0" * (-len(data) % 512)
    b = _template_v7plus(h, name, linkname, ascii_only=True)
    return out + _finish_block(b)


def is_header_only(typeflag: bytes) -> bool:
    return typeflag in (TYPE_LINK, TYPE_SYMLINK, TYPE_CHAR, TYPE_BLOCK, TYPE_DIR, TYPE_FIFO)


TRAILER = b"\0" * 1024  # tar.Writer.Close(): exactly two zero blocks



# ---- go1.14 archive/tar Reader (stdlib, not vendored), restated ----------------------
# reader.go: Reader.next / readHeader / parsePAX / mergePAX, strconv.go: parseNumeric / parseOctal / parsePAXTime,
# format.go: getFormat.  Sparse files (GNU 'S', PAX GNU.sparse.*) are rejected: docker layers do not carry them.
class TarHeaderError(ValueError):
    """tar.ErrHeader"""


def _c_string(b: bytes) -> bytes:
    i = b.find(b"\0")
    return b if i < 0 else b[:i]


def _parse_octal(b: bytes) -> int:
    b = b.strip(b" \0")
    if not b:
        return 0
    b = _c_string(b)
    if not b or any(c not in b"01234567" for c in b):
        raise TarHeaderError("archive/tar: invalid tar header")
    return int(b, 8)


def _parse_numeric(b: bytes) -> int:
    if b and b[0] & 0x80:  # base-256 (binary), two's complement when 0x40 is set
        inv = 0xFF if b[0] & 0x40 else 0x00
        x = 0
        for i, c in enumerate(b):
            c ^= inv
            if i == 0:
                c &= 0x7F
            if x >> 56:
                raise TarHeaderError("archive/tar: invalid tar header")
            x = (x << 8) | c
        if x >> 63:
            raise TarHeaderError("archive/tar: invalid tar header")
        return ~x if inv else x
    return _parse_octal(b)


def _parse_pax_time(s: str) -> int:
    """seconds[.fraction] -> integer nanoseconds (strconv.go parsePAXTime: fraction truncated to 9 digits)."""
    ss, _, sn = s.partition(".")
    try:
        secs = int(ss, 10)
    except ValueError:
        raise TarHeaderError("archive/tar: invalid tar header")
    if not sn:
        return secs * 10**9
    if not sn.isdigit():
        raise TarHeaderError("archive/tar: invalid tar header")
    nsecs = int((sn + "0" * 9)[:9])
    return secs * 10**9 - nsecs if ss.startswith("-") else secs * 10**9 + nsecs


def _parse_pax_records(body: bytes) -> Dict[str, str]:
    out: Dict[str, str] = {}
    while body:
        sp = body.find(b" ")
        if sp < 0 or not body[:sp].isdigit():
            raise TarHeaderError("archive/tar: invalid tar header")
        n = int(body[:sp])
        if n < 5 or n > len(body):
            raise TarHeaderError("archive/tar: invalid tar header")
        rec, body = body[sp + 1:n], body[n:]
        if not rec.endswith(b"\n") or b"=" not in rec:
            raise TarHeaderError("archive/tar: invalid tar header")
        k, _, v = rec[:-1].partition(b"=")
        key, val = k.decode("utf-8", "surrogateescape"), v.decode("utf-8", "surrogateescape")
        if key.startswith("GNU.sparse."):
            raise TarHeaderError("archive/tar: sparse entries are not supported by this restatement")
        if val:
            out[key] = val
        else:
            out.pop(key, None)
    return out


@dataclass
class TarMember:
    hdr: Header
    data_off: int   # offset of the member's data in the stream
    data_len: int   # bytes of data the Reader exposes (0 for header-only types)


def read_tar(data: bytes) -> List[TarMember]:
    """tar.Reader.Next() until io.EOF over an in-memory stream."""
    members: List[TarMember] = []
    pos = 0
    pax: Dict[str, str] = {}
    gnu_name = gnu_link = ""
    while True:
        blk = data[pos:pos + 512]
        if len(blk) == 0:
            return members                      # io.EOF at a block boundary
        if len(blk) < 512:
            raise TarHeaderError("unexpected EOF")
        if blk == b"\0" * 512:
            nxt = data[pos + 512:pos + 1024]
            if len(nxt) == 0 or nxt == b"\0" * 512:
                return members                  # one zero block at EOF, or two: end of archive
            if len(nxt) < 512:
                raise TarHeaderError("unexpected EOF")
            raise TarHeaderError("archive/tar: invalid tar header")
        pos += 512
        # checksum: unsigned or signed sum with the field read as spaces
        want = _parse_octal(blk[148:156])
        b2 = blk[:148] + b" " * 8 + blk[156:]
        unsigned = sum(b2)
        signed = sum(c - 256 if c > 127 else c for c in b2)
        if want != unsigned and want != signed:
            raise TarHeaderError("archive/tar: invalid tar header")
        magic, version = blk[257:263], blk[263:265]
        if magic == b"ustar\0" and blk[508:512] == b"tar\0":
            fmt = "star"
        elif magic == b"ustar\0":
            fmt = "ustar"
        elif magic == b"ustar " and version == b" \0":
            fmt = "gnu"
        else:
            fmt = "v7"
        h = Header(name=_c_string(blk[0:100]).decode("utf-8", "surrogateescape"), mode=_parse_numeric(blk[100:108]),
                   uid=_parse_numeric(blk[108:116]), gid=_parse_numeric(blk[116:124]), size=_parse_numeric(blk[124:136]),
                   mtime_ns=_parse_numeric(blk[136:148]) * 10**9, typeflag=blk[156:157],
                   linkname=_c_string(blk[157:257]).decode("utf-8", "surrogateescape"))
        if fmt != "v7":
            h.uname = _c_string(blk[265:297]).decode("utf-8", "surrogateescape")
            h.gname = _c_string(blk[297:329]).decode("utf-8", "surrogateescape")
            h.devmajor, h.devminor = _parse_numeric(blk[329:337]), _parse_numeric(blk[337:345])
            prefix = b""
            if fmt == "ustar":
                prefix = _c_string(blk[345:500])
            elif fmt == "star":
                prefix = _c_string(blk[345:476])
            if prefix:
                h.name = prefix.decode("utf-8", "surrogateescape") + "/" + h.name
        nb = 0 if is_header_only(h.typeflag) else h.size
        if nb < 0:
            raise TarHeaderError("archive/tar: invalid tar header")
        if h.typeflag in (b"x", b"g"):
            body = data[pos:pos + nb]
            if len(body) < nb:
                raise TarHeaderError("unexpected EOF")
            pos += (nb + 511) // 512 * 512
            recs = _parse_pax_records(body)
            if h.typeflag == b"g":
                raise TarHeaderError("unsupported type 1100111")  # Next() returns it; IsSimilarHeader then rejects 'g'
            pax = recs
            continue
        if h.typeflag in (b"L", b"K"):
            body = data[pos:pos + nb]
            if len(body) < nb:
                raise TarHeaderError("unexpected EOF")
            pos += (nb + 511) // 512 * 512
            if h.typeflag == b"L":
                gnu_name = _c_string(body).decode("utf-8", "surrogateescape")
            else:
                gnu_link = _c_string(body).decode("utf-8", "surrogateescape")
            continue
        if h.typeflag == b"S":
            raise TarHeaderError("archive/tar: sparse entries are not supported by this restatement")
        for k, v in pax.items():  # mergePAX
            try:
                if k == "path":
                    h.name = v
                elif k == "linkpath":
                    h.linkname = v
                elif k == "uname":
                    h.uname = v
                elif k == "gname":
                    h.gname = v
                elif k == "uid":
                    h.uid = int(v, 10)
                elif k == "gid":
                    h.gid = int(v, 10)
                elif k == "mtime":
                    h.mtime_ns = _parse_pax_time(v)
                elif k == "size":
                    h.size = int(v, 10)
            except ValueError:
                raise TarHeaderError("archive/tar: invalid tar header")
        if gnu_name:
            h.name = gnu_name
        if gnu_link:
            h.linkname = gnu_link
        if h.typeflag == b"\0":  # TypeRegA
            h.typeflag = TYPE_DIR if h.name.endswith("/") else TYPE_REG
        nb = 0 if is_header_only(h.typeflag) else h.size
        if nb < 0:
            raise TarHeaderError("archive/tar: invalid tar header")
        if pos + nb > len(data):
            raise TarHeaderError("unexpected EOF")
        members.append(TarMember(h, pos, nb))
        pos += (nb + 511) // 512 * 512
        pax, gnu_name, gnu_link = {}, "", ""

# ---- tario.IsSimilarHeader (lib/tario/compare.go) -----------------------------------
def is_similar_header(h: Header, nh: Header, ignore_time: bool = False) -> bool:
    if h.name == "" and nh.name == "":
        return True
    time_eq = ignore_time or h.mtime_ns // 10**9 == nh.mtime_ns // 10**9
    if h.typeflag == TYPE_SYMLINK:
        return nh.typeflag == TYPE_SYMLINK and h.linkname == nh.linkname
    if h.typeflag == TYPE_LINK:
        return (nh.typeflag == TYPE_LINK and time_eq and h.linkname == nh.linkname and h.uid == nh.uid
                and h.gid == nh.gid and h.file_mode_bits() == nh.file_mode_bits())
    if h.typeflag == TYPE_DIR:
        return (nh.typeflag == TYPE_DIR and time_eq and h.uid == nh.uid and h.gid == nh.gid
                and h.file_mode_bits() == nh.file_mode_bits())
    if h.typeflag == TYPE_REG:
        return (nh.typeflag == TYPE_REG and time_eq and h.uid == nh.uid and h.gid == nh.gid and h.size == nh.size
                and h.file_mode_bits() == nh.file_mode_bits())
    raise ValueError("unsupported type %r" % h.typeflag)


# ---- MemFS ---------------------------------------------------------------------------
@dataclass
class MemFile:
    src: str
    dst: str
    hdr: Header
    whiteout: bool = False
    deleted: str = ""
    digest: Optional[bytes] = None  # SHA-256 of the content the tree believes the file has (content-aware scan; ours)


@dataclass
class Node:
    mf: MemFile
    children: Dict[str, "Node"] = field(default_factory=dict)


@dataclass
class CopyOperation:
    """lib/snapshot/copy_op.go:29-80"""
    src_root: str
    srcs: List[str]
    dst: str
    uid: int = 0
    gid: int = 0

    @staticmethod
    def new(srcs: List[str], src_root: str, work_dir: str, dst: str, uid: int = 0, gid: int = 0) -> "CopyOperation":
        if not srcs:
            raise ValueError("srcs cannot be empty")
        is_dir_fmt = dst.endswith("/") or dst in (".", "..")
        if len(srcs) > 1 and not is_dir_fmt:
            raise ValueError('tarring multiple sources, destination must end with "/"')
        if not posixpath.isabs(dst):
            if not posixpath.isabs(work_dir):
                raise ValueError("dst is not absolute path, must specify absolute working directory")
            d = posixpath.normpath(posixpath.join(work_dir, dst))
            dst = d + "/" if is_dir_fmt else d
        return CopyOperation(src_root, [rel_path(s) for s in srcs], dst, uid, gid)


class MemFS:
    def __init__(self, now: Callable[[], float], root: str, blacklist: Optional[List[str]] = None,
                 is_mountpoint: Callable[[str], bool] = lambda p: False):
        self.now = now
        self.root = root
        self.blacklist = blacklist or []
        self.is_mountpoint = is_mountpoint
        st = os.lstat(root)
        hdr = self.create_header(root, "/", st)
        self.tree = Node(MemFile(root, "/", hdr))
        self.layers: List[Dict[str, MemFile]] = []

    # mem_layer.go:152-190
    def create_header(self, src: str, dst: str, st: os.stat_result, from_header: Optional[Header] = None) -> Header:
        if from_header is not None:
            # tar.FileInfoHeader(hdr.FileInfo()): type + perm bits + owner of the ancestor header
            hdr = Header(mode=from_header.mode & 0o7777, uid=from_header.uid, gid=from_header.gid,
                         typeflag=from_header.typeflag, mtime_ns=from_header.mtime_ns,
                         size=from_header.size if from_header.typeflag == TYPE_REG else 0)
        else:
            link = os.readlink(src) if stat.S_ISLNK(st.st_mode) else ""
            hdr = file_info_header(st, link)
        hdr.name = rel_path(dst)
        hdr.uname = hdr.gname = ""
        asrc = abs_path(src)
        if hdr.typeflag == TYPE_DIR:
            if not asrc.endswith("/"):
                hdr.name += "/"
        elif hdr.typeflag == TYPE_SYMLINK and from_header is None:
            target = os.readlink(asrc)
            if posixpath.isabs(target):
                target = trim_root(target, self.root)
            hdr.linkname = target
        return hdr

    # utils.go:37-75
    def _should_skip(self, p: str, st: os.stat_result, blacklist: List[str]) -> bool:
        if posixpath.basename(p).startswith(WHITEOUT_META_PREFIX):
            return True
        if is_descendant_of_any(p, blacklist) or is_special_file(st):
            return True
        return self.is_mountpoint(p)

    def _walk(self, src_root: str, blacklist: List[str], f: Callable[[str, os.stat_result], None]) -> None:
        def visit(p: str, st: os.stat_result):
            if self._should_skip(p, st, blacklist):
                return "skipdir" if stat.S_ISDIR(st.st_mode) else None
            f(p, st)
            return None
        go_walk(src_root, visit)

    # mem_layer.go:192-211 + updateMemFS
    def _add_header(self, layer: Dict[str, MemFile], src: str, dst: str, hdr: Header) -> None:
        src, dst = abs_path(src), abs_path(dst)
        d, b = posixpath.split(dst)
        if b.startswith(WHITEOUT_PREFIX):
            deleted = posixpath.join(d, b[len(WHITEOUT_PREFIX):])
            mf = MemFile("", dst, Header(name=rel_path(dst)), whiteout=True, deleted=deleted)
            layer[deleted] = mf
            self._tree_delete(deleted)
            return
        mf = MemFile(src, dst, hdr)
        layer[dst] = mf
        self._tree_put(mf)

    def _tree_put(self, mf: MemFile) -> None:
        node = self.tree
        parts = split_path(mf.dst)
        for i, part in enumerate(parts):
            last = i == len(parts) - 1
            if part in node.children:
                if last:
                    old = node.children[part]
                    nn = Node(mf)
                    if mf.hdr.typeflag == TYPE_DIR:
                        nn.children.update(old.children)
                    node.children[part] = nn
                else:
                    node = node.children[part]
            else:
                if last:
                    node.children[part] = Node(mf)
                else:
                    raise ValueError(f"missing intermediate directory {part} in {mf.dst}")

    def _tree_delete(self, path: str) -> None:
        node = self.tree
        parts = split_path(path)
        for i, part in enumerate(parts):
            if part in node.children:
                if i == len(parts) - 1:
                    del node.children[part]
                else:
                    node = node.children[part]
            elif i != len(parts) - 1:
                raise ValueError(f"missing intermediate dir {part} in {path}")

    # mem_fs.go:487-503
    def _is_updated(self, p: str, hdr: Header) -> Tuple[bool, Optional[Node]]:
        cur = self.tree
        for part in split_path(p):
            if part in cur.children:
                cur = cur.children[part]
            else:
                return True, None
        return not is_similar_header(cur.mf.hdr, hdr, False), cur

    # mem_fs.go:509-569
    def _add_ancestors(self, layer, dst: str, inclusive: bool, depth: int, uid: int, gid: int) -> str:
        if depth >= 1024:
            raise ValueError(f"symlink loop at {dst}")
        if depth == 0:  # the reference recurses 1024 deep before it gives up; CPython's default limit is 1000 frames
            import sys
            if sys.getrecursionlimit() < 4000:
                sys.setrecursionlimit(4000)
        last_ancestor = self.tree
        cur = self.tree
        parts = split_path(dst)
        end = len(parts) if inclusive else len(parts) - 1
        i = 0
        while i < end:
            part = parts[i]
            n = cur.children.get(part)
            if n is None:
                break
            self._add_header(layer, n.mf.src, n.mf.dst, n.mf.hdr)
            n = cur.children[part]
            if n.mf.hdr.typeflag == TYPE_DIR:
                last_ancestor = n
                cur = n
            elif n.mf.hdr.typeflag == TYPE_SYMLINK:
                remaining = posixpath.join(*parts[i + 1:]) if parts[i + 1:] else ""
                target = posixpath.join(n.mf.hdr.linkname, remaining)
                return self._add_ancestors(layer, target, inclusive, depth + 1, uid, gid)
            i += 1
        for j in range(i, end):
            cur_path = abs_path(posixpath.join(*parts[:j + 1]))
            hdr = self.create_header("", cur_path, None, from_header=last_ancestor.mf.hdr)  # type: ignore[arg-type]
            hdr.mtime_ns = int(self.now()) * 10**9  # clk.Now(); tests inject whole seconds
            hdr.uid, hdr.gid = uid, gid
            self._add_header(layer, "", cur_path, hdr)
        return dst

    # mem_fs.go:440-482
    def _maybe_add(self, layer, src: str, dst: str, hdr: Header, create_whiteout: bool) -> None:
        updated, node = self._is_updated(dst, hdr)
        if updated and dst != "/":
            self._add_ancestors(layer, abs_path(dst), False, 0, 0, 0)
            self._add_header(layer, src, dst, hdr)
        if create_whiteout and hdr.typeflag == TYPE_DIR and node is not None:
            for child in list(node.children.values()):
                on_disk = os.path.lexists(child.mf.src)
                if not on_disk:
                    d, b = posixpath.split(abs_path(child.mf.dst))
                    wpath = posixpath.join(d, WHITEOUT_PREFIX + b)
                    layer[child.mf.dst] = MemFile("", wpath, Header(name=rel_path(wpath)), whiteout=True,
                                                  deleted=child.mf.dst)
                    self._tree_delete(child.mf.dst)
                    self._add_ancestors(layer, child.mf.dst, False, 0, 0, 0)

    # mem_fs.go:353-420
    def _add_to_layer(self, layer, c: CopyOperation) -> None:
        create_dst = True
        if len(c.srcs) == 1:
            src = posixpath.normpath(posixpath.join(c.src_root, c.srcs[0])) if c.srcs[0] else c.src_root
            if not os.path.isdir(src):
                os.stat(src)
                create_dst = False
        if create_dst:
            resolved = self._add_ancestors(layer, abs_path(c.dst), True, 0, c.uid, c.gid)
            if not resolved.endswith("/"):
                resolved += "/"
            c.dst = resolved
        for s in c.srcs:
            # mem_fs.go:380-384: src, err = evalSymlinks(src, c.srcRoot); src = filepath.Join(c.srcRoot, src)
            from .copier import eval_symlinks  # local import: copier.py imports this module
            s = eval_symlinks(s, c.src_root)
            src = go_clean(c.src_root + "/" + s) if s else c.src_root

            def visit(cur_src: str, st: os.stat_result, src=src):
                if cur_src == src:
                    if stat.S_ISDIR(st.st_mode):
                        return
                    elif not c.dst.endswith("/"):
                        cur_dst = c.dst
                    else:
                        cur_dst = posixpath.normpath(posixpath.join(c.dst, posixpath.basename(src)))
                else:
                    cur_dst = posixpath.normpath(c.dst + "/" + cur_src[len(src):])
                hdr = self.create_header(cur_src, cur_dst, st)
                hdr.uid, hdr.gid = c.uid, c.gid
                self._maybe_add(layer, cur_src, cur_dst, hdr, False)

            self._walk(src, [], visit)

    def add_layer_by_copy_ops(self, ops: List[CopyOperation]) -> List[MemFile]:
        """AddLayerByCopyOps (mem_fs.go:276-289) minus sync(): returns the entries in tar order."""
        layer: Dict[str, MemFile] = {}
        for c in ops:
            self._add_to_layer(layer, c)
        self.layers.append(layer)
        return [layer[k] for k in sorted(layer, key=os.fsencode)]  # mem_layer.go:232-244 sort.Strings

    def add_layer_by_scan(self, content_aware: bool = False) -> List[MemFile]:
        """AddLayerByScan (mem_fs.go:260-270,315-341).  content_aware (ours): regular files whose
        header is "similar" but whose remembered content digest differs from the file's current SHA-256 join the
        layer as if isUpdated had said so."""
        layer: Dict[str, MemFile] = {}
        root = self.root
        suspects: List[Tuple[str, str, Header, bytes]] = []

        def visit(src: str, st: os.stat_result):
            dst = trim_root(src, root)
            hdr = self.create_header(src, dst, st)
            if content_aware and hdr.typeflag == TYPE_REG:
                updated, node = self._is_updated(dst, hdr)
                if not updated and node is not None and node.mf.digest is not None:
                    suspects.append((src, dst, hdr, node.mf.digest))
            self._maybe_add(layer, src, dst, hdr, True)

        self._walk(root, self.blacklist, visit)
        for src, dst, hdr, known in suspects:
            with open(src, "rb") as f:
                now = hashlib.sha256(f.read()).digest()
            if now != known:
                self._add_ancestors(layer, abs_path(dst), False, 0, 0, 0)
                self._add_header(layer, src, dst, hdr)
        self.layers.append(layer)
        return [layer[k] for k in sorted(layer, key=os.fsencode)]

    def _set_digest(self, dst: str, digest: bytes) -> None:
        cur = self.tree
        for part in split_path(dst):
            if part not in cur.children:
                return
            cur = cur.children[part]
        cur.mf.digest = digest

    def remember_content(self, entries: List[MemFile]) -> None:
        """MKHOST_FILE_DIGESTS: after a layer is committed, remember the SHA-256 of every regular file in it."""
        for e in entries:
            if not e.whiteout and e.hdr.typeflag == TYPE_REG and e.hdr.size:
                with open(e.src, "rb") as f:
                    self._set_digest(e.dst, hashlib.sha256(f.read()).digest())


    def update_from_tar(self, data: bytes, remember: bool = False, untar: bool = False) -> List[MemFile]:
        """UpdateFromTarReader(r, untar) (mem_fs.go:165-255): merge the headers of an (uncompressed) layer tar into the
        tree, hard links in a second pass.  untar=True also writes the members under the root (untarOneItem,
        mem_fs.go:574-716; tario.ApplyHeader, lib/tario/apply.go:26-49) and restores the parent directories' mtimes.
        Returns the merged layer in key order (the reference only logs its count)."""
        layer: Dict[str, MemFile] = {}
        hardlinks: Dict[str, Header] = {}
        modtimes: Dict[str, int] = {}
        members = read_tar(data)
        for m in members:
            hdr = m.hdr
            path = go_clean(self.root + "/" + hdr.name)  # filepath.Join(fs.tree.src, hdr.Name)
            if posixpath.basename(path).startswith(WHITEOUT_META_PREFIX):
                continue
            if is_descendant_of_any(path, self.blacklist) or hdr.is_special() or self.is_mountpoint(path):
                continue
            if untar:
                parent = posixpath.dirname(path)
                if parent not in modtimes:
                    modtimes[parent] = os.lstat(parent).st_mtime_ns      # "stat parent dir of ..." when it is missing
            hdr = replace(hdr, name=rel_path(hdr.name))
            if hdr.typeflag == TYPE_LINK:
                hdr.linkname = abs_path(hdr.linkname)
                hardlinks[path] = hdr
            else:
                if untar:
                    self._untar_one_item(path, hdr, data[m.data_off:m.data_off + m.data_len])
                self._maybe_add(layer, abs_path(hdr.name), abs_path(hdr.name), hdr, False)
        for path in sorted(hardlinks, key=os.fsencode):  # Go ranges over a map: order is unspecified, result is not affected
            hdr = hardlinks[path]
            if untar:
                self._untar_one_item(path, hdr, b"")
            self._maybe_add(layer, abs_path(hdr.name), abs_path(hdr.name), hdr, False)
        for path, ns in modtimes.items():
            os.utime(path, ns=(ns, ns))
        if remember:
            for m in members:
                dst = abs_path(m.hdr.name)
                if m.hdr.typeflag == TYPE_REG and m.data_len and dst in layer and not layer[dst].whiteout \
                        and layer[dst].hdr.typeflag == TYPE_REG:
                    self._set_digest(dst, hashlib.sha256(data[m.data_off:m.data_off + m.data_len]).digest())
        self.layers.append(layer)
        return [layer[k] for k in sorted(layer, key=os.fsencode)]

    # lib/tario/apply.go:26-49
    @staticmethod
    def _apply_header(path: str, hdr: Header) -> None:
        st = os.lstat(path)
        if stat.S_ISLNK(st.st_mode) or hdr.file_mode_bits() & GO_MODE_SYMLINK:
            raise OSError("update symlink instead of file: %s" % path)
        os.chown(path, hdr.uid, hdr.gid)
        os.chmod(path, go_chmod_bits(hdr.file_mode_bits()))   # after chown: setuid/setgid survive
        os.utime(path, ns=(hdr.mtime_ns, hdr.mtime_ns))

    # mem_fs.go:574-716
    def _untar_one_item(self, path: str, hdr: Header, body: bytes) -> None:
        base = posixpath.basename(path)
        if base.startswith(WHITEOUT_PREFIX):
            victim = posixpath.join(posixpath.dirname(path), base[len(WHITEOUT_PREFIX):])
            if os.path.lexists(victim):
                if os.path.isdir(victim) and not os.path.islink(victim):
                    import shutil
                    shutil.rmtree(victim)
                else:
                    os.remove(victim)
            return
        if os.path.lexists(path):
            st = os.lstat(path)
            link = ""
            if stat.S_ISLNK(st.st_mode):
                link = os.readlink(path)
                if posixpath.isabs(link):
                    link = trim_root(link, self.root)
            local = file_info_header(st, link)
            local.name = base
            if is_similar_header(local, hdr, False):
                return
            if hdr.file_mode_bits() & GO_MODE_DIR and stat.S_ISDIR(st.st_mode):
                self._apply_header(path, hdr)
                return
            if stat.S_ISDIR(st.st_mode):
                import shutil
                shutil.rmtree(path)
            else:
                os.remove(path)
        if hdr.typeflag == TYPE_DIR:
            os.mkdir(path, go_chmod_bits(hdr.file_mode_bits()))
            self._apply_header(path, hdr)
        elif hdr.typeflag == TYPE_SYMLINK:
            target = hdr.linkname
            if posixpath.isabs(target):
                target = go_clean(self.root + "/" + target)
            os.symlink(target, path)
            os.lchown(path, hdr.uid, hdr.gid)
        elif hdr.typeflag == TYPE_LINK:
            os.link(go_clean(self.root + "/" + hdr.linkname), path)
            self._apply_header(path, hdr)
        else:
            fd = os.open(path, os.O_CREAT | os.O_TRUNC | os.O_WRONLY, go_chmod_bits(hdr.file_mode_bits()))
            try:
                os.write(fd, body)
            finally:
                os.close(fd)
            self._apply_header(path, hdr)


# ---- tario.WriteEntry / WriteHeader (lib/tario/write.go) -----------------------------
def entry_header_bytes(mf: MemFile) -> bytes:
    h = replace(mf.hdr, name=mf.hdr.name.lstrip("/"))
    h.mtime_ns = (h.mtime_ns // 10**9) * 10**9  # ModTime.Truncate(1s)
    return encode_header(h)


def layer_tar_chunks(entries: List[MemFile]) -> Iterator[bytes]:
    """The exact byte stream tar.Writer emits for the committed layer."""
    for mf in entries:
        yield entry_header_bytes(mf)
        if not mf.whiteout and mf.hdr.typeflag == TYPE_REG and mf.hdr.size:
            left = mf.hdr.size
            with open(mf.src, "rb") as f:
                while left:
                    buf = f.read(min(left, 1 << 20))
                    if not buf:
                        raise IOError(f"copy file {mf.src} to tar writer: unexpected EOF")  # io.CopyN
                    left -= len(buf)
                    yield buf
            yield b"\0" * (-mf.hdr.size % 512)
    yield TRAILER


def tar_digest(entries: List[MemFile]) -> str:
    h = hashlib.sha256()
    for c in layer_tar_chunks(entries):
        h.update(c)
    return "sha256:" + h.hexdigest()
