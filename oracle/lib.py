"""ctypes wrapper of oracle/mkoracle.c (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
SO = _HERE / "_build" / "libmkoracle.so"


def build(force: bool = False) -> Path:
    src = _HERE / "mkoracle.c"
    if force or not SO.exists() or SO.stat().st_mtime < max(src.stat().st_mtime, (_HERE / "mkoracle.h").stat().st_mtime):
        SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-fPIC", "-shared", "-Wall", "-Wextra", "-o", str(SO), str(src)])
    return SO


class CdcParams(C.Structure):
    _fields_ = [("min_size", C.c_uint32), ("normal_size", C.c_uint32), ("max_size", C.c_uint32),
                ("strict_bits", C.c_uint32), ("loose_bits", C.c_uint32)]


class TableSummary(C.Structure):
    _fields_ = [("n_chunks", C.c_uint64), ("n_unique", C.c_uint64), ("root", C.c_uint8 * 32)]


_L = None


def L():
    global _L
    if _L is None:
        l = C.CDLL(str(build()))
        vp, u32, u64, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
        l.mko_crc32_update.restype = u32; l.mko_crc32_update.argtypes = [u32, vp, sz]
        l.mko_crc32_pure.restype = u32; l.mko_crc32_pure.argtypes = [vp, sz]
        l.mko_crc32_xpow8n.restype = u32; l.mko_crc32_xpow8n.argtypes = [u64]
        l.mko_crc32_mulmod.restype = u32; l.mko_crc32_mulmod.argtypes = [u32, u32]
        l.mko_crc32_combine.restype = u32; l.mko_crc32_combine.argtypes = [u32, u32, u64]
        l.mko_sha256.restype = None; l.mko_sha256.argtypes = [vp, sz, vp]
        l.mko_cdc_default_params.restype = None; l.mko_cdc_default_params.argtypes = [C.POINTER(CdcParams)]
        l.mko_roll_multiplier.restype = u32; l.mko_roll_multiplier.argtypes = []
        l.mko_roll_at.restype = u32; l.mko_roll_at.argtypes = [vp, sz]
        l.mko_cdc_cuts.restype = sz; l.mko_cdc_cuts.argtypes = [vp, sz, C.POINTER(CdcParams), vp, sz]
        l.mko_sort_unique_digests.restype = sz; l.mko_sort_unique_digests.argtypes = [vp, sz]
        l.mko_merkle_root.restype = None; l.mko_merkle_root.argtypes = [vp, sz, vp]
        l.mko_chunk_table.restype = C.c_int
        l.mko_chunk_table.argtypes = [vp, vp, vp, sz, C.POINTER(CdcParams), vp, vp, vp, sz, C.POINTER(TableSummary)]
        l.mko_synth_fill.restype = None; l.mko_synth_fill.argtypes = [vp, u64, u64, u64]
        l.mko_step_same_work.restype = sz
        l.mko_step_same_work.argtypes = [vp, vp, vp, sz, C.POINTER(CdcParams), vp, vp, sz, vp]
        _L = l
    return _L


def _buf(data):
    a = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def crc32(data, crc: int = 0) -> int:
    a = _buf(data)
    return int(L().mko_crc32_update(crc, a.ctypes.data, a.size))


def crc32_pure(data) -> int:
    a = _buf(data)
    return int(L().mko_crc32_pure(a.ctypes.data, a.size))


def sha256(data) -> bytes:
    a = _buf(data)
    out = np.empty(32, dtype=np.uint8)
    L().mko_sha256(a.ctypes.data, a.size, out.ctypes.data)
    return out.tobytes()


def default_params() -> CdcParams:
    p = CdcParams()
    L().mko_cdc_default_params(C.byref(p))
    return p


def roll_multiplier() -> int:
    return int(L().mko_roll_multiplier())


def roll_at(data, i: int) -> int:
    a = _buf(data)
    return int(L().mko_roll_at(a.ctypes.data, i))


def cdc_cuts(data, params: CdcParams | None = None) -> np.ndarray:
    a = _buf(data)
    p = params or default_params()
    cap = max(16, a.size // p.min_size + 2)
    ends = np.empty(cap, dtype=np.uint64)
    n = L().mko_cdc_cuts(a.ctypes.data, a.size, C.byref(p), ends.ctypes.data, cap)
    assert n <= cap
    return ends[:n].copy()


def merkle_root(table: np.ndarray) -> bytes:
    t = np.ascontiguousarray(table, dtype=np.uint8).reshape(-1, 32)
    out = np.empty(32, dtype=np.uint8)
    L().mko_merkle_root(t.ctypes.data, t.shape[0], out.ctypes.data)
    return out.tobytes()


def sort_unique(digests: np.ndarray) -> np.ndarray:
    d = np.array(digests, dtype=np.uint8, copy=True).reshape(-1, 32)
    m = L().mko_sort_unique_digests(d.ctypes.data, d.shape[0])
    return d[:m].copy()


def chunk_table(arena, offs, lens, params: CdcParams | None = None):
    """-> dict(ends, digests, table, n_chunks, n_unique, root) for files packed in `arena`."""
    a = _buf(arena)
    o = np.ascontiguousarray(offs, dtype=np.uint64)
    ln = np.ascontiguousarray(lens, dtype=np.uint64)
    p = params or default_params()
    s = TableSummary()
    L().mko_chunk_table(a.ctypes.data, o.ctypes.data, ln.ctypes.data, o.size, C.byref(p), None, None, None, 0, C.byref(s))
    n = int(s.n_chunks)
    ends = np.empty(max(n, 1), dtype=np.uint64)
    dig = np.empty((max(n, 1), 32), dtype=np.uint8)
    tab = np.empty((max(n, 1), 32), dtype=np.uint8)
    rc = L().mko_chunk_table(a.ctypes.data, o.ctypes.data, ln.ctypes.data, o.size, C.byref(p), ends.ctypes.data,
                             dig.ctypes.data, tab.ctypes.data, max(n, 1), C.byref(s))
    assert rc == 0
    return dict(ends=ends[:n], digests=dig[:n], table=tab[: int(s.n_unique)], n_chunks=n,
                n_unique=int(s.n_unique), root=bytes(s.root))


def synth_fill(byte_off: int, n: int, seed: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    L().mko_synth_fill(out.ctypes.data, byte_off, n, seed)
    return out
