"""Shared helpers for the parity tests: pack byte strings into an arena the way the
product's host packer does (16-byte aligned extents, 512-byte aligned files)."""
from __future__ import annotations

import numpy as np

from makisu_b200.abi import Extent, Range, MKSNAP_X_CDC, MKSNAP_X_CRC


def align(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def pack(segments, file_align: int = 512):
    """segments: list of bytes/ndarray.  -> (arena ndarray, offsets list)."""
    offs, pos = [], 0
    for s in segments:
        pos = align(pos, file_align)
        offs.append(pos)
        pos += len(s)
    arena = np.zeros(align(max(pos, 16), 512), dtype=np.uint8)
    for s, o in zip(segments, offs):
        arena[o:o + len(s)] = np.frombuffer(bytes(s), dtype=np.uint8) if not isinstance(s, np.ndarray) else s
    return arena, offs


def crc_extents(offs, lens, order, flags_extra=None):
    """Extents for a CRC stream that visits segments in `order`; returns (extents list, total len)."""
    total = sum(lens[i] for i in order)
    after = total
    ext = []
    for i in order:
        after -= lens[i]
        e = Extent()
        e.arena_off, e.len, e.crc_suffix = offs[i], lens[i], after
        e.flags = MKSNAP_X_CRC | (flags_extra[i] if flags_extra else 0)
        ext.append(e)
    return ext, total


def cdc_extents(offs, lens):
    out = []
    for o, l in zip(offs, lens):
        e = Extent()
        e.arena_off, e.len, e.crc_suffix, e.flags = o, l, 0, MKSNAP_X_CDC
        out.append(e)
    return out


def ranges(offs, lens):
    out = []
    for i, (o, l) in enumerate(zip(offs, lens)):
        r = Range()
        r.arena_off, r.len, r.stream, r.flags = o, l, i, 0
        out.append(r)
    return out


def table_fingerprint(n_chunks, n_unique, root, ends, digests, crc):
    """What tests/golden/config1_1k_x_1MiB.json pins (same code for the oracle on the CPU and the engine on the GPU)."""
    import hashlib
    return {"n_chunks": int(n_chunks), "n_unique": int(n_unique), "root": bytes(root).hex(), "crc32_of_content": int(crc),
            "sha256_of_chunk_ends_u64le": hashlib.sha256(np.ascontiguousarray(ends, dtype=np.uint64).tobytes()).hexdigest(),
            "sha256_of_chunk_digests": hashlib.sha256(np.ascontiguousarray(digests, dtype=np.uint8).tobytes()).hexdigest(),
            "first_ends": [int(x) for x in ends[:4]]}
