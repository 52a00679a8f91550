"""ORACLE (test infrastructure only): cache.Manager key / entry strings.

Restates reference lib/cache/cache_manager.go:34-35 (_cachePrefix, _cacheEmptyEntry), :239-252 (parseEntry,
createEntry).  Pinned by the reference's own test vectors in tests/golden/reference_fixtures.json where present
(cache_manager_test.go uses createEntry/parseEntry round trips only: "parity unpinned" beyond the format string).
The "_chunks" companion entry is ours (no reference counterpart).
"""
from typing import Optional, Tuple

CACHE_PREFIX = "makisu_builder_cache_"
CACHE_EMPTY_ENTRY = "MAKISU_CACHE_EMPTY"


def cache_key(cache_id: str, chunk_table: bool = False) -> str:
    return CACHE_PREFIX + cache_id + ("_chunks" if chunk_table else "")


def create_entry(tar_hex: Optional[str], gzip_hex: str = "") -> str:
    if tar_hex is None:
        return CACHE_EMPTY_ENTRY
    return "%s,%s" % (tar_hex, gzip_hex)


def parse_entry(entry: str) -> Tuple[str, str]:
    if "," not in entry:
        raise ValueError("parse redis entry: %s" % entry)
    a, b = entry.split(",", 1)
    return "sha256:" + a, "sha256:" + b


def create_chunk_entry(root: bytes, n_unique: int) -> str:
    return "%s,%d" % (root.hex(), n_unique)


def parse_chunk_entry(entry: str) -> Tuple[bytes, int]:
    h, sep, n = entry.partition(",")
    if not sep or len(h) != 64 or not n.isdigit() or any(c not in "0123456789abcdef" for c in h):
        raise ValueError("parse chunk table entry: %s" % entry)
    return bytes.fromhex(h), int(n)
