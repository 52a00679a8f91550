"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle, bit exact.

Reads like the reference's own tests where they exist:
  reference lib/builder/step/copy_step_test.go:51-169 (TestCopyStepSetCacheID): same content => same
  ID, changed content / order => different ID -- here additionally checked against the exact CRC value.
"""
import hashlib
import zlib

import numpy as np
import pytest

from tests.util import pack, crc_extents, cdc_extents, ranges

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from makisu_b200.abi import Engine
    e = Engine(device=0, device_arena_bytes=256 << 20, n_host_arenas=2, host_arena_bytes=64 << 20,
               max_extents=1 << 16)
    yield e
    e.close()


def _rand(rng, n):
    return rng.integers(0, 256, n, dtype=np.uint8)


SIZES = [0, 1, 3, 15, 16, 17, 63, 64, 65, 511, 512, 513, 4095, 4096, 4097, 70000, 262143, 262144, 262145,
         524288, 1000003, 3 * 262144 + 77]


def test_crc32_ragged_extents(eng):
    rng = np.random.default_rng(7)
    segs = [_rand(rng, n) for n in SIZES]
    arena, offs = pack(segs, file_align=16)
    lens = [len(s) for s in segs]
    order = list(rng.permutation(len(segs)))
    ext, total = crc_extents(offs, lens, order)
    eng.begin()
    eng.device_upload(0, 0, arena)
    eng.device_submit(0, arena.size, ext)
    res = eng.finish()
    want = zlib.crc32(b"".join(segs[i].tobytes() for i in order))
    assert res.crc_bytes == total
    assert eng.ctx_crc32(res) == want
    # split over two submits (session accumulates, suffixes are global)
    half = len(ext) // 2
    eng.begin()
    eng.device_submit(0, arena.size, ext[:half])
    eng.device_submit(0, arena.size, ext[half:])
    res2 = eng.finish()
    assert eng.ctx_crc32(res2) == want
    # relational properties of TestCopyStepSetCacheID: changed content => different id
    arena2 = arena.copy()
    arena2[offs[-1] + 5] ^= 1
    eng.begin()
    eng.device_upload(0, 0, arena2)
    eng.device_submit(0, arena2.size, ext)
    assert eng.ctx_crc32(eng.finish()) != want


def test_crc32_empty_stream(eng):
    eng.begin()
    res = eng.finish()
    assert eng.ctx_crc32(res) == zlib.crc32(b"") == 0
    assert res.n_chunks == 0 and res.n_unique == 0
    assert bytes(res.root) == hashlib.sha256(b"").digest()


def test_sha256_streams(eng):
    rng = np.random.default_rng(11)
    sizes = [0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 121, 128, 1000, 1024, 65536 + 3, 1 << 20]
    segs = [_rand(rng, n) for n in sizes]
    arena, offs = pack(segs, file_align=16)
    eng.begin()
    eng.device_upload(0, 0, arena)
    eng.device_submit(0, arena.size, [], ranges(offs, sizes))
    res = eng.finish()
    got = eng.get_stream_digests(res.n_streams)
    assert res.n_streams == len(sizes)
    for s, g in zip(segs, got):
        assert g.tobytes() == hashlib.sha256(s.tobytes()).digest()


def _check_table(eng, oracle_lib, arena, offs, lens, via_host=False):
    eng.begin()
    ext = cdc_extents(offs, lens)
    if via_host:
        ptr, cap, aid = eng.arena_acquire()
        import ctypes
        ctypes.memmove(ptr, arena.ctypes.data, arena.size)
        eng.arena_submit(aid, arena.size, ext)
    else:
        eng.device_upload(0, 0, arena)
        eng.device_submit(0, arena.size, ext)
    res = eng.finish()
    want = oracle_lib.chunk_table(arena, offs, lens)
    assert res.n_chunks == want["n_chunks"]
    ends, dig = eng.get_chunks(res.n_chunks)
    np.testing.assert_array_equal(ends, want["ends"])
    np.testing.assert_array_equal(dig, want["digests"])
    assert res.n_unique == want["n_unique"]
    np.testing.assert_array_equal(eng.get_table(res.n_unique), want["table"])
    assert bytes(res.root) == want["root"]
    return res


def test_chunk_table_random_files(eng, oracle_lib):
    rng = np.random.default_rng(3)
    lens = [0, 10, 4095, 4096, 4097, 8192, 40000, 131072, 131073, 300000, 1 << 20, 2500000, 31, 77]
    segs = [_rand(rng, n) for n in lens]
    arena, offs = pack(segs)
    _check_table(eng, oracle_lib, arena, offs, lens)
    _check_table(eng, oracle_lib, arena, offs, lens, via_host=True)


def test_chunk_table_low_entropy_and_duplicates(eng, oracle_lib):
    rng = np.random.default_rng(5)
    zeros = np.zeros(700000, dtype=np.uint8)                       # no candidates: forced max cuts
    ones = np.full(300001, 0xFF, dtype=np.uint8)
    text = np.frombuffer((b"the quick brown fox jumps over the lazy dog\n" * 20000), dtype=np.uint8)
    period = np.tile(_rand(rng, 37), 20000)                        # short period
    blob = _rand(rng, 600000)
    dup = np.concatenate([blob, blob, blob[:250000]])               # repeated content resynchronises
    segs = [zeros, ones, text, period, blob, dup, blob.copy()]
    lens = [len(s) for s in segs]
    arena, offs = pack(segs)
    res = _check_table(eng, oracle_lib, arena, offs, lens)
    assert res.n_unique < res.n_chunks  # dedup happened


def test_chunk_table_many_small_files(eng, oracle_lib):
    rng = np.random.default_rng(9)
    lens = [int(x) for x in rng.integers(0, 20000, 3000)]
    segs = [_rand(rng, n) for n in lens]
    arena, offs = pack(segs)
    _check_table(eng, oracle_lib, arena, offs, lens)


def test_everything_in_one_submit(eng, oracle_lib):
    """CRC + CDC + a serial stream over the same arena, like a COPY step that is both fingerprinted and committed."""
    rng = np.random.default_rng(21)
    lens = [123456, 7, 999999, 512, 65536]
    segs = [_rand(rng, n) for n in lens]
    arena, offs = pack(segs)
    from makisu_b200.abi import MKSNAP_X_CDC
    ext, total = crc_extents(offs, lens, [4, 0, 1, 2, 3], flags_extra=[MKSNAP_X_CDC] * 5)
    eng.begin()
    eng.device_upload(0, 0, arena)
    eng.device_submit(0, arena.size, ext, ranges([0], [arena.size]))
    res = eng.finish()
    assert eng.ctx_crc32(res) == zlib.crc32(b"".join(segs[i].tobytes() for i in [4, 0, 1, 2, 3]))
    # CDC extents are reported in submit order: [4,0,1,2,3]
    want = oracle_lib.chunk_table(arena, [offs[i] for i in [4, 0, 1, 2, 3]], [lens[i] for i in [4, 0, 1, 2, 3]])
    assert bytes(res.root) == want["root"] and res.n_chunks == want["n_chunks"]
    assert eng.get_stream_digests(1)[0].tobytes() == hashlib.sha256(arena.tobytes()).digest()


def test_synth_fill_matches_oracle(eng, oracle_lib):
    eng.begin()
    eng.synth_fill(0, 4096, 1 << 20, 0xC2)
    got = eng.device_download(0, 4096, 1 << 20)
    np.testing.assert_array_equal(got, oracle_lib.synth_fill(4096, 1 << 20, 0xC2))
    eng.finish()


def test_sha256_stream_continues_across_submits(eng):
    """One serial stream fed in three pieces over three submits (MKSNAP_R_MORE), next to an unrelated one-shot
    stream: the midstate survives between submits (a layer tar larger than one arena)."""
    from makisu_b200.abi import Range, MKSNAP_R_MORE
    rng = np.random.default_rng(31)
    data = _rand(rng, 3 * 65536 + 4096 + 77)
    other = _rand(rng, 1000)
    cuts = [0, 65536, 65536 + 128 * 1024, data.size]
    eng.begin()
    for i in range(3):
        piece = data[cuts[i]:cuts[i + 1]]
        arena, offs = pack([piece, other])
        eng.device_upload(0, 0, arena)
        r = Range()
        r.arena_off, r.len, r.stream, r.flags = offs[0], piece.size, 5, (MKSNAP_R_MORE if i < 2 else 0)
        rs = [r]
        if i == 1:
            r2 = Range()
            r2.arena_off, r2.len, r2.stream, r2.flags = offs[1], other.size, 2, 0
            rs.append(r2)
        eng.device_submit(0, arena.size, [], rs)
    res = eng.finish()
    assert res.n_streams == 6
    got = eng.get_stream_digests(6)
    assert got[5].tobytes() == hashlib.sha256(data.tobytes()).digest()
    assert got[2].tobytes() == hashlib.sha256(other.tobytes()).digest()
    # misuse is rejected loudly
    from makisu_b200.abi import MksnapError
    eng.begin()
    bad = Range()
    bad.arena_off, bad.len, bad.stream, bad.flags = 0, 100, 0, MKSNAP_R_MORE
    with pytest.raises(MksnapError):
        eng.device_submit(0, 4096, [], [bad])
    eng.finish()


def test_chunk_order_knob_is_result_neutral(oracle_lib, monkeypatch):
    """K2 hands chunks out longest length class first (k_len_order); the digests are stored by chunk index, so the
    chunk list, table and root must not depend on it -- also when a session spans several submits (order indices
    are batch relative)."""
    from makisu_b200.abi import Engine
    rng = np.random.default_rng(33)
    lens = [int(x) for x in rng.integers(0, 400000, 60)] + [0, 1, 4096, 131072, 131073]
    segs = [_rand(rng, n) for n in lens]
    arena, offs = pack(segs)
    want = oracle_lib.chunk_table(arena, offs, lens)
    got = []
    for knob in ("1", "0"):
        monkeypatch.setenv("MKSNAP_SHA_ORDER", knob)
        with Engine(device=0, device_arena_bytes=64 << 20, max_extents=1 << 12) as e:
            ext = cdc_extents(offs, lens)
            e.begin()
            e.device_upload(0, 0, arena)
            half = len(ext) // 2
            e.device_submit(0, arena.size, ext[:half])
            e.device_submit(0, arena.size, ext[half:])
            res = e.finish()
            ends, dig = e.get_chunks(res.n_chunks)
            got.append((bytes(res.root), res.n_chunks, res.n_unique, ends.copy(), dig.copy()))
    # chunk ends are reported in session-stream coordinates: the second submit starts `arena.size` later
    boundary = int(ext[half - 1].arena_off) + int(ext[half - 1].len)
    want_ends = want["ends"] + (want["ends"] > boundary).astype(np.uint64) * np.uint64(arena.size)
    for root, n, u, ends, dig in got:
        assert root == want["root"] and n == want["n_chunks"] and u == want["n_unique"]
        np.testing.assert_array_equal(ends, want_ends)
        np.testing.assert_array_equal(dig, want["digests"])


def _split_submit(e, files, pieces_of, carry, use_host_arenas):
    """Drive one session: `files` = list of byte arrays in order; pieces_of[i] = piece lengths of file i (sum = len).
    Every piece after the first of a file opens a new submit (MKSNAP_X_MORE / MKSNAP_X_CONT)."""
    from makisu_b200.abi import Extent, MKSNAP_X_CDC, MKSNAP_X_CONT, MKSNAP_X_MORE
    import ctypes
    cap = e.limits().host_arena_bytes if use_host_arenas else e.limits().device_arena_bytes
    arena = np.zeros(cap, dtype=np.uint8)
    ext, pos = [], 0

    def flush():
        nonlocal arena, ext, pos
        if use_host_arenas:
            ptr, c, aid = e.arena_acquire()
            ctypes.memmove(ptr, arena.ctypes.data, pos)
            e.arena_submit(aid, pos, ext)
        else:
            e.sync()
            e.device_upload(0, 0, arena[:pos])
            e.device_submit(0, pos, ext)
        arena = np.zeros(cap, dtype=np.uint8)
        ext, pos = [], 0
    e.begin()
    for data, pieces in zip(files, pieces_of):
        done = 0
        for k, n in enumerate(pieces):
            first, last = k == 0, k == len(pieces) - 1
            if not first:
                flush()
                pos = carry
            pos = (pos + 511) // 512 * 512
            assert pos + n <= cap
            arena[pos:pos + n] = data[done:done + n]
            x = Extent()
            x.arena_off, x.len, x.crc_suffix = pos, n, 0
            x.flags = MKSNAP_X_CDC | (0 if first else MKSNAP_X_CONT) | (0 if last else MKSNAP_X_MORE)
            x.reserved = 0 if last else min(len(data) - done - n, 0xFFFFFFFF)
            ext.append(x)
            pos += n
            done += n
        assert done == len(data)
    flush()
    return e.finish()


@pytest.mark.parametrize("use_host_arenas", [False, True])
def test_files_larger_than_an_arena_span_submits(oracle_lib, use_host_arenas):
    """lib/tario/write.go:45 copies a file of ANY size; here a file larger than an arena travels in pieces
    (MKSNAP_X_MORE / MKSNAP_X_CONT): the open chunk at the end of a piece is carried to the next submit, so chunk
    count, every digest, the table and the root equal the oracle's over the UNDIVIDED files -- for random content,
    zero runs (forced max-size cuts straddling the boundary), pieces shorter than min_size, three-way splits, a
    >4 MiB file (CTA-per-file selection path) and small neighbours sharing the arenas."""
    from makisu_b200.abi import Engine, MksnapError, Extent, MKSNAP_X_CDC, MKSNAP_X_MORE
    rng = np.random.default_rng(2024)
    big = _rand(rng, 9_000_000)
    zr = _rand(rng, 1_500_000)
    zr[200_000:900_000] = 0                                   # forced cuts across the split
    files = [_rand(rng, 70_000), big, _rand(rng, 5), zr, _rand(rng, 300_000), _rand(rng, 600_000), _rand(rng, 131_072 * 3)]
    pieces = [[70_000], [3_145_728, 3_145_728, 9_000_000 - 2 * 3_145_728], [5], [524_288 - 512, 512 * 3, 1_500_000 - 524_288 - 1024],
              [300_000], [1024, 600_000 - 1024], [131_072, 131_072 * 2]]
    arena, offs = pack([f for f in files])
    want = oracle_lib.chunk_table(arena, offs, [len(f) for f in files])
    kw = dict(n_host_arenas=2, host_arena_bytes=8 << 20) if use_host_arenas else {}
    with Engine(device=0, device_arena_bytes=8 << 20, max_extents=256, **kw) as e:
        carry = e.limits().carry_bytes
        assert carry >= 131072 and carry % 512 == 0
        res = _split_submit(e, files, pieces, carry, use_host_arenas)
        assert (res.n_chunks, res.n_unique, bytes(res.root)) == (want["n_chunks"], want["n_unique"], want["root"])
        assert res.n_files == len(files) and res.cdc_bytes == sum(len(f) for f in files)
        ends, dig = e.get_chunks(res.n_chunks)
        np.testing.assert_array_equal(dig, want["digests"])    # same chunks in the same order
        # contract errors: a file left open at finish; a continuation that is missing
        e.begin()
        x = Extent()
        x.arena_off, x.len, x.flags, x.reserved = 0, 4096, MKSNAP_X_CDC | MKSNAP_X_MORE, 100
        if use_host_arenas:
            ptr, c, aid = e.arena_acquire()
            e.arena_submit(aid, 8192, [x])
        else:
            e.device_submit(0, 8192, [x])
        y = Extent()
        y.arena_off, y.len, y.flags = 0, 4096, MKSNAP_X_CDC
        with pytest.raises(MksnapError):
            if use_host_arenas:
                ptr, c, aid = e.arena_acquire()
                e.arena_submit(aid, 8192, [y])
            else:
                e.device_submit(0, 8192, [y])
        with pytest.raises(MksnapError):
            e.finish()
