"""-m gpu: bigger cases -- offsets beyond 4 GiB, Zipf-sized files (BASELINE configs[3]), duplicate-heavy
content (configs[4]), one long file, and size-independent properties at multi-GiB where the oracle is too slow."""
import zlib

import numpy as np
import pytest

from tests.util import cdc_extents, crc_extents

pytestmark = pytest.mark.gpu
GiB = 1 << 30


def _ext(off, ln, suffix, flags):
    from makisu_b200.abi import Extent
    e = Extent()
    e.arena_off, e.len, e.crc_suffix, e.flags = off, ln, suffix, flags
    return e


def test_offsets_beyond_4gib_match_low_offsets(oracle_lib):
    """Same 48 MiB at offset 0 and at 5 GiB + 512: identical cuts/digests, CRC of the concatenation."""
    from makisu_b200.abi import Engine, MKSNAP_X_CDC, MKSNAP_X_CRC
    n = 48 << 20
    hi = 5 * GiB + 512
    with Engine(device=0, device_arena_bytes=6 * GiB, max_extents=64) as eng:
        eng.begin()
        eng.synth_fill(0, 0, n, 77)
        data = eng.device_download(0, 0, n)
        np.testing.assert_array_equal(data[:4096], oracle_lib.synth_fill(0, 4096, 77))
        eng.device_upload(0, hi, data)
        eng.device_submit(0, hi + n, [_ext(0, n, n, MKSNAP_X_CRC | MKSNAP_X_CDC), _ext(hi, n, 0, MKSNAP_X_CRC | MKSNAP_X_CDC)])
        res = eng.finish()
        assert eng.ctx_crc32(res) == zlib.crc32(data.tobytes(), zlib.crc32(data.tobytes()))
        want = oracle_lib.chunk_table(data, [0], [n])
        assert res.n_chunks == 2 * want["n_chunks"] and res.n_unique == want["n_unique"]
        ends, digs = eng.get_chunks(res.n_chunks)
        k = want["n_chunks"]
        np.testing.assert_array_equal(ends[:k], want["ends"])
        np.testing.assert_array_equal(ends[k:], want["ends"] + np.uint64(hi))
        np.testing.assert_array_equal(digs[:k], want["digests"])
        np.testing.assert_array_equal(digs[k:], want["digests"])
        assert bytes(res.root) == want["root"]


@pytest.fixture(scope="module")
def eng():
    from makisu_b200.abi import Engine
    e = Engine(device=0, device_arena_bytes=3 * GiB, n_host_arenas=0, max_extents=1 << 17)
    yield e
    e.close()


def _pack_files(files):
    offs, pos = [], 0
    for f in files:
        pos = (pos + 511) // 512 * 512
        offs.append(pos)
        pos += len(f)
    arena = np.zeros((pos + 511) // 512 * 512 + 512, dtype=np.uint8)
    for f, o in zip(files, offs):
        arena[o:o + len(f)] = f
    return arena, offs


def _check_against_oracle(eng, oracle_lib, files):
    arena, offs = _pack_files(files)
    lens = [len(f) for f in files]
    eng.begin()
    eng.device_upload(0, 0, arena)
    ext, total = crc_extents(offs, lens, list(range(len(files))), flags_extra=[2] * len(files))
    eng.device_submit(0, arena.size, ext)
    res = eng.finish()
    want = oracle_lib.chunk_table(arena, offs, lens)
    crc = 0
    for f in files:
        crc = zlib.crc32(f.tobytes(), crc)
    assert eng.ctx_crc32(res) == crc
    assert (res.n_chunks, res.n_unique) == (want["n_chunks"], want["n_unique"]) and bytes(res.root) == want["root"]
    ends, digs = eng.get_chunks(res.n_chunks)
    np.testing.assert_array_equal(ends, want["ends"])
    np.testing.assert_array_equal(digs, want["digests"])
    return res


def test_zipf_sized_files(eng, oracle_lib):
    """BASELINE configs[3] at reduced scale: sizes ~ Zipf(1.1) clipped to [10 B, 64 MiB], long tail included."""
    rng = np.random.default_rng(0xC4)
    sizes = np.clip((rng.zipf(1.1, 600).astype(np.float64) * 10), 10, 64 << 20).astype(np.int64)
    sizes[0] = 64 << 20
    sizes = sizes[np.cumsum(sizes) < 400 << 20]
    files = [rng.integers(0, 256, int(s), dtype=np.uint8) for s in sizes]
    res = _check_against_oracle(eng, oracle_lib, files)
    assert res.n_files == len(files)


def test_duplicate_heavy_context(eng, oracle_lib):
    """BASELINE configs[4] at reduced scale: files assembled from a pool holding 20 % unique content, in
    256 KiB pieces (>= max chunk, so CDC resynchronises inside every piece)."""
    rng = np.random.default_rng(0xC5)
    piece = 256 << 10
    n_pieces = 800
    pool = [rng.integers(0, 256, piece, dtype=np.uint8) for _ in range(n_pieces // 5)]
    files = []
    for _ in range(40):
        k = int(rng.integers(5, 40))
        files.append(np.concatenate([pool[int(rng.integers(0, len(pool)))] for _ in range(k)]))
    res = _check_against_oracle(eng, oracle_lib, files)
    ratio = res.n_unique / res.n_chunks
    assert 0.1 < ratio < 0.5, ratio


def test_one_long_file(eng, oracle_lib):
    rng = np.random.default_rng(8)
    _check_against_oracle(eng, oracle_lib, [rng.integers(0, 256, 200 << 20, dtype=np.uint8)])


def test_properties_at_2gib(eng, oracle_lib):
    """No oracle pass over the data: (i) one submit == two submits == three reordered submits (CRC linearity,
    table order independence); (ii) idempotence; (iii) crc(A||B) == combine(crc(A), crc(B), |B|);
    (iv) duplicating the context leaves the table and root unchanged and doubles n_chunks."""
    from makisu_b200.abi import MKSNAP_X_CDC, MKSNAP_X_CRC
    n_files, fb = 4096, 512 << 10
    total = n_files * fb
    eng.begin()
    eng.synth_fill(0, 0, total, 0xABC)
    eng.finish()
    fl = MKSNAP_X_CRC | MKSNAP_X_CDC

    def run(groups):
        eng.begin()
        for g in groups:
            eng.device_submit(0, total, [_ext(i * fb, fb, total - (i + 1) * fb, fl) for i in g])
        r = eng.finish()
        return eng.ctx_crc32(r), r.n_chunks, r.n_unique, bytes(r.root)

    allf = list(range(n_files))
    one = run([allf])
    assert one == run([allf])                                        # idempotent
    assert one == run([allf[: n_files // 3], allf[n_files // 3:]])   # split submits
    perm = list(np.random.default_rng(1).permutation(n_files))
    assert one == run([perm[:1000], perm[1000:3000], perm[3000:]])   # any order: suffixes carry the position
    # (iii) CRC of the two halves, combined on the host
    half = n_files // 2
    eng.begin()
    eng.device_submit(0, total, [_ext(i * fb, fb, (half - 1 - i) * fb, MKSNAP_X_CRC) for i in range(half)])
    ra = eng.finish()
    eng.begin()
    eng.device_submit(0, total, [_ext(i * fb, fb, (n_files - 1 - i) * fb, MKSNAP_X_CRC) for i in range(half, n_files)])
    rb = eng.finish()
    assert oracle_lib.L().mko_crc32_combine(eng.ctx_crc32(ra), eng.ctx_crc32(rb), (n_files - half) * fb) == one[0]
    # (iv) the same files twice (CDC only)
    eng.begin()
    eng.device_submit(0, total, [_ext(i * fb, fb, 0, MKSNAP_X_CDC) for i in allf + allf])
    rd = eng.finish()
    assert rd.n_chunks == 2 * one[1] and rd.n_unique == one[2] and bytes(rd.root) == one[3]
    # spot check: first and last file against the oracle
    for i in (0, n_files - 1):
        d = eng.device_download(0, i * fb, fb)
        np.testing.assert_array_equal(d, oracle_lib.synth_fill(i * fb, fb, 0xABC))


def test_25m_row_table_sort_unique():
    """Scale test of K3 (radix sort + unique) and the Merkle root: tiny CDC parameters turn 4 GiB of random
    bytes into > 20 M distinct chunks.  The table must equal numpy's
    sorted-unique of the chunk digests (bytewise order == big-endian u64 column order)."""
    from makisu_b200.abi import CdcParams, Engine, MKSNAP_X_CDC
    p = CdcParams(64, 128, 256, 8, 6)
    n_files, fb = 4096, 1 << 20
    with Engine(device=0, device_arena_bytes=4 * GiB, max_extents=8192, max_chunks=40_000_000, cdc=p) as e:
        e.begin()
        e.synth_fill(0, 0, n_files * fb, 0x5157)
        e.device_submit(0, n_files * fb, [_ext(i * fb, fb, 0, MKSNAP_X_CDC) for i in range(n_files)])
        r = e.finish()
        assert r.n_chunks > 24_000_000
        _, d = e.get_chunks(r.n_chunks)
        k = d.view(">u8").reshape(-1, 4)
        order = np.lexsort((k[:, 3], k[:, 2], k[:, 1], k[:, 0]))
        ks = k[order]
        distinct = 1 + int((ks[1:] != ks[:-1]).any(axis=1).sum())
        # (a handful of 1-byte file tails share a byte value, so distinct is a few less than n_chunks)
        assert r.n_chunks - 50 < distinct <= r.n_chunks
        assert r.n_unique == distinct
        t = e.get_table(r.n_unique).view(">u8").reshape(-1, 4)
        np.testing.assert_array_equal(t, ks[np.concatenate([[True], (ks[1:] != ks[:-1]).any(axis=1)])])


def test_nccl_gather_world_of_one():
    """The all-gather/merge path with a 1-rank communicator (runs on a single-GPU box): result == finish()."""
    from makisu_b200.abi import Engine, MKSNAP_X_CDC, MKSNAP_X_CRC
    fb = 1 << 20
    with Engine(device=0, device_arena_bytes=256 << 20, max_extents=4096) as e:
        e.comm_init(Engine.comm_unique_id(), 1, 0)
        e.begin()
        e.synth_fill(0, 0, 200 * fb, 99)
        e.device_submit(0, 200 * fb, [_ext(i * fb, fb, (199 - i) * fb, MKSNAP_X_CDC | MKSNAP_X_CRC) for i in range(200)])
        r = e.finish()
        g = e.allgather_tables()
        assert (g.n_chunks, g.n_unique, bytes(g.root), e.ctx_crc32(g)) == (r.n_chunks, r.n_unique, bytes(r.root), e.ctx_crc32(r))


@pytest.mark.parametrize("params", [(64, 128, 256, 8, 6), (64, 256, 1024, 4, 3), (4096, 16384, 131072, 16, 12)])
def test_big_file_selection_paths(oracle_lib, params):
    """Files >= 4 MiB take the CTA-per-file selection kernel (shared-memory window); very dense masks force its
    global-memory fallback.  Mixed with small files; everything must equal the oracle."""
    from makisu_b200.abi import CdcParams, Engine
    import oracle.lib as olib
    p = CdcParams(*params)
    op = olib.CdcParams(*params)
    rng = np.random.default_rng(17)
    files = [rng.integers(0, 256, n, dtype=np.uint8) for n in (9 << 20, 1000, (4 << 20), (4 << 20) - 1, 5 << 20, 0, 70000)]
    files.append(np.zeros(6 << 20, dtype=np.uint8))
    arena, offs = _pack_files(files)
    lens = [len(f) for f in files]
    with Engine(device=0, device_arena_bytes=64 << 20, max_extents=64, max_chunks=2_000_000, cdc=p) as e:
        e.begin()
        e.device_upload(0, 0, arena)
        e.device_submit(0, arena.size, cdc_extents(offs, lens))
        res = e.finish()
        want = oracle_lib.chunk_table(arena, offs, lens, op)
        assert (res.n_chunks, res.n_unique) == (want["n_chunks"], want["n_unique"]) and bytes(res.root) == want["root"]
        ends, digs = e.get_chunks(res.n_chunks)
        np.testing.assert_array_equal(ends, want["ends"])
        np.testing.assert_array_equal(digs, want["digests"])


def test_config1_exact_1k_files_of_1MiB(oracle_lib):
    """BASELINE configs[1] exactly: synthetic 1k files x 1 MiB on one B200, every cut point, every chunk digest, the
    sorted-unique table and the root diffed bit-exact against the oracle, plus the cacheID CRC against zlib."""
    import zlib
    from makisu_b200.abi import Engine
    n, fb = 1000, 1 << 20
    with Engine(device=0, device_arena_bytes=(n * fb) + (1 << 20), max_extents=2 * n + 16) as e:
        e.begin()
        e.synth_fill(0, 0, n * fb, 0xC2)
        arena = e.device_download(0, 0, n * fb)
        np.testing.assert_array_equal(arena[:4096], oracle_lib.synth_fill(0, 4096, 0xC2))   # same generator on both sides
        offs, lens = [i * fb for i in range(n)], [fb] * n
        ext, total = crc_extents(offs, lens, list(range(n)))
        for x in ext:
            x.flags |= 2  # MKSNAP_X_CDC
        e.device_submit(0, n * fb, ext)
        res = e.finish()
        want = oracle_lib.chunk_table(arena, offs, lens)
        assert (res.n_chunks, res.n_unique, bytes(res.root)) == (want["n_chunks"], want["n_unique"], want["root"])
        ends, dig = e.get_chunks(res.n_chunks)
        np.testing.assert_array_equal(ends, want["ends"])
        np.testing.assert_array_equal(dig, want["digests"])
        np.testing.assert_array_equal(e.get_table(res.n_unique), want["table"])
        assert e.ctx_crc32(res) == zlib.crc32(arena.tobytes()) and res.crc_bytes == n * fb
        # ... and against the committed fingerprint (tests/golden/config1_1k_x_1MiB.json, written by make_golden.py)
        import json
        import os
        from tests.util import table_fingerprint
        g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_1k_x_1MiB.json")))
        got = table_fingerprint(res.n_chunks, res.n_unique, bytes(res.root), ends, dig, e.ctx_crc32(res))
        assert got == {k: g[k] for k in got}


def test_all_zero_context_has_one_long_run_of_duplicates(oracle_lib):
    """2 GiB of zeros: every chunk is the same max-size chunk, i.e. one run of ~16 k identical digests for the tie
    fix-up after the radix sort (k_fix_ties hands runs longer than 64 rows to k_fix_long_runs: a whole CTA checks that
    they are duplicates instead of one thread walking them).  Table = 1 row (+ the tail chunk), root vs the oracle."""
    from makisu_b200.abi import Engine
    n, fb = 64, 32 << 20
    with Engine(device=0, device_arena_bytes=n * fb, max_extents=n + 16) as e:
        e.begin()
        e.memset(0, 0, n * fb, 0)
        e.device_submit(0, n * fb, cdc_extents([i * fb for i in range(n)], [fb] * n))
        res = e.finish()
        small = np.zeros(fb, dtype=np.uint8)
        want = oracle_lib.chunk_table(small, [0], [fb])     # one file is enough: all files are identical
        assert res.n_chunks == n * want["n_chunks"] and res.n_unique == want["n_unique"] == 1
        assert bytes(res.root) == want["root"]


def test_one_gib_file_through_256_mib_arenas(oracle_lib):
    """BASELINE configs[3] has 1 GiB files; the e2e path uses arenas far smaller than that.  One 1 GiB file (content from
    the device generator, with a 6 MiB zero run across an arena boundary) travels through 256 MiB pinned arenas in five
    pieces (MKSNAP_X_MORE / MKSNAP_X_CONT) next to small neighbours; chunk count, every digest, table and root equal the
    oracle's over the undivided file (lib/tario/write.go:45: io.CopyN of any size)."""
    from makisu_b200.abi import Engine
    from tests.test_gpu_parity import _split_submit
    big_n = 1 << 30
    arena = 256 << 20
    with Engine(device=0, device_arena_bytes=arena, n_host_arenas=2, host_arena_bytes=arena, max_extents=64,
                max_chunks=(big_n >> 12) + 4096) as e:
        carry = e.limits().carry_bytes
        # content: generator output fetched in slices (the engine's own slot is only 256 MiB)
        big = np.empty(big_n, dtype=np.uint8)
        e.begin()
        for o in range(0, big_n, arena):
            e.synth_fill(0, 0, arena, 0xB16 + o)
            big[o:o + arena] = e.device_download(0, 0, arena)
        e.finish()
        first = arena - 4096 - (1 << 20)                       # first piece: what is left of arena 0 behind a 1 MiB neighbour
        big[first - (3 << 20):first + (3 << 20)] = 0            # zero run straddling the first boundary: forced cuts across it
        per = (arena - carry) // 512 * 512
        pieces, left = [first], big_n - first
        while left:
            n = min(per, left)
            pieces.append(n)
            left -= n
        small1 = oracle_lib.synth_fill(0, 1 << 20, 7)
        small2 = oracle_lib.synth_fill(0, 300_001, 8)
        files = [small1, big, small2]
        res = _split_submit(e, files, [[len(small1)], pieces, [len(small2)]], carry, True)
        offs = [0, 1 << 20, (1 << 20) + big_n]
        whole = np.concatenate([small1, big, small2])
        want = oracle_lib.chunk_table(whole, offs, [len(f) for f in files])
        assert (res.n_chunks, res.n_unique, bytes(res.root)) == (want["n_chunks"], want["n_unique"], want["root"])
        assert res.n_files == 3 and res.cdc_bytes == whole.size
        _, dig = e.get_chunks(res.n_chunks)
        np.testing.assert_array_equal(dig, want["digests"])
