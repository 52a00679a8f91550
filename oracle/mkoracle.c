/*
 * mkoracle.c — CPU ORACLE (test infrastructure only; see mkoracle.h).
 *
 * Scalar single-thread restatement of the arithmetic on the makisu
 * snapshot+hash hot path.  Reference call sites (all under /root/reference):
 *   CRC-32/IEEE   lib/builder/step/add_copy_step.go:104-119 (context cacheID),
 *                 lib/builder/step/base_step.go:62-67, lib/builder/build_plan.go:96-97
 *   SHA-256       lib/builder/step/common.go:44-55,86-87 (TarDigest, gzip digest),
 *                 lib/docker/image/digester.go:35-55
 * Both algorithms live in the Go standard library (go1.14, hash/crc32 and
 * crypto/sha256), which is not vendored under /root/reference; the published
 * algorithms (reflected CRC-32 poly 0xEDB88320; FIPS 180-4) are restated and
 * pinned against the reference's own fixtures in tests/test_oracle_golden.py.
 *
 * Roll-32 CDC / chunk table / Merkle root: no reference counterpart
 * ("parity unpinned"); this file is the normative statement of DESIGN.md section 3.
 */
#include "mkoracle.h"

#include <stdlib.h>
#include <string.h>

/* ======================================================================= */
/* CRC-32/IEEE                                                             */
/* ======================================================================= */

#define CRC_POLY_REFLECTED 0xEDB88320u

static uint32_t crc_tab[8][256];
static int crc_tab_ready = 0;

static void crc_init_tables(void)
{
    if (crc_tab_ready)
        return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++)
            c = (c & 1) ? (c >> 1) ^ CRC_POLY_REFLECTED : c >> 1;
        crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++)
            crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xFF];
    crc_tab_ready = 1;
}

/* register update without init/xorout handling (slicing-by-8) */
static uint32_t crc_raw(uint32_t r, const uint8_t *p, size_t n)
{
    crc_init_tables();
    while (n && ((uintptr_t)p & 7)) {
        r = (r >> 8) ^ crc_tab[0][(r ^ *p++) & 0xFF];
        n--;
    }
    while (n >= 8) {
        uint32_t lo, hi;
        memcpy(&lo, p, 4);
        memcpy(&hi, p + 4, 4);
        lo ^= r;
        r = crc_tab[7][lo & 0xFF] ^ crc_tab[6][(lo >> 8) & 0xFF] ^
            crc_tab[5][(lo >> 16) & 0xFF] ^ crc_tab[4][lo >> 24] ^
            crc_tab[3][hi & 0xFF] ^ crc_tab[2][(hi >> 8) & 0xFF] ^
            crc_tab[1][(hi >> 16) & 0xFF] ^ crc_tab[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    while (n--)
        r = (r >> 8) ^ crc_tab[0][(r ^ *p++) & 0xFF];
    return r;
}

uint32_t mko_crc32_update(uint32_t crc, const uint8_t *p, size_t n)
{
    return crc_raw(crc ^ 0xFFFFFFFFu, p, n) ^ 0xFFFFFFFFu;
}

uint32_t mko_crc32_pure(const uint8_t *p, size_t n)
{
    return crc_raw(0, p, n);
}

/* Reflected representation: bit 31 is the x^0 coefficient, bit 0 is x^31. */
uint32_t mko_crc32_mulmod(uint32_t a, uint32_t b)
{
    uint32_t acc = 0;
    for (int k = 31; k >= 0; k--) { /* walk a from x^0 up to x^31 */
        if ((a >> k) & 1)
            acc ^= b;
        b = (b & 1) ? (b >> 1) ^ CRC_POLY_REFLECTED : b >> 1; /* b *= x */
    }
    return acc;
}

uint32_t mko_crc32_xpow8n(uint64_t nbytes)
{
    uint32_t result = 0x80000000u; /* x^0 */
    uint32_t sq = 0x00800000u;     /* x^8 */
    while (nbytes) {
        if (nbytes & 1)
            result = mko_crc32_mulmod(result, sq);
        sq = mko_crc32_mulmod(sq, sq);
        nbytes >>= 1;
    }
    return result;
}

uint32_t mko_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
    return mko_crc32_mulmod(mko_crc32_xpow8n(len_b), crc_a) ^ crc_b;
}

/* ======================================================================= */
/* SHA-256 (FIPS 180-4)                                                    */
/* ======================================================================= */

static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha256_block(uint32_t st[8], const uint8_t *p)
{
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) |
               ((uint32_t)p[4 * i + 2] << 8) | (uint32_t)p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    uint32_t e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + SHA_K[i] + w[i];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1;
        d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d;
    st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

void mko_sha256_init(mko_sha256_ctx *c)
{
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                   0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(c->h, iv, sizeof iv);
    c->nbytes = 0;
    c->fill = 0;
}

void mko_sha256_update(mko_sha256_ctx *c, const uint8_t *p, size_t n)
{
    c->nbytes += n;
    if (c->fill) {
        size_t take = 64 - c->fill;
        if (take > n)
            take = n;
        memcpy(c->buf + c->fill, p, take);
        c->fill += (uint32_t)take;
        p += take;
        n -= take;
        if (c->fill < 64)
            return;
        sha256_block(c->h, c->buf);
        c->fill = 0;
    }
    while (n >= 64) {
        sha256_block(c->h, p);
        p += 64;
        n -= 64;
    }
    if (n) {
        memcpy(c->buf, p, n);
        c->fill = (uint32_t)n;
    }
}

void mko_sha256_final(mko_sha256_ctx *c, uint8_t out[32])
{
    uint64_t bits = c->nbytes * 8;
    uint8_t pad[72];
    size_t padlen = (c->fill < 56) ? 56 - c->fill : 120 - c->fill;
    memset(pad, 0, sizeof pad);
    pad[0] = 0x80;
    for (int i = 0; i < 8; i++)
        pad[padlen + i] = (uint8_t)(bits >> (56 - 8 * i));
    mko_sha256_update(c, pad, padlen + 8);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(c->h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(c->h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(c->h[i] >> 8);
        out[4 * i + 3] = (uint8_t)c->h[i];
    }
}

void mko_sha256(const uint8_t *p, size_t n, uint8_t out[32])
{
    mko_sha256_ctx c;
    mko_sha256_init(&c);
    mko_sha256_update(&c, p, n);
    mko_sha256_final(&c, out);
}

/* ======================================================================= */
/* Roll-32 CDC (no reference counterpart)                   */
/* ======================================================================= */
/* Rabin-Karp style polynomial rolling hash over the little-endian 32-bit word that ENDS at each byte:
 *   u_i = b[i-3] | b[i-2]<<8 | b[i-1]<<16 | b[i]<<24
 *   h_i = h_{i-1} * M + u_i   (mod 2^32),   M = 2 * (odd)  =>  M^32 = 0 (mod 2^32):
 *   h_i = sum_{k<32} u_{i-k} * M^k  -- a window of 32 positions = 35 bytes, no table.
 * Candidate at byte i iff the top `bits` bits of h_i are all ONES (h_i >= 2^32 - 2^(32-bits)); a run of zero bytes
 * hashes to 0 and is therefore never a candidate. */

#define ROLL_MULT 0x9E3779BAu /* 2 * 0x4F1BBCDD */

/* splitmix64 finaliser: the synthetic-content generator below (mko_synth_fill) */
static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
#define GOLDEN64 0x9E3779B97F4A7C15ull

uint32_t mko_roll_multiplier(void) { return ROLL_MULT; }

void mko_cdc_default_params(mko_cdc_params *p)
{
    p->min_size = 4096;
    p->normal_size = 16384;
    p->max_size = 131072;
    p->strict_bits = 16;
    p->loose_bits = 12;
}

static inline uint32_t word_ending_at(const uint8_t *data, size_t k)
{
    /* bytes before data[0] do not exist for the callers below except in mko_roll_at, which treats them as absent (0) */
    uint32_t u = (uint32_t)data[k] << 24;
    if (k >= 1) u |= (uint32_t)data[k - 1] << 16;
    if (k >= 2) u |= (uint32_t)data[k - 2] << 8;
    if (k >= 3) u |= (uint32_t)data[k - 3];
    return u;
}

uint32_t mko_roll_at(const uint8_t *data, size_t i)
{
    size_t start = i >= 31 ? i - 31 : 0;
    uint32_t h = 0;
    for (size_t k = start; k <= i; k++)
        h = h * ROLL_MULT + word_ending_at(data, k);
    return h;
}

size_t mko_cdc_cuts(const uint8_t *data, size_t len, const mko_cdc_params *p,
                    uint64_t *ends, size_t cap)
{
    const uint32_t strict_thr = 0u - (1u << (32 - p->strict_bits));
    const uint32_t loose_thr = 0u - (1u << (32 - p->loose_bits));
    size_t n = 0, prev = 0;
    while (prev < len) {
        size_t rem = len - prev;
        size_t cut;
        if (rem <= p->min_size) {
            cut = len;
        } else {
            size_t limit = rem < p->max_size ? rem : p->max_size;
            /* window of 32 positions (35 bytes) ending at the last byte of a min-size chunk: min_size >= 64, so it
             * lies inside the chunk */
            uint32_t h = 0;
            for (size_t k = prev + p->min_size - 32; k < prev + p->min_size; k++)
                h = h * ROLL_MULT + word_ending_at(data + prev, k - prev);
            size_t L = p->min_size;
            cut = 0;
            for (;;) {
                uint32_t thr = L < p->normal_size ? strict_thr : loose_thr;
                if (h >= thr) {
                    cut = prev + L;
                    break;
                }
                if (L == limit)
                    break;
                h = h * ROLL_MULT + word_ending_at(data + prev, L);
                L++;
            }
            if (!cut)
                cut = prev + limit; /* forced at max, or end of file */
        }
        if (n < cap)
            ends[n] = cut;
        n++;
        prev = cut;
    }
    return n;
}

/* ======================================================================= */
/* chunk table                                                             */
/* ======================================================================= */

static int cmp_digest(const void *a, const void *b) { return memcmp(a, b, 32); }

size_t mko_sort_unique_digests(uint8_t *d, size_t n)
{
    if (n == 0)
        return 0;
    qsort(d, n, 32, cmp_digest);
    size_t m = 1;
    for (size_t i = 1; i < n; i++)
        if (memcmp(d + 32 * i, d + 32 * (m - 1), 32) != 0) {
            if (i != m)
                memcpy(d + 32 * m, d + 32 * i, 32);
            m++;
        }
    return m;
}

void mko_merkle_root(const uint8_t *digests, size_t n, uint8_t out[32])
{
    if (n == 0) {
        mko_sha256((const uint8_t *)"", 0, out);
        return;
    }
    size_t cur_n = n;
    const uint8_t *cur = digests;
    uint8_t *owned = NULL;
    for (;;) {
        size_t next_n = (cur_n + 255) / 256;
        uint8_t *next = (uint8_t *)malloc(next_n * 32);
        for (size_t j = 0; j < next_n; j++) {
            size_t cnt = cur_n - j * 256 < 256 ? cur_n - j * 256 : 256;
            mko_sha256(cur + j * 256 * 32, cnt * 32, next + j * 32);
        }
        free(owned);
        owned = next;
        cur = next;
        cur_n = next_n;
        if (cur_n == 1)
            break;
    }
    memcpy(out, cur, 32);
    free(owned);
}

int mko_chunk_table(const uint8_t *arena, const uint64_t *offs, const uint64_t *lens,
                    size_t n_files, const mko_cdc_params *p, uint64_t *cut_ends,
                    uint8_t *digests, uint8_t *table, size_t cap,
                    mko_table_summary *summary)
{
    /* pass 1: count */
    size_t total = 0;
    for (size_t f = 0; f < n_files; f++)
        total += mko_cdc_cuts(arena + offs[f], lens[f], p, NULL, 0);
    summary->n_chunks = total;
    summary->n_unique = 0;
    if ((cut_ends || digests || table) && total > cap)
        return -1;
    uint64_t *ends = cut_ends ? cut_ends : (uint64_t *)malloc((total ? total : 1) * 8);
    uint8_t *dg = digests ? digests : (uint8_t *)malloc((total ? total : 1) * 32);
    size_t k = 0;
    for (size_t f = 0; f < n_files; f++) {
        size_t c = mko_cdc_cuts(arena + offs[f], lens[f], p, ends + k, total - k);
        uint64_t prev = 0;
        for (size_t j = 0; j < c; j++) {
            uint64_t e = ends[k + j];
            mko_sha256(arena + offs[f] + prev, e - prev, dg + 32 * (k + j));
            prev = e;
            ends[k + j] = offs[f] + e; /* absolute arena offset */
        }
        k += c;
    }
    uint8_t *tb = table ? table : (uint8_t *)malloc((total ? total : 1) * 32);
    memcpy(tb, dg, total * 32);
    size_t uniq = mko_sort_unique_digests(tb, total);
    summary->n_unique = uniq;
    mko_merkle_root(tb, uniq, summary->root);
    if (!table)
        free(tb);
    if (!digests)
        free(dg);
    if (!cut_ends)
        free(ends);
    return 0;
}

/* Baseline timing only (bench.py cpu_best_effort): the work of one GPU step over a slice of files -- CRC-32 of every
 * file, Roll-32 CDC, SHA-256 of every chunk through the SHA-NI path when the CPU has it.  Thread safe (no shared
 * state): bench.py runs one call per host thread on disjoint slices.  Returns the number of chunks; digests may be
 * NULL (then the last digest is folded into *sink so the work cannot be optimised away). */
size_t mko_step_same_work(const uint8_t *arena, const uint64_t *offs, const uint64_t *lens, size_t n_files,
                          const mko_cdc_params *p, uint32_t *crcs, uint8_t *digests, size_t cap, uint32_t *sink)
{
    size_t k = 0;
    uint64_t max_len = 0;
    for (size_t f = 0; f < n_files; f++)
        if (lens[f] > max_len)
            max_len = lens[f];
    const size_t ends_cap = (size_t)(max_len / (p->min_size ? p->min_size : 1)) + 2;
    uint64_t *ends = (uint64_t *)malloc(ends_cap * 8);
    uint8_t dg[32] = {0};
    uint32_t acc = 0;
    for (size_t f = 0; f < n_files; f++) {
        const uint8_t *base = arena + offs[f];
        uint32_t c = mko_crc32_update(0, base, lens[f]);
        if (crcs)
            crcs[f] = c;
        acc ^= c;
        size_t n = mko_cdc_cuts(base, lens[f], p, ends, ends_cap);
        uint64_t prev = 0;
        for (size_t j = 0; j < n; j++) {
            mko_sha256_ctx ctx;
            mko_sha256_init(&ctx);
            mko_sha256_update_fast(&ctx, base + prev, ends[j] - prev);
            mko_sha256_final(&ctx, (digests && k < cap) ? digests + 32 * k : dg);
            prev = ends[j];
            k++;
        }
    }
    free(ends);
    if (sink)
        *sink = acc ^ dg[0];
    return k;
}

/* ======================================================================= */
/* synthetic content                                                       */
/* ======================================================================= */

void mko_synth_fill(uint8_t *dst, uint64_t byte_off, uint64_t n, uint64_t seed)
{
    uint64_t w0 = byte_off / 8;
    for (uint64_t i = 0; i < n / 8; i++) {
        uint64_t v = mix64(seed + (w0 + i + 1) * GOLDEN64);
        memcpy(dst + 8 * i, &v, 8);
    }
}

/* ======================================================================= */
/* SHA-256 with the x86 SHA extensions (baseline timing only)              */
/* ======================================================================= */
/* The reference's Go 1.14 crypto/sha256 uses hand-written AVX2 assembly on amd64; a scalar C loop would
 * understate the CPU baseline.  This path (SHA-NI, the fastest single-thread SHA-256 this host offers) is used
 * by bench.py's cpu_baseline / --impl reference when the CPU has it; tests check it against the scalar code. */
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>

int mko_have_sha_ni(void)
{
    unsigned a, b, c, d;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d))
        return 0;
    return (b >> 29) & 1; /* CPUID.7.0:EBX.SHA */
}

__attribute__((target("sha,sse4.1,ssse3"))) static void sha256_blocks_ni(uint32_t st[8], const uint8_t *p, size_t nblk)
{
    const __m128i MASK = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i TMP = _mm_loadu_si128((const __m128i *)&st[0]);
    __m128i STATE1 = _mm_loadu_si128((const __m128i *)&st[4]);
    TMP = _mm_shuffle_epi32(TMP, 0xB1);          /* CDAB */
    STATE1 = _mm_shuffle_epi32(STATE1, 0x1B);    /* EFGH */
    __m128i STATE0 = _mm_alignr_epi8(TMP, STATE1, 8); /* ABEF */
    STATE1 = _mm_blend_epi16(STATE1, TMP, 0xF0);      /* CDGH */
    while (nblk--) {
        const __m128i S0 = STATE0, S1 = STATE1;
        __m128i M[4], MSG;
        for (int i = 0; i < 4; i++)
            M[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + 16 * i)), MASK);
        for (int r = 0; r < 16; r++) {
            MSG = _mm_add_epi32(M[r & 3], _mm_loadu_si128((const __m128i *)&SHA_K[4 * r]));
            STATE1 = _mm_sha256rnds2_epu32(STATE1, STATE0, MSG);
            if (r >= 3 && r < 15) { /* schedule words 16+4(r-3) .. : W[t] for the group used at round r+1 */
                __m128i t = _mm_alignr_epi8(M[r & 3], M[(r + 3) & 3], 4);
                M[(r + 1) & 3] = _mm_sha256msg2_epu32(_mm_add_epi32(M[(r + 1) & 3], t), M[r & 3]);
            }
            MSG = _mm_shuffle_epi32(MSG, 0x0E);
            STATE0 = _mm_sha256rnds2_epu32(STATE0, STATE1, MSG);
            if (r >= 1 && r < 13)
                M[(r + 3) & 3] = _mm_sha256msg1_epu32(M[(r + 3) & 3], M[r & 3]);
        }
        STATE0 = _mm_add_epi32(STATE0, S0);
        STATE1 = _mm_add_epi32(STATE1, S1);
        p += 64;
    }
    TMP = _mm_shuffle_epi32(STATE0, 0x1B);       /* FEBA */
    STATE1 = _mm_shuffle_epi32(STATE1, 0xB1);    /* DCHG */
    STATE0 = _mm_blend_epi16(TMP, STATE1, 0xF0); /* DCBA */
    STATE1 = _mm_alignr_epi8(STATE1, TMP, 8);    /* ABEF -> HGFE */
    _mm_storeu_si128((__m128i *)&st[0], STATE0);
    _mm_storeu_si128((__m128i *)&st[4], STATE1);
}
#else
int mko_have_sha_ni(void) { return 0; }
#endif

/* streaming update that uses SHA-NI for whole blocks when available (same context layout) */
void mko_sha256_update_fast(mko_sha256_ctx *c, const uint8_t *p, size_t n)
{
#if defined(__x86_64__)
    static int have = -1;
    if (have < 0)
        have = mko_have_sha_ni();
    if (have) {
        c->nbytes += n;
        if (c->fill) {
            size_t take = 64 - c->fill;
            if (take > n)
                take = n;
            memcpy(c->buf + c->fill, p, take);
            c->fill += (uint32_t)take;
            p += take;
            n -= take;
            if (c->fill < 64)
                return;
            sha256_blocks_ni(c->h, c->buf, 1);
            c->fill = 0;
        }
        if (n >= 64) {
            sha256_blocks_ni(c->h, p, n / 64);
            p += n / 64 * 64;
            n %= 64;
        }
        if (n) {
            memcpy(c->buf, p, n);
            c->fill = (uint32_t)n;
        }
        return;
    }
#endif
    mko_sha256_update(c, p, n);
}
