/*
 * mkoracle.h — CPU ORACLE for the makisu snapshot+hash hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (makisu_b200/csrc, libmksnap.so) never links, imports or calls it and
 * fails loudly when the CUDA library is missing.
 *
 * What is restated here (plain C, scalar, single thread):
 *   - CRC-32/IEEE as Go's hash/crc32 computes it for
 *     reference lib/builder/step/add_copy_step.go:104 (crc32.NewIEEE) and
 *     base_step.go:64 (crc32.ChecksumIEEE).  hash/crc32 is Go stdlib
 *     (go1.14, not vendored under /root/reference); the algorithm is the
 *     published reflected CRC-32, poly 0xEDB88320, init/xorout 0xFFFFFFFF.
 *   - SHA-256 (FIPS 180-4) as Go's crypto/sha256 computes it for
 *     reference lib/builder/step/common.go:44-45 (tarDigester/gzipDigester)
 *     and lib/docker/image/digester.go:35-55.
 *   - The chunk-table spec frozen in DESIGN.md section 3 (Roll-32 CDC, per-chunk
 *     SHA-256, sorted-unique table, fan-out-256 Merkle root).  This part has
 *     NO reference counterpart => "parity unpinned": it is
 *     pinned only by this restatement and the golden vectors in tests/golden.
 */
#ifndef MKORACLE_H
#define MKORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- CRC-32/IEEE (zlib / Go hash/crc32 semantics) ---------------------- */
/* crc = running value as returned by Sum32(); start with 0. */
uint32_t mko_crc32_update(uint32_t crc, const uint8_t *p, size_t n);
/* "pure" CRC: init 0, no final xor -- the GF(2)-linear part used by the
 * device decomposition; restated here so tests can check the algebra. */
uint32_t mko_crc32_pure(const uint8_t *p, size_t n);
/* x^(8*nbytes) mod P in the reflected 32-bit representation. */
uint32_t mko_crc32_xpow8n(uint64_t nbytes);
/* (a*b) mod P, reflected representation. */
uint32_t mko_crc32_mulmod(uint32_t a, uint32_t b);
/* crc(A||B) from crc(A), crc(B), |B| (zlib crc32_combine semantics). */
uint32_t mko_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);

/* ---- SHA-256 ----------------------------------------------------------- */
typedef struct {
    uint32_t h[8];
    uint64_t nbytes;
    uint8_t buf[64];
    uint32_t fill;
} mko_sha256_ctx;
void mko_sha256_init(mko_sha256_ctx *c);
void mko_sha256_update(mko_sha256_ctx *c, const uint8_t *p, size_t n);
void mko_sha256_final(mko_sha256_ctx *c, uint8_t out[32]);
void mko_sha256(const uint8_t *p, size_t n, uint8_t out[32]);
/* baseline timing: SHA-NI for whole blocks when the CPU has it (else identical to mko_sha256_update) */
int mko_have_sha_ni(void);
void mko_sha256_update_fast(mko_sha256_ctx *c, const uint8_t *p, size_t n);

/* ---- Roll-32 content-defined chunking ------------------ */
typedef struct {
    uint32_t min_size;     /* 4096   */
    uint32_t normal_size;  /* 16384: strict mask below, loose mask at/after */
    uint32_t max_size;     /* 131072 */
    uint32_t strict_bits;  /* 16: candidate iff the top strict_bits bits of h are all ones */
    uint32_t loose_bits;   /* 12 */
} mko_cdc_params;
void mko_cdc_default_params(mko_cdc_params *p);
/* the multiplier M of h_i = h_{i-1}*M + u_i (u_i = little-endian 32-bit word ending at byte i) */
uint32_t mko_roll_multiplier(void);
/* Rolling hash of the 32-position (35-byte) window ending at data[i] (bytes before the start
 * of the buffer count as absent, i.e. state starts at 0 at data[0] and missing bytes read as 0). */
uint32_t mko_roll_at(const uint8_t *data, size_t i);
/* Chunk one file.  Writes chunk END offsets (exclusive, relative to data)
 * into ends[0..cap); returns the number of chunks (may exceed cap: call
 * again with a larger buffer).  An empty file has 0 chunks. */
size_t mko_cdc_cuts(const uint8_t *data, size_t len, const mko_cdc_params *p,
                    uint64_t *ends, size_t cap);

/* ---- chunk table ------------------------------------------------------- */
/* Sort n 32-byte digests bytewise ascending and drop duplicates in place;
 * returns the unique count. */
size_t mko_sort_unique_digests(uint8_t *digests, size_t n);
/* Fan-out-256 Merkle root over n 32-byte digests (already sorted-unique). */
void mko_merkle_root(const uint8_t *digests, size_t n, uint8_t out[32]);

typedef struct {
    uint64_t n_chunks;
    uint64_t n_unique;
    uint8_t root[32];
} mko_table_summary;
/* Whole chunk-table pipeline over files packed in one arena.
 * offs/lens: per-file extents.  If cut_ends != NULL it receives absolute
 * arena END offsets of every chunk in file order (cap entries at most);
 * if digests != NULL it receives per-chunk digests in the same order; if
 * table != NULL it receives the sorted-unique table.  Returns 0, or -1 if
 * cap was too small (summary->n_chunks then holds the needed size). */
int mko_chunk_table(const uint8_t *arena, const uint64_t *offs,
                    const uint64_t *lens, size_t n_files,
                    const mko_cdc_params *p, uint64_t *cut_ends,
                    uint8_t *digests, uint8_t *table, size_t cap,
                    mko_table_summary *summary);

/* baseline timing (bench.py): CRC-32 + Roll-32 CDC + per-chunk SHA-256 (fast path) of a slice of files; thread safe */
size_t mko_step_same_work(const uint8_t *arena, const uint64_t *offs, const uint64_t *lens, size_t n_files,
                          const mko_cdc_params *p, uint32_t *crcs, uint8_t *digests, size_t cap, uint32_t *sink);

/* ---- synthetic content generator shared with the device -
 * byte stream = little-endian u64 words, word i = mix64(seed + i), where i is
 * the absolute 8-byte word index in the arena.  dst covers arena bytes
 * [byte_off, byte_off+n); byte_off and n must be multiples of 8. */
void mko_synth_fill(uint8_t *dst, uint64_t byte_off, uint64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
