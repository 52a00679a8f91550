// mksnap_sha_stream.cuh -- K4: SHA-256 of SERIAL streams (TarDigest, per-file digests), sm_100a.
//
// Replaces the tarDigester sink of lib/builder/step/common.go:44-55 (`sha256.New()` behind the tar.Writer):
// one stream = one Merkle-Damgard chain, so a stream cannot be split; what bounds it on a GPU is the time ONE
// warp needs per 64-byte block.  On B200 the rotate/logic ops (SHF, LOP3) issue on the half-rate ALU pipe: a
// warp instruction occupies it for 2 cycles whatever the number of active lanes.  A block costs
//   rounds   64 x (6 SHF + 4 LOP3)        = 640 ALU-pipe ops = 1280 cycles
//   schedule 48 x (6 SHF/SHR + 2 LOP3)    = 384 ALU-pipe ops =  768 cycles  (+ byte swaps, loads, K adds)
// With both in one warp (the round-1 kernel, lane per stream) a stream gets one block per >2000 cycles, and that
// kernel also waited for its un-prefetched global loads: 34.5 MB/s measured.
//
// Here a CTA is a PAIR of warps on two different SM sub-partitions (each with its own ALU pipe):
//   warp 1 ("schedule")  owns the pieces: fetches work, prefetches the next block of every lane one iteration
//                        ahead (ld.global.nc, 4 x 16 B per lane), byte-swaps, builds the padding blocks,
//                        expands W[0..63] and stores W[t]+K[t] to shared memory ([t][lane]: conflict-free);
//   warp 0 ("rounds")    runs nothing but the 64-round chain: per round one LDS (prefetched by the unrolled
//                        code), 6 SHF + 4 LOP3 + the additions; the state never leaves its registers.
// Lane l of both warps works on the same stream, 32 streams per pair, double-buffered hand-over through two
// named barriers per buffer (bar.sync / bar.arrive, 64 threads).  The rounds warp is the bound:
// 1280 cycles per block = 98 MB/s per stream at 1965 MHz, 3.1 GB/s per pair, independent of how many of the 32
// lanes carry a stream; pairs are spread one per SM first so a pair owns its sub-partitions.
// Lanes pull the next piece from a global counter when theirs ends (ragged per-file streams).
//
// Bit-exactness: FIPS 180-4; checked against hashlib and the oracle in tests/test_gpu_parity.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mk {

// midstate of a serial stream that continues in a later submit
struct StreamState {
    uint32_t st[8];
    unsigned long long bytes; // bytes compressed so far (multiple of 64)
    uint32_t open;            // 1 = a piece with MKSNAP_R_MORE was seen and the stream is not finished
    uint32_t pad;
};

#ifndef SS_ADD_MODE
#define SS_ADD_MODE 0 // 0: 3-input adds (IADD3), fewest instructions; 1: every add as IMAD (FMA pipe)
#endif
#ifndef SS_SCHED_WARP
#define SS_SCHED_WARP 1 // which warp of the CTA is the schedule warp (the rounds warp is warp 0); others exit at once
#endif
constexpr int SS_THREADS = 32 * (SS_SCHED_WARP + 1);
constexpr uint32_t SS_ACTIVE = 1u, SS_START_IV = 2u, SS_START_RESUME = 4u, SS_FINAL = 8u, SS_PARK = 16u, SS_EXIT = 32u;

// W[t] + K[t] of the block in flight, [buffer][lane][t]: a lane's 64 words are contiguous (LDS.128 / STS.128, 16
// instead of 64 shared-memory instructions per block and warp) and the lane stride is 68 words so that the 8 lanes of
// one quarter-warp phase of a 128-bit access fall into 8 disjoint groups of 4 banks (68 mod 32 = 4): conflict-free.
constexpr uint32_t SS_LANE_WORDS = 68;
struct SsShared {
    uint4 kw[2][32 * SS_LANE_WORDS / 4];
    uint32_t ctrl[2][32];
    uint32_t sid[2][32];
};

#ifdef SS_PROFILE
// microbenchmark only (k4_microbench.cu): [0] rounds-warp cycles, [1] of which waiting for a full buffer,
// [2] schedule-warp cycles, [3] of which waiting for an empty buffer, [4] blocks; CTA 0 only
__device__ unsigned long long ss_prof[8];
__device__ int ss_dbg; // 1: the schedule warp skips its global loads; 2: the rounds warp skips the rounds; 3: the schedule warp skips the expansion (timing only)
#endif
__device__ __forceinline__ void ss_bar_sync(uint32_t id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void ss_bar_arrive(uint32_t id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ uint32_t ss_rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t ss_add(uint32_t a, uint32_t b, uint32_t one)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}
// a + b + c meant to become ONE IADD3.  There is no 3-input add in PTX; ptxas fuses two dependent adds, but the
// compiler, left alone, shares the partial sum S1 + ch between the two state updates of a round (one instruction
// fewer, one more on the dependent chain).  The empty asm makes the first operand opaque, so no subexpression is
// shared across two ss_add3 calls, and costs no instruction.
__device__ __forceinline__ uint32_t ss_add3(uint32_t a, uint32_t b, uint32_t c)
{
    asm("" : "+r"(a));
    return (a + b) + c;
}
__device__ __forceinline__ uint4 ss_ldg(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}


__constant__ uint32_t SS_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

// One piece per (start[i], len[i], stream[i], flags[i]); start multiple of 16.  Digest of a finished stream goes to
// row stream[i] of `out`; a piece with flag bit 0 (MKSNAP_R_MORE, len multiple of 64) parks the midstate in
// sstate[stream[i]] instead.  max_lanes: lanes of a pair that take pieces (32 normally; a tuning knob).
__global__ void __launch_bounds__(SS_THREADS)
k_sha256_streams(const uint8_t *__restrict__ data, const uint64_t *__restrict__ start, const uint64_t *__restrict__ len,
                 uint32_t n, const uint32_t *__restrict__ rng_stream, const uint32_t *__restrict__ rng_flags,
                 StreamState *__restrict__ sstate, uint8_t *__restrict__ out, uint32_t *__restrict__ work_counter,
                 uint32_t max_lanes, const uint32_t one /* = 1, see ss_add */,
                 uint64_t uni_len = 0, uint64_t uni_total = 0 /* start == nullptr: piece i = [i*uni_len, ...) of `data`, row i
                                                                   of `out` (Merkle levels: 8 KiB groups of a digest table) */)
{
    __shared__ SsShared sh;
    const uint32_t lane = threadIdx.x & 31;
    if ((threadIdx.x >> 5) != 0 && (threadIdx.x >> 5) != SS_SCHED_WARP)
        return;
#ifdef SS_PROFILE
    if (blockIdx.x == 0 && lane == 0) {
        uint32_t wid, sid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sid));
        ss_prof[(threadIdx.x >> 5) ? 6 : 5] = ((unsigned long long)sid << 32) | wid;
    }
#endif
    // named barriers: 1+b = "buffer b is full", 3+b = "buffer b is empty"
    if (threadIdx.x < 32) {
        // ------------------------------ rounds warp ------------------------------
        uint32_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#ifdef SS_PROFILE
        long long p_wait = 0, p_t0 = clock64();
#endif
        for (uint32_t it = 0;; ++it) {
            const uint32_t b = it & 1u;
#ifdef SS_PROFILE
            const long long p_a = clock64();
#endif
            ss_bar_sync(1 + b);
#ifdef SS_PROFILE
            p_wait += clock64() - p_a;
#endif
            const uint32_t c = sh.ctrl[b][lane];
            if (c & SS_EXIT)
                break; // warp-uniform
            const uint32_t sid = sh.sid[b][lane];
            if (c & SS_START_IV) {
                st[0] = 0x6a09e667; st[1] = 0xbb67ae85; st[2] = 0x3c6ef372; st[3] = 0xa54ff53a;
                st[4] = 0x510e527f; st[5] = 0x9b05688c; st[6] = 0x1f83d9ab; st[7] = 0x5be0cd19;
            }
            if (c & SS_START_RESUME) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    st[k] = sstate[sid].st[k];
            }
            uint32_t a = st[0], bb = st[1], cc = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
            // all 64 words of the block into registers before round 0: this warp is alone on its sub-partition, so
            // a shared-memory load issued next to its use (ptxas's choice when the loads are left to it) stalls the
            // round for the full ~30-cycle LDS latency -- 64 times per block (measured: 3250 cycles per block)
            uint32_t kw[64];
            {
                const uint32_t base = (uint32_t)__cvta_generic_to_shared(&sh.kw[b][lane * (SS_LANE_WORDS / 4)]);
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(kw[4 * q]), "=r"(kw[4 * q + 1]), "=r"(kw[4 * q + 2]), "=r"(kw[4 * q + 3])
                                 : "r"(base + 16u * q));
            }
#ifdef SS_PROFILE
            if (ss_dbg != 2)
#endif
#pragma unroll
            for (int t = 0; t < 64; ++t) {
                // Measured (k4_microbench, B200): a warp that is alone on its sub-partition issues one instruction per
                // TWO cycles, whichever pipe it goes to (2430 cycles per block for 1185 instructions; 1577 for the
                // schedule warp's ~800) -- so the per-stream rate is set by the INSTRUCTION COUNT of this loop, not by
                // pipe balance: 6 SHF + 4 LOP3 + four additions per round.  The e-chain per round is
                // SHF -> LOP3 -> IADD3; h + kw + d and S0 + mj + (h + kw) are ready before S1 is.
#if SS_ADD_MODE == 1
                const uint32_t x = ss_add(h, kw[t], one);
                const uint32_t y = ss_add(x, d, one);
                const uint32_t S1 = ss_rotr(e, 6) ^ ss_rotr(e, 11) ^ ss_rotr(e, 25);
                const uint32_t ch = (e & f) ^ (~e & g);
                const uint32_t S0 = ss_rotr(a, 2) ^ ss_rotr(a, 13) ^ ss_rotr(a, 22);
                const uint32_t mj = (a & bb) ^ (a & cc) ^ (bb & cc);
                const uint32_t z = ss_add(ss_add(S0, mj, one), x, one);
                const uint32_t chy = ss_add(ch, y, one), chz = ss_add(ch, z, one);
                h = g; g = f; f = e;
                e = ss_add(S1, chy, one);
                d = cc; cc = bb; bb = a;
                a = ss_add(S1, chz, one);
#else
                // four additions per round: with T1 = h + S1 + ch + kw the new state words are
                //   e' = d + T1 = S1 + ch + (h + kw + d)         a' = T1 + S0 + mj = e' + (S0 + mj - d)
                const uint32_t y = ss_add3(h, kw[t], d);
                const uint32_t S1 = ss_rotr(e, 6) ^ ss_rotr(e, 11) ^ ss_rotr(e, 25);
                const uint32_t ch = (e & f) ^ (~e & g);
                const uint32_t S0 = ss_rotr(a, 2) ^ ss_rotr(a, 13) ^ ss_rotr(a, 22);
                const uint32_t mj = (a & bb) ^ (a & cc) ^ (bb & cc);
                const uint32_t w_ = ss_add3(S0, mj, 0u - d);
                h = g; g = f; f = e;
                e = ss_add3(S1, ch, y);
                d = cc; cc = bb; bb = a;
                a = e + w_;
#endif
            }
            if (c & SS_ACTIVE) {
                st[0] += a; st[1] += bb; st[2] += cc; st[3] += d;
                st[4] += e; st[5] += f; st[6] += g; st[7] += h;
            }
            if (c & SS_FINAL) {
                uint4 o0, o1;
                o0.x = __byte_perm(st[0], 0, 0x0123); o0.y = __byte_perm(st[1], 0, 0x0123);
                o0.z = __byte_perm(st[2], 0, 0x0123); o0.w = __byte_perm(st[3], 0, 0x0123);
                o1.x = __byte_perm(st[4], 0, 0x0123); o1.y = __byte_perm(st[5], 0, 0x0123);
                o1.z = __byte_perm(st[6], 0, 0x0123); o1.w = __byte_perm(st[7], 0, 0x0123);
                uint4 *dst = reinterpret_cast<uint4 *>(out + (uint64_t)sid * 32);
                dst[0] = o0;
                dst[1] = o1;
            }
            if (c & SS_PARK) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    sstate[sid].st[k] = st[k];
            }
            ss_bar_arrive(3 + b);
        }
#ifdef SS_PROFILE
        if (blockIdx.x == 0 && lane == 0) {
            ss_prof[0] = (unsigned long long)(clock64() - p_t0);
            ss_prof[1] = (unsigned long long)p_wait;
        }
#endif
        return;
    }

    // ------------------------------ schedule warp ------------------------------
    const uint8_t *p = nullptr; // next block to take
    uint64_t total = 0, done = 0, prior = 0;
    uint32_t sid = 0;
    bool more = false, first = false;
    uint32_t phase = 0; // 0 idle, 1 data blocks, 2 needs the extra length block
    bool exhausted = lane >= max_lanes;
    uint4 pre[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};

#ifdef SS_PROFILE
    long long p_wait = 0, p_t0 = clock64();
#endif
    for (uint32_t it = 0;; ++it) {
        const uint32_t b = it & 1u;
        // ---- refill idle lanes (warp-aggregated fetch) ----
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, phase == 0 && !exhausted);
        if (need) {
            uint32_t basei = 0;
            const uint32_t leader = __ffs(need) - 1;
            if (lane == leader)
                basei = atomicAdd(work_counter, __popc(need));
            basei = __shfl_sync(0xFFFFFFFFu, basei, leader);
            if (phase == 0 && !exhausted) {
                const uint32_t idx = basei + __popc(need & ((1u << lane) - 1u));
                if (idx < n) {
                    if (start) {
                        p = data + start[idx];
                        total = len[idx];
                        sid = rng_stream[idx];
                        more = (rng_flags[idx] & 1u) != 0;
                    } else {
                        const uint64_t o = (uint64_t)idx * uni_len;
                        p = data + o;
                        total = uni_total - o < uni_len ? uni_total - o : uni_len;
                        sid = idx;
                        more = false;
                    }
                    done = 0;
                    prior = 0;
                    first = true;
                    phase = 1;
                    if (start && sstate[sid].open)
                        prior = sstate[sid].bytes | (1ull << 63); // bit 63: resume the parked midstate
#ifdef SS_PROFILE
                    if (ss_dbg != 1)
#endif
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((uint64_t)(16 * k) < total)
                            pre[k] = ss_ldg(reinterpret_cast<const uint4 *>(p) + k);
                } else {
                    exhausted = true;
                }
            }
        }
        const bool any = __any_sync(0xFFFFFFFFu, phase != 0);
#ifdef SS_PROFILE
        const long long p_a = clock64();
#endif
        if (it >= 2)
            ss_bar_sync(3 + b); // the rounds warp is done with buffer b
#ifdef SS_PROFILE
        p_wait += clock64() - p_a;
#endif
        if (!any) {
            sh.ctrl[b][lane] = SS_EXIT;
            ss_bar_arrive(1 + b);
#ifdef SS_PROFILE
            if (blockIdx.x == 0 && lane == 0) {
                ss_prof[2] = (unsigned long long)(clock64() - p_t0);
                ss_prof[3] = (unsigned long long)p_wait;
                ss_prof[4] = it;
            }
#endif
            break;
        }

        // ---- this lane's block -> w[0..15] ----
        uint32_t w[16];
        uint32_t ctrl = 0;
        const bool resume = (prior >> 63) != 0;
        const uint64_t prior_b = prior & ~(1ull << 63);
        if (phase != 0) {
            ctrl = SS_ACTIVE;
            if (first)
                ctrl |= resume ? SS_START_RESUME : SS_START_IV;
            first = false;
        }
        const uint64_t rem = total - done;
        if (phase == 1) {
            w[0] = pre[0].x; w[1] = pre[0].y; w[2] = pre[0].z; w[3] = pre[0].w;
            w[4] = pre[1].x; w[5] = pre[1].y; w[6] = pre[1].z; w[7] = pre[1].w;
            w[8] = pre[2].x; w[9] = pre[2].y; w[10] = pre[2].z; w[11] = pre[2].w;
            w[12] = pre[3].x; w[13] = pre[3].y; w[14] = pre[3].z; w[15] = pre[3].w;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                w[i] = __byte_perm(w[i], 0, 0x0123);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                w[i] = 0;
        }
        const bool tail = (phase == 1 && rem < 64) || phase == 2;
        if (__any_sync(0xFFFFFFFFu, tail)) { // rare: some lane is at the end of its stream
            if (phase == 1 && rem < 64) {
                // final data block: keep `rem` bytes, append 0x80, zero the rest
                const uint32_t rb = (uint32_t)rem;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int32_t k = (int32_t)rb - 4 * i; // valid bytes in this word
                    const uint32_t keep = k >= 4 ? 0xFFFFFFFFu : (k <= 0 ? 0u : ~(0xFFFFFFFFu >> (8 * k)));
                    uint32_t v = w[i] & keep;
                    if (k >= 0 && k < 4)
                        v |= 0x80000000u >> (8 * k);
                    w[i] = v;
                }
                if (rb < 56) {
                    w[14] = (uint32_t)(((prior_b + total) * 8) >> 32);
                    w[15] = (uint32_t)((prior_b + total) * 8);
                    ctrl |= SS_FINAL;
                    phase = 0;
                } else {
                    phase = 2;
                }
            } else if (phase == 2) {
                w[14] = (uint32_t)(((prior_b + total) * 8) >> 32);
                w[15] = (uint32_t)((prior_b + total) * 8);
                ctrl |= SS_FINAL;
                phase = 0;
            }
            if ((ctrl & SS_FINAL) && start)
                sstate[sid].open = 0;
        }
        if (phase == 1) { // a full data block was taken (rem >= 64)
            p += 64;
            done += 64;
            if (done == total && more) { // piece boundary of a stream that continues: park the midstate
                ctrl |= SS_PARK;
                sstate[sid].bytes = prior_b + total;
                sstate[sid].open = 1;
                phase = 0;
            } else {
                // prefetch the next block while this one is expanded (the only global-memory latency on the path)
                const uint64_t left = total - done;
#ifdef SS_PROFILE
                if (ss_dbg != 1)
#endif
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((uint64_t)(16 * k) < left)
                        pre[k] = ss_ldg(reinterpret_cast<const uint4 *>(p) + k);
            }
        }
        sh.ctrl[b][lane] = ctrl;
        sh.sid[b][lane] = sid;

        // ---- message schedule: W[t] + K[t] for the 64 rounds ----
        // Rolled: 16 words are copied, then 3 x 16 are expanded by ONE unrolled body (the ring index t & 15 is static
        // inside it, K[t] comes from constant memory).  Fully unrolled, this loop was 11 KB of code running beside the
        // rounds warp's 15 KB and the rounds warp -- the one that bounds a stream -- needed 2219 cycles per block; with
        // the rolled loop it needs 2022 although this warp itself got slower (1484 -> 1863 cycles alone, it has the
        // slack): the two warps compete for instruction fetch, not for a pipe (profiles/r2b_k4_microbench.txt).
        const uint32_t kbase = (uint32_t)__cvta_generic_to_shared(&sh.kw[b][lane * (SS_LANE_WORDS / 4)]);
#ifdef SS_PROFILE
        if (ss_dbg != 3)
#endif
        {
            uint32_t o4[4];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                o4[t & 3] = w[t] + SS_K[t];
                if ((t & 3) == 3)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(kbase + 4u * (t - 3)), "r"(o4[0]), "r"(o4[1]), "r"(o4[2]),
                                 "r"(o4[3])
                                 : "memory");
            }
#pragma unroll 1
            for (int q = 1; q < 4; ++q) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                    const uint32_t s0 = ss_rotr(w15, 7) ^ ss_rotr(w15, 18) ^ (w15 >> 3);
                    const uint32_t s1 = ss_rotr(w2, 17) ^ ss_rotr(w2, 19) ^ (w2 >> 10);
#if SS_ADD_MODE == 1
                    const uint32_t wt = ss_add(ss_add(w[i], s0, one), ss_add(w[(i + 9) & 15], s1, one), one);
                    o4[i & 3] = ss_add(wt, SS_K[16 * q + i], one);
#else
                    const uint32_t wt = ss_add3(ss_add3(w[i], s0, w[(i + 9) & 15]), s1, 0u);
                    o4[i & 3] = wt + SS_K[16 * q + i];
#endif
                    w[i] = wt;
                    if ((i & 3) == 3)
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(kbase + 4u * (16 * q + i - 3)), "r"(o4[0]), "r"(o4[1]),
                                     "r"(o4[2]), "r"(o4[3])
                                     : "memory");
                }
            }
        }
        ss_bar_arrive(1 + b);
    }
}

} // namespace mk
