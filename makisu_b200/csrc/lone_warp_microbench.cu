// lone_warp_microbench.cu -- what ONE warp, alone on its SM sub-partition, can issue per cycle (B200, sm_100a).
// The serial-stream SHA-256 kernel (mksnap_sha_stream.cuh) is bounded by exactly this: its rounds warp has no other
// warp to share the scheduler with.  Each variant runs ITERS x 16 instructions on 8 independent accumulators
// (dependent distance 8 > the 4-cycle ALU latency), one warp per CTA, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lw makisu_b200/csrc/lone_warp_microbench.cu && /tmp/lw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 8192

template <int OP> __global__ void __launch_bounds__(32, 1) bench(uint32_t *out, uint32_t seed, uint32_t one, unsigned long long *cycles)
{
    uint32_t a[8], b = seed | 1u, c = seed * 7u + 3u;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i * 0x9E3779B9u + threadIdx.x;
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));                                  // ALU only
                if (OP == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one), "r"(c));                  // FMA only
                if (OP == 2) { if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one), "r"(c));     // alternate ALU / FMA
                               else asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i])); }
                if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));                // ALU only (LOP3)
                if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));                                   // whatever ptxas picks for add
                if (OP == 5) { if (i & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));                      // alternate SHF / add
                               else asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i])); }
                if (OP == 6) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[0]));                                   // ONE dependent chain
                if (OP == 7) { if (i % 3 == 0) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[0]));                 // SHF -> LOP3 -> add chain
                               else if (i % 3 == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[0]) : "r"(b), "r"(c));
                               else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[0]) : "r"(b)); }
                if (OP == 8) { if (i % 3 == 0) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[0]));                 // SHF -> LOP3 -> IMAD chain
                               else if (i % 3 == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[0]) : "r"(b), "r"(c));
                               else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[0]) : "r"(one), "r"(c)); }
            }
    }
    unsigned long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i];
    out[blockIdx.x * 32 + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, int threads = 32)
{
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 8);
    bench<OP><<<148, threads>>>(out, 12345, 1, cyc);
    cudaDeviceSynchronize();
    bench<OP><<<148, threads>>>(out, 12345, 1, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    printf("%-44s %6.3f cycles per instruction (one warp per SM)  (%s)\n", name, avg / (16.0 * ITERS), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

// ---- one SHA-256 round chain per lane, rotates as SHF (ALU pipe) or as IMAD.WIDE by 2^(32-n) (FMA pipe: hi ^ lo of the
// 64-bit product IS the rotation, the two halves have disjoint bits) ----
__device__ __forceinline__ uint32_t rot_shf(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ void rot_mul(uint32_t x, int n, uint32_t &lo, uint32_t &hi)
{
    asm("{ .reg .b64 p; mul.wide.u32 p, %2, %3; mov.b64 {%0, %1}, p; }" : "=r"(lo), "=r"(hi) : "r"(x), "r"(1u << (32 - n)));
}
__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b, uint32_t c) { asm("" : "+r"(a)); return (a + b) + c; }
__device__ __forceinline__ uint32_t addsub(uint32_t a, uint32_t b, uint32_t c) { asm("" : "+r"(a)); return (a + b) - c; }
// NMUL = how many of the three rotates of each big sigma go through the multiplier
template <int NMUL> __device__ __forceinline__ uint32_t big_sigma(uint32_t x, int r0, int r1, int r2)
{
    if (NMUL == 0) return rot_shf(x, r0) ^ rot_shf(x, r1) ^ rot_shf(x, r2);
    uint32_t l0, h0, l1, h1, l2, h2;
    if (NMUL == 1) { rot_mul(x, r2, l2, h2); return (rot_shf(x, r0) ^ rot_shf(x, r1)) ^ (l2 ^ h2); }
    if (NMUL == 2) { rot_mul(x, r1, l1, h1); rot_mul(x, r2, l2, h2); return (rot_shf(x, r0) ^ l1 ^ h1) ^ (l2 ^ h2); }
    rot_mul(x, r0, l0, h0); rot_mul(x, r1, l1, h1); rot_mul(x, r2, l2, h2);
    return (l0 ^ h0 ^ l1) ^ (h1 ^ l2 ^ h2);
}
template <int N1, int N0> __global__ void __launch_bounds__(32, 1) bench_round(uint32_t *out, uint32_t seed, unsigned long long *cycles)
{
    uint32_t a = seed + threadIdx.x, b = a * 3u, c = a * 5u, d = a * 7u, e = a * 11u, f = a * 13u, g = a * 17u, h = a * 19u;
    uint32_t kw[16];
#pragma unroll
    for (int i = 0; i < 16; i++) kw[i] = seed * (2 * i + 1);
    unsigned long long t0 = clock64();
    for (int it = 0; it < 512; it++) {
#pragma unroll
        for (int t = 0; t < 64; t++) {
            const uint32_t y = add3(h, kw[t & 15], d);
            const uint32_t S1 = big_sigma<N1>(e, 6, 11, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t S0 = big_sigma<N0>(a, 2, 13, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t w_ = addsub(S0, mj, d);
            h = g; g = f; f = e; e = add3(S1, ch, y);
            d = c; c = b; b = a; a = e + w_;
        }
    }
    unsigned long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int N1, int N0> void run_round(const char *name, int threads = 32)
{
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 8);
    bench_round<N1, N0><<<148, threads>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    bench_round<N1, N0><<<148, threads>>>(out, 12345, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    uint32_t o0; cudaMemcpy(&o0, out, 4, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    printf("%-44s %6.2f cycles per round, %6.0f per 64-byte block  out %08x (%s)\n", name, avg / (64.0 * 512), avg / 512, o0, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

// ---- does a second warp with its OWN large straight-line body (the schedule warp) slow the rounds warp down?
// warp 0: 64 unrolled rounds (ADDM 0: IADD3-friendly adds, 1: every add as IMAD on the FMA pipe);
// warp 1 (if `second`): 3 x 16 unrolled message-schedule words per iteration, as the real schedule warp does.
__device__ __forceinline__ uint32_t madd(uint32_t a, uint32_t b, uint32_t one)
{
    uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b)); return d;
}
template <int ADDM, int UNROLL> __global__ void __launch_bounds__(64, 1)
bench_pair(uint32_t *out, uint32_t seed, uint32_t one, int second, unsigned long long *cycles)
{
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) {
        if (!second) return;
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = seed * (2 * i + 3) + lane;
        uint32_t acc = 0;
        for (int it = 0; it < 512; it++) {
#pragma unroll
            for (int t = 0; t < 64; t++) {
                const uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
                const uint32_t s0 = rot_shf(w15, 7) ^ rot_shf(w15, 18) ^ (w15 >> 3);
                const uint32_t s1 = rot_shf(w2, 17) ^ rot_shf(w2, 19) ^ (w2 >> 10);
                w[t & 15] = add3(w[t & 15], s0, w[(t + 9) & 15]) + s1;
                acc += w[t & 15] + t * 0x9E3779B9u;
            }
        }
        out[blockIdx.x * 64 + threadIdx.x] = acc;
        return;
    }
    uint32_t a = seed + lane, b = a * 3u, c = a * 5u, d = a * 7u, e = a * 11u, f = a * 13u, g = a * 17u, h = a * 19u;
    uint32_t kw[16];
#pragma unroll
    for (int i = 0; i < 16; i++) kw[i] = seed * (2 * i + 1);
    unsigned long long t0 = clock64();
    for (int it = 0; it < 512 * (64 / UNROLL); it++) {
#pragma unroll
        for (int t = 0; t < UNROLL; t++) {
            const uint32_t S1 = big_sigma<0>(e, 6, 11, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t S0 = big_sigma<0>(a, 2, 13, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            if (ADDM == 0) {
                const uint32_t y = add3(h, kw[t & 15], d);
                const uint32_t w_ = addsub(S0, mj, d);
                h = g; g = f; f = e; e = add3(S1, ch, y);
                d = c; c = b; b = a; a = e + w_;
            } else {
                const uint32_t y = madd(madd(h, kw[t & 15], one), d, one);
                const uint32_t w_ = madd(madd(S0, mj, one), 0u - d, one);
                h = g; g = f; f = e; e = madd(madd(ch, y, one), S1, one);
                d = c; c = b; b = a; a = madd(e, w_, one);
            }
        }
    }
    unsigned long long t1 = clock64();
    out[blockIdx.x * 64 + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int ADDM, int UNROLL> void run_pair(const char *name, int second)
{
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, 148 * 64 * 4); cudaMalloc(&cyc, 148 * 8);
    bench_pair<ADDM, UNROLL><<<148, 64>>>(out, 12345, 1, second, cyc);
    cudaDeviceSynchronize();
    bench_pair<ADDM, UNROLL><<<148, 64>>>(out, 12345, 1, second, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    uint32_t o0; cudaMemcpy(&o0, out, 4, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    printf("%-58s %6.2f cycles per round, %6.0f per block  out %08x (%s)\n", name, avg / (64.0 * 512), avg / 512, o0, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}


// ---- the split-rounds step of mksnap_sha_stream.cuh MODE 1 (two lanes per stream) on registers only.
// XCH 0: no exchange (the word a lane needs is faked from its own), 1: one butterfly shuffle per step, consumed DELAY steps later
template <int XCH, int DELAY> __global__ void __launch_bounds__(32, 1) bench_split(uint32_t *out, uint32_t seed, unsigned long long *cycles)
{
    const uint32_t lane = threadIdx.x & 31, half = lane >> 4;
    uint32_t n1 = half ? 2u : 6u, n2 = half ? 13u : 11u, n3 = half ? 22u : 25u, nmask = half ? 0xFFFFFFFFu : 0u, sg = half ? 0xFFFFFFFFu : 1u;
    asm volatile("" : "+r"(n1), "+r"(n2), "+r"(n3), "+r"(nmask), "+r"(sg));
    uint32_t x0 = seed + lane, x1 = x0 * 3u, x2 = x0 * 5u, x3 = x0 * 7u;
    uint32_t kw[16];
#pragma unroll
    for (int i = 0; i < 16; i++) kw[i] = seed * (2 * i + 1);
    uint32_t r[4] = {x3, x2, x1, x0};
    unsigned long long t0 = clock64();
    for (int it = 0; it < 512; it++) {
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            uint32_t r1, r2, r3, S, p, F, u;
            asm("shf.r.wrap.b32 %0, %1, %1, %2;" : "=r"(r1) : "r"(x0), "r"(n1));
            asm("shf.r.wrap.b32 %0, %1, %1, %2;" : "=r"(r2) : "r"(x0), "r"(n2));
            asm("shf.r.wrap.b32 %0, %1, %1, %2;" : "=r"(r3) : "r"(x0), "r"(n3));
            asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(S) : "r"(r1), "r"(r2), "r"(r3));
            asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(p) : "r"(x0), "r"(x1), "r"(nmask));
            asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(F) : "r"(p), "r"(x1), "r"(x2));
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(x3), "r"(sg), "r"(kw[k & 15]));
            const uint32_t Y = u + r[0];
            uint32_t a = S;
            asm("" : "+r"(a));
            const uint32_t nx = (a + F) + Y;
#pragma unroll
            for (int j = 0; j < DELAY - 1; ++j) r[j] = r[j + 1];
            r[DELAY - 1] = XCH ? __shfl_xor_sync(0xFFFFFFFFu, nx, 16) : (nx ^ nmask);
            x3 = x2; x2 = x1; x1 = x0; x0 = nx;
        }
    }
    unsigned long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int XCH, int DELAY> void run_split(const char *name)
{
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 8);
    bench_split<XCH, DELAY><<<148, 32>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    bench_split<XCH, DELAY><<<148, 32>>>(out, 12345, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    printf("%-64s %6.2f cycles per step, %6.0f per 66-step block (%s)\n", name, avg / (64.0 * 512), avg / 512 / 64 * 66, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char **argv)
{
    if (argc > 1 && argv[1][0] == 's') {
        run_split<0, 2>("split step, no exchange (own word, 2 steps old)");
        run_split<1, 2>("split step, SHFL.BFLY per step, consumed 2 steps later");
        run_split<1, 3>("split step, SHFL.BFLY per step, consumed 3 steps later");
        run_split<1, 4>("split step, SHFL.BFLY per step, consumed 4 steps later");
        run_split<1, 1>("split step, SHFL.BFLY per step, consumed in the NEXT step");
        return 0;
    }
    if (argc > 1) { // does a warp with fewer ACTIVE lanes issue faster?  (it does not, if the pipe takes two passes per warp anyway)
        run<0>("SHF only, 32 lanes", 32);
        run<0>("SHF only, 16 lanes (half warp launched)", 16);
        run<0>("SHF only, 8 lanes", 8);
        run<0>("SHF only, 1 lane", 1);
        run<1>("IMAD only, 16 lanes", 16);
        run<2>("SHF, IMAD alternating, 16 lanes", 16);
        run<4>("add.u32 only, 16 lanes", 16);
        run_round<0, 0>("SHA round, 6 SHF rotates, 32 lanes", 32);
        run_round<0, 0>("SHA round, 6 SHF rotates, 16 lanes", 16);
        run_round<0, 0>("SHA round, 6 SHF rotates, 1 lane", 1);
        return 0;
    }
    run_pair<0, 64>("rounds warp alone, IADD3 adds, 64 rounds unrolled", 0);
    run_pair<0, 64>("  + schedule-like warp on the next sub-partition", 1);
    run_pair<1, 64>("rounds warp alone, IMAD adds, 64 rounds unrolled", 0);
    run_pair<1, 64>("  + schedule-like warp", 1);
    run_pair<0, 8>("rounds warp alone, IADD3 adds, 8 rounds unrolled", 0);
    run_pair<0, 8>("  + schedule-like warp", 1);
    run_pair<1, 8>("rounds warp alone, IMAD adds, 8 rounds unrolled", 0);
    run_pair<1, 8>("  + schedule-like warp", 1);

    run_round<0, 0>("SHA round, 6 SHF rotates");
    run_round<1, 1>("SHA round, 4 SHF + 2 IMAD.WIDE rotates");
    run_round<2, 1>("SHA round, 3 SHF + 3 IMAD.WIDE rotates");
    run_round<2, 2>("SHA round, 2 SHF + 4 IMAD.WIDE rotates");
    run_round<3, 2>("SHA round, 1 SHF + 5 IMAD.WIDE rotates");
    run_round<3, 3>("SHA round, 6 IMAD.WIDE rotates");
    run_round<1, 0>("SHA round, S1 one IMAD.WIDE, S0 all SHF");
    run_round<2, 0>("SHA round, S1 two IMAD.WIDE, S0 all SHF");

    run<0>("SHF only (ALU pipe)");
    run<3>("LOP3 only (ALU pipe)");
    run<1>("IMAD only (FMA pipe)");
    run<2>("SHF, IMAD alternating (ALU + FMA)");
    run<4>("add.u32 only (IADD3 / IMAD.IADD, ptxas's pick)");
    run<5>("SHF, add.u32 alternating");
    run<6>("SHF, ONE dependent chain (latency)");
    run<7>("SHF -> LOP3 -> add, one dependent chain");
    run<8>("SHF -> LOP3 -> IMAD, one dependent chain");
    return 0;
}
