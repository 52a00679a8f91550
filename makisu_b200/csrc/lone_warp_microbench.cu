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

template <int OP> void run(const char *name)
{
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 8);
    bench<OP><<<148, 32>>>(out, 12345, 1, cyc);
    cudaDeviceSynchronize();
    bench<OP><<<148, 32>>>(out, 12345, 1, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    printf("%-44s %6.3f cycles per instruction (one warp per SM)  (%s)\n", name, avg / (16.0 * ITERS), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
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
