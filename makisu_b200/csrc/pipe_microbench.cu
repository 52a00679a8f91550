// pipe_microbench.cu -- measured per-SM instruction throughput of the integer ops the hash kernels are
// built from (B200, sm_100a).  Not part of libmksnap: a measuring stick for DESIGN.md section 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pm makisu_b200/csrc/pipe_microbench.cu && /tmp/pm
// Each kernel runs ITERS x UNR independent ops per thread on 8 accumulators (no dependent chain shorter
// than 8 ops); result = warp-instructions / cycle / SM with 32 resident warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNR 8

template <int OP> __global__ void __launch_bounds__(1024, 1) bench(uint32_t *out, uint32_t seed, unsigned long long *cycles)
{
    uint32_t a[UNR], b = seed | 1u, c = seed * 7u + 3u;
#pragma unroll
    for (int i = 0; i < UNR; i++) a[i] = seed + i * 0x9E3779B9u + threadIdx.x;
    __shared__ uint32_t tab[256 * 32];
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) tab[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(tab) + (threadIdx.x & 31) * 4;
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNR; i++) {
            if (OP == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 2) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
            if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, 0x0123;" : "+r"(a[i]) : "r"(b));
            if (OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 5) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, 0x02000000;" : "=l"(w) : "r"(a[i])); a[i] = (uint32_t)w ^ (uint32_t)(w >> 32); }
            if (OP == 6) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, 0x02000000;" : "=l"(w) : "r"(a[i])); a[i] = (uint32_t)(w >> 32); }
            if (OP == 7) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 8) { uint32_t v, ad; asm volatile("prmt.b32 %0, %1, 0, 0x4440;" : "=r"(ad) : "r"(a[i])); asm volatile("mad.lo.u32 %0, %0, 128, %1;" : "+r"(ad) : "r"(tbase)); asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ad)); asm volatile("mad.lo.u32 %0, %0, 2, %1;" : "+r"(a[i]) : "r"(v)); }
            if (OP == 9) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[(i + 4) % UNR]) : "r"(b), "r"(c)); }
            if (OP == 10) asm volatile("mad.lo.u32 %0, %0, 2, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 11) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 12) asm volatile("shl.b32 %0, %0, 3;" : "+r"(a[i]));
            if (OP == 13) asm volatile("shr.u32 %0, %0, 3;" : "+r"(a[i]));
            if (OP == 14) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tbase + ((i * 37 + it) & 255) * 128)); a[i] ^= v; }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < UNR; i++) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, int ops_per_iter)
{
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    bench<OP><<<148, 1024>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    bench<OP><<<148, 1024>>>(out, 12345, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    double winst = 32.0 * ITERS * UNR * ops_per_iter;
    printf("%-34s %7.3f warp-inst/clk/SM  (%s)\n", name, winst / avg, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("IADD (add.u32)", 1);
    run<1>("LOP3 xor3", 1);
    run<2>("SHF.R.W rotate", 1);
    run<3>("PRMT", 1);
    run<4>("IMAD (mad.lo)", 1);
    run<5>("IMAD.WIDE + LOP3", 2);
    run<6>("IMAD.WIDE (hi only)", 1);
    run<7>("VIMNMX (min.u32)", 1);
    run<8>("gear step PRMT+IMAD+LDS+IMAD", 4);
    run<9>("IADD + IMAD pair", 2);
    run<10>("IMAD h*2+g", 1);
    run<11>("IMAD.HI (mul.hi)", 1);
    run<12>("SHL imm", 1);
    run<13>("SHR imm", 1);
    run<14>("LDS only (conflict-free)", 2);
    return 0;
}
