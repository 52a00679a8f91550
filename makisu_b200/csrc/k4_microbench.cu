// k4_microbench.cu -- measuring stick for the serial-stream SHA-256 kernel (mksnap_sha_stream.cuh), outside the
// engine: N independent streams of L bytes of device-generated content, digests checked against a scalar host
// SHA-256 (test tool only, not part of libmksnap), per-stream and aggregate MB/s printed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/k4 makisu_b200/csrc/k4_microbench.cu && /tmp/k4
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define SS_PROFILE 1
#include "mksnap_sha_stream.cuh"

using namespace mk;

static const uint32_t HK[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void host_sha256(const uint8_t *m, uint64_t n, uint8_t out[32])
{
    uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::vector<uint8_t> buf(m, m + n);
    buf.push_back(0x80);
    while (buf.size() % 64 != 56)
        buf.push_back(0);
    for (int i = 7; i >= 0; --i)
        buf.push_back((uint8_t)((n * 8) >> (8 * i)));
    for (size_t o = 0; o < buf.size(); o += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; ++i)
            w[i] = (buf[o + 4 * i] << 24) | (buf[o + 4 * i + 1] << 16) | (buf[o + 4 * i + 2] << 8) | buf[o + 4 * i + 3];
        for (int i = 16; i < 64; ++i)
            w[i] = w[i - 16] + (rr(w[i - 15], 7) ^ rr(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] +
                   (rr(w[i - 2], 17) ^ rr(w[i - 2], 19) ^ (w[i - 2] >> 10));
        uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
        for (int i = 0; i < 64; ++i) {
            uint32_t t1 = h + (rr(e, 6) ^ rr(e, 11) ^ rr(e, 25)) + ((e & f) ^ (~e & g)) + HK[i] + w[i];
            uint32_t t2 = (rr(a, 2) ^ rr(a, 13) ^ rr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
    }
    for (int i = 0; i < 8; ++i)
        for (int k = 0; k < 4; ++k)
            out[4 * i + k] = (uint8_t)(st[i] >> (24 - 8 * k));
}

__global__ void fill(uint64_t *d, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        d[i] = z ^ (z >> 31);
    }
}

#define CK(x)                                                                                \
    do {                                                                                     \
        cudaError_t e = (x);                                                                 \
        if (e != cudaSuccess) {                                                              \
            printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));                 \
            exit(1);                                                                         \
        }                                                                                    \
    } while (0)

int main(int argc, char **argv)
{
    const uint64_t total_cap = 1ull << 30;
    uint8_t *d_data;
    CK(cudaMalloc(&d_data, total_cap + 8192));
    fill<<<148 * 8, 256>>>((uint64_t *)d_data, (total_cap + 8192) / 8);
    CK(cudaDeviceSynchronize());
    int sm = 148;
    const int ns[] = {1, 32, 256, 1024, 4736, 9472, 20000};
    const int lanes[] = {32};
    printf("schedule warp = warp %d of the CTA (%d threads)\n", SS_SCHED_WARP, SS_THREADS);
    for (int dbg = 0; dbg < 4; ++dbg)
    for (int max_lanes : lanes)
        for (int n : ns) {
            if (dbg && n != 32) continue;
            CK(cudaMemcpyToSymbol(mk::ss_dbg, &dbg, sizeof dbg));
            if (dbg) printf("debug mode %d (1 = no global loads in the schedule warp, 2 = no rounds in the rounds warp, 3 = no expansion in the schedule warp): digests will not match\n", dbg);
            // ragged lengths around L, L chosen so the run takes a few ms
            uint64_t L = n <= 64 ? (1u << 20) : (n <= 1024 ? (256u << 10) : (48u << 10));
            std::vector<uint64_t> start(n), len(n);
            std::vector<uint32_t> sid(n), flags(n, 0);
            uint64_t off = 0;
            for (int i = 0; i < n; ++i) {
                len[i] = L - (uint64_t)(i * 37 % 1000);
                start[i] = off;
                off += (len[i] + 511) / 512 * 512;
                sid[i] = i;
            }
            if (off > total_cap) {
                printf("skip n=%d\n", n);
                continue;
            }
            uint64_t *d_start, *d_len;
            uint32_t *d_sid, *d_flags, *d_work;
            StreamState *d_ss;
            uint8_t *d_out;
            CK(cudaMalloc(&d_start, n * 8)); CK(cudaMalloc(&d_len, n * 8)); CK(cudaMalloc(&d_sid, n * 4));
            CK(cudaMalloc(&d_flags, n * 4)); CK(cudaMalloc(&d_work, 4)); CK(cudaMalloc(&d_ss, n * sizeof(StreamState)));
            CK(cudaMalloc(&d_out, n * 32));
            CK(cudaMemcpy(d_start, start.data(), n * 8, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_len, len.data(), n * 8, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_sid, sid.data(), n * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_flags, flags.data(), n * 4, cudaMemcpyHostToDevice));
            CK(cudaMemset(d_ss, 0, n * sizeof(StreamState)));
            const int pairs = (n + max_lanes - 1) / max_lanes;
            const int grid = pairs < sm * 4 ? pairs : sm * 4;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemset(d_work, 0, 4));
                CK(cudaEventRecord(e0));
                k_sha256_streams<<<grid, SS_THREADS>>>(d_data, d_start, d_len, (uint32_t)n, d_sid, d_flags, d_ss, d_out, d_work,
                                                       (uint32_t)max_lanes, 1u);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            // verify a few
            std::vector<uint8_t> out(n * 32);
            CK(cudaMemcpy(out.data(), d_out, n * 32, cudaMemcpyDeviceToHost));
            int bad = 0;
            const int checks[] = {0, n / 2, n - 1};
            for (int ci = 0; ci < 3; ++ci) {
                const int i = checks[ci];
                std::vector<uint8_t> m(len[i]);
                CK(cudaMemcpy(m.data(), d_data + start[i], len[i], cudaMemcpyDeviceToHost));
                uint8_t ref[32];
                host_sha256(m.data(), len[i], ref);
                if (memcmp(ref, out.data() + 32 * i, 32)) bad++;
            }
            unsigned long long prof[8];
            CK(cudaMemcpyFromSymbol(prof, mk::ss_prof, sizeof prof));
            if (n <= 64 || n == 9472)
                printf("    CTA0: rounds warp %.0f cyc/block (%.0f waiting), schedule warp %.0f cyc/block (%.0f waiting), %llu blocks, %.0f MHz; "
                       "hw warp slots %llu / %llu (sm %llu)\n",
                       (double)prof[0] / prof[4], (double)prof[1] / prof[4], (double)prof[2] / prof[4], (double)prof[3] / prof[4], prof[4],
                       (double)prof[0] / best / 1e3, prof[5] & 0xffffffffu, prof[6] & 0xffffffffu, prof[5] >> 32);
            double bytes = 0;
            for (int i = 0; i < n; ++i) bytes += len[i];
            printf("lanes/pair %2d  streams %6d  grid %4d  L %8llu  %8.3f ms  per-stream %7.1f MB/s  aggregate %8.2f GB/s  %s\n",
                   max_lanes, n, grid, (unsigned long long)L, best, (double)L / best / 1e3, bytes / best / 1e6, bad ? "MISMATCH" : "ok");
            cudaFree(d_start); cudaFree(d_len); cudaFree(d_sid); cudaFree(d_flags); cudaFree(d_work); cudaFree(d_ss); cudaFree(d_out);
        }
    return 0;
}
