// shuffle_cta_probe.cu -- stand-alone check + timing of the CTA-parallel exact Fisher-Yates replay
// (genomic_b200/csrc/shuffle.cuh) against the sequential loop of CBS.cpp:487-493 run on the host.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -o tools/shuffle_cta_probe tools/shuffle_cta_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <random>
#include <vector>
#include <cuda_runtime.h>

#include "../genomic_b200/csrc/cbs_core.h"
#include "../genomic_b200/csrc/shuffle.cuh"

using namespace cbsg;

template <int T, int K, bool GLOBAL>
__global__ void __launch_bounds__(T) k_probe(const uint64_t* words, const double* cur, double* out, unsigned* lastg, int n, int P,
                                            int hbits, unsigned* ctr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* claim = (unsigned*)smem_raw;
    const int H = 1 << hbits;
    for (int k = threadIdx.x; k < H; k += T) claim[k] = 0;
    __shared__ int s_g;
    unsigned epoch = 0;
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) s_g = (int)atomicAdd(ctr, 1u);
        __syncthreads();
        const int p = s_g;
        __syncthreads();
        if (p >= P) break;
        ShufDraws src;
        src.win = words + (size_t)p * n; src.mt = true;
        double* sx = out + (size_t)p * (n + 1);
        if (GLOBAL) shuffle_cta<T, K>(LastGlobal32{lastg + (size_t)blockIdx.x * (n + 1)}, claim, H - 1, epoch, n, src, cur, nullptr, sx);
        else shuffle_cta<T, K>(LastSmem16{(unsigned short*)(smem_raw + (size_t)H * 4)}, claim, H - 1, epoch, n, src, cur, nullptr, sx);
    }
}

template <int T, int K, bool GLOBAL>
float run(const uint64_t* dw, const double* dcur, double* dout, unsigned* dlast, int n, int P, int hbits, int ctas_per_sm, unsigned* dctr) {
    size_t smem = ((size_t)4 << hbits) + (GLOBAL ? 0 : 2 * (size_t)(n + 2));
    cudaFuncSetAttribute(k_probe<T, K, GLOBAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_probe<T, K, GLOBAL>, T, smem);
    if (occ < 1) return -1.f;
    if (occ > ctas_per_sm) occ = ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(dctr, 0, 4);
        cudaEventRecord(e0);
        k_probe<T, K, GLOBAL><<<148 * occ, T, smem>>>(dw, dcur, dout, dlast, n, P, hbits, dctr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return -2.f; }
    return best;
}

int main(int argc, char** argv) {
    const int sizes[] = {37, 500, 3000, 9000, 20000, 33000, 53000, 65535, 91958, 147726};
    std::mt19937_64 rng(1);
    for (int n : sizes) {
        const int P = (n <= 9000) ? 4096 : (n <= 65535 ? 1184 : 592);
        std::vector<uint64_t> w((size_t)n * P);
        for (auto& v : w) v = rng();
        std::vector<double> cur(n);
        for (int k = 0; k < n; ++k) cur[k] = (double)k + 0.25;
        // host reference for a few permutations
        const int check[] = {0, 1, P / 2, P - 1};
        uint64_t *dw; double *dcur, *dout; unsigned *dlast, *dctr;
        cudaMalloc(&dw, w.size() * 8); cudaMalloc(&dcur, n * 8); cudaMalloc(&dout, (size_t)P * (n + 1) * 8);
        cudaMalloc(&dlast, (size_t)148 * 8 * (n + 1) * 4); cudaMalloc(&dctr, 4);
        cudaMemcpy(dw, w.data(), w.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dcur, cur.data(), n * 8, cudaMemcpyHostToDevice);
        auto verify = [&](const char* tag) {
            std::vector<double> got((size_t)n + 1);
            int bad = 0;
            for (int p : check) {
                std::vector<double> px(cur);
                for (int i = n; i >= 1; --i) {
                    const int j = draw_index(mt_temper(w[(size_t)p * n + (n - i)]), i);
                    std::swap(px[i - 1], px[j - 1]);
                }
                cudaMemcpy(got.data(), dout + (size_t)p * (n + 1), (n + 1) * 8, cudaMemcpyDeviceToHost);
                for (int k = 0; k < n; ++k) if (got[k + 1] != px[k]) { if (bad < 3) printf("  MISMATCH %s n=%d p=%d k=%d got %.2f want %.2f\n", tag, n, p, k, got[k + 1], px[k]); ++bad; }
            }
            return bad;
        };
        struct Cfg { int T, K, hbits, cps; bool glob; };
        const Cfg cfgs[] = {{128, 2, 11, 16, false}, {256, 2, 11, 8, false}, {256, 4, 12, 8, false}, {512, 2, 12, 4, false}, {512, 4, 13, 4, false}, {1024, 2, 13, 2, false},
                            {1024, 4, 13, 2, false}, {512, 4, 13, 4, true}, {1024, 4, 14, 2, true}, {1024, 4, 15, 1, true}, {512, 8, 14, 2, true}};
        for (const Cfg& c : cfgs) {
            if (!c.glob && n > 65535) continue;
            if (c.glob && n < 20000) continue;
            cudaMemset(dout, 0, (size_t)P * (n + 1) * 8);
            float ms = -1.f;
#define RUN(TT, KK) if (c.T == TT && c.K == KK) ms = c.glob ? run<TT, KK, true>(dw, dcur, dout, dlast, n, P, c.hbits, c.cps, dctr) : run<TT, KK, false>(dw, dcur, dout, dlast, n, P, c.hbits, c.cps, dctr);
            RUN(128, 2) RUN(256, 2) RUN(256, 4) RUN(512, 2) RUN(512, 4) RUN(1024, 2) RUN(1024, 4) RUN(512, 8)
            if (ms < 0) { printf("n=%6d T=%4d K=%d H=2^%d %s: not launchable (%g)\n", n, c.T, c.K, c.hbits, c.glob ? "global" : "smem", ms); continue; }
            const int bad = verify(c.glob ? "global" : "smem");
            // latency of a single permutation
            float ms1 = -1.f;
#define RUN1(TT, KK) if (c.T == TT && c.K == KK) ms1 = c.glob ? run<TT, KK, true>(dw, dcur, dout, dlast, n, 1, c.hbits, c.cps, dctr) : run<TT, KK, false>(dw, dcur, dout, dlast, n, 1, c.hbits, c.cps, dctr);
            RUN1(128, 2) RUN1(256, 2) RUN1(256, 4) RUN1(512, 2) RUN1(512, 4) RUN1(1024, 2) RUN1(1024, 4) RUN1(512, 8)
            printf("n=%6d P=%5d T=%4d K=%d H=2^%d cps=%d %-6s: %8.3f ms  %7.3f ns/elem  %7.1f GB/s(16B/elem)  one perm %8.1f us  %s\n", n, P, c.T, c.K, c.hbits, c.cps,
                   c.glob ? "global" : "smem", ms, ms * 1e6 / ((double)n * P), 16.0 * n * P / (ms * 1e6), ms1 * 1e3, bad ? "WRONG" : "ok");
        }
        cudaFree(dw); cudaFree(dcur); cudaFree(dout); cudaFree(dlast); cudaFree(dctr);
    }
    return 0;
}
