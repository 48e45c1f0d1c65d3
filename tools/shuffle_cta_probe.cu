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
        ShufDraws<true> src;
        src.win = words + (size_t)p * n;
        double* sx = out + (size_t)p * (n + 1);
        if (GLOBAL) shuffle_cta<T, K>(LastGlobal32{lastg + (size_t)blockIdx.x * (n + 1)}, claim, H - 1, epoch, n, src, cur, nullptr, sx);
        else shuffle_cta<T, K>(LastSmem16{(unsigned short*)(smem_raw + (size_t)H * 4)}, claim, H - 1, epoch, n, src, cur, nullptr, sx);
    }
}

template <int T, int K, int R>
__global__ void __cluster_dims__(R, 1, 1) __launch_bounds__(T) k_probe_cluster(const uint64_t* words, const double* cur, double* out, int n, int P,
                                                                             int hbits, unsigned* ctr) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* claim = (unsigned*)smem_raw;
    unsigned* last = (unsigned*)(smem_raw + ((size_t)4 << hbits));
    const int H = 1 << hbits;
    for (int k = threadIdx.x; k < H; k += T) claim[k] = 0;
    __shared__ int s_g;
    unsigned epoch = 0;
    cg::cluster_group cl = cg::this_cluster();
    for (;;) {
        cl.sync();
        if (cl.block_rank() == 0 && threadIdx.x == 0) {
            const int g = (int)atomicAdd(ctr, 1u);
            for (int r = 0; r < R; ++r) *cl.map_shared_rank(&s_g, r) = g;
        }
        cl.sync();
        const int p = s_g;
        if (p >= P) break;
        ShufDraws<true> src;
        src.win = words + (size_t)p * n;
        shuffle_cluster<T, K, R>(last, claim, H - 1, epoch, n, src, cur, nullptr, out + (size_t)p * (n + 1));
    }
}

template <int T, int K, int R>
float run_cluster(const uint64_t* dw, const double* dcur, double* dout, int n, int P, int hbits, unsigned* dctr) {
    size_t smem = ((size_t)4 << hbits) + 4 * (size_t)(n / R + 4);
    if (cudaFuncSetAttribute(k_probe_cluster<T, K, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    if (R > 8) cudaFuncSetAttribute(k_probe_cluster<T, K, R>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / R * R); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = R; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    cudaOccupancyMaxActiveClusters(&ncl, k_probe_cluster<T, K, R>, &cfg);
    if (ncl < 1) { cudaGetLastError(); return -1.f; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(dctr, 0, 4);
        cudaEventRecord(e0);
        k_probe_cluster<T, K, R><<<ncl * R, T, smem>>>(dw, dcur, dout, n, P, hbits, dctr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return -2.f; }
    printf("   [clusters resident: %d]", ncl);
    return best;
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
        const Cfg cfgs[] = {{128, 2, 11, 16, false}, {256, 2, 11, 8, false}, {512, 2, 12, 4, false}, {1024, 2, 13, 2, false}, {1024, 4, 15, 1, true}};
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
        if (n > 20000) {
            struct CC { int T, K, R, hbits; };
            const CC ccs[] = {{1024, 2, 4, 12}, {1024, 4, 4, 12}, {1024, 4, 4, 13}, {512, 4, 4, 12}, {1024, 2, 8, 12}, {1024, 4, 8, 12}, {1024, 4, 2, 12}};
            for (const CC& c : ccs) {
                cudaMemset(dout, 0, (size_t)P * (n + 1) * 8);
                float ms = -1.f, ms1 = -1.f;
#define RUNC(TT, KK, RR) if (c.T == TT && c.K == KK && c.R == RR) { ms = run_cluster<TT, KK, RR>(dw, dcur, dout, n, P, c.hbits, dctr); }
                RUNC(1024, 2, 4) RUNC(1024, 4, 4) RUNC(512, 4, 4) RUNC(1024, 2, 8) RUNC(1024, 4, 8) RUNC(1024, 4, 2)
                if (ms < 0) { printf("n=%6d cluster T=%d K=%d R=%d: not launchable\n", n, c.T, c.K, c.R); continue; }
                const int bad = verify("cluster");
#define RUNC1(TT, KK, RR) if (c.T == TT && c.K == KK && c.R == RR) { ms1 = run_cluster<TT, KK, RR>(dw, dcur, dout, n, 1, c.hbits, dctr); }
                RUNC1(1024, 2, 4) RUNC1(1024, 4, 4) RUNC1(512, 4, 4) RUNC1(1024, 2, 8) RUNC1(1024, 4, 8) RUNC1(1024, 4, 2)
                printf("\nn=%6d P=%5d T=%4d K=%d H=2^%d R=%d cluster: %8.3f ms  %7.3f ns/elem  %7.1f GB/s(16B/elem)  one perm %8.1f us  %s\n", n, P, c.T, c.K, c.hbits, c.R,
                       ms, ms * 1e6 / ((double)n * P), 16.0 * n * P / (ms * 1e6), ms1 * 1e3, bad ? "WRONG" : "ok");
            }
        }
        cudaFree(dw); cudaFree(dcur); cudaFree(dout); cudaFree(dlast); cudaFree(dctr);
    }
    return 0;
}
