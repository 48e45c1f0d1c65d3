// shuffle_probe.cu -- latency of the pieces of one 32-step Fisher-Yates group (one warp per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__device__ __forceinline__ uint64_t temper(uint64_t y) {
    y ^= (y >> 29) & 0x5555555555555555ULL; y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
    y ^= (y << 37) & 0xFFF7EEE000000000ULL; y ^= (y >> 43); return y;
}
__device__ __forceinline__ int draw_index(uint64_t v, int i) {
    double u = (double)v * 5.421010862427522170037264004349708557128906250e-20;
    if (u >= 1.0) u = 0.99999999999999988897769753748434595763683319091796875;
    return (int)(u * (double)i) + 1;
}
__global__ void k(int mode, int iters, long long* out, unsigned* sink) {
    __shared__ unsigned short idx[16384];
    __shared__ unsigned char tab[2048 + 32];
    if (threadIdx.x < 32) tab[2048 + threadIdx.x] = 0;
    const int lane = threadIdx.x;
    for (int k2 = lane; k2 < 16384; k2 += 32) idx[k2] = k2;
    __syncwarp();
    uint64_t x = 0x9E3779B97F4A7C15ULL * (lane + 1);
    unsigned acc = 0;
    int i0 = 16384;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;  // stand-in for the raw word
        const int i = i0 - lane;
        int j;
        if (mode == 0 || mode == 6) { j = (int)(((x >> 40) * (uint64_t)i) >> 24) + 1; }                      // cheap index (baseline loop cost)
        else j = draw_index(temper(x), i);                                   // temper + u64->double + mul + f2i
        if (mode >= 2 && mode < 5) { const unsigned same = __match_any_sync(FULL, j); acc += __popc(same); }
        if (mode >= 3 && mode < 5) {
            const int m = i0 - j;
            const bool tgt = (m >= 0) && (m < 32) && (m != lane);
            const unsigned tmask = __reduce_or_sync(FULL, tgt ? (1u << m) : 0u);
            acc += tmask;
        }
        if (mode == 4) {
            const int vi = idx[i - 1], vj = idx[j - 1];
            __syncwarp();
            idx[i - 1] = vj; idx[j - 1] = vi;
            __syncwarp();
        }
        if (mode == 5 || mode == 6) {   // claim-table detection + swap (the production scheme)
            unsigned char* cflag = tab + 2048;
            const int slot = j & 2047;
            const int m = i0 - j;
            const bool tgt = (m >= 0) && (m < 32) && (m != lane);
            tab[slot] = (unsigned char)lane;
            if (tgt) cflag[m] = 1;
            __syncwarp();
            const int w = tab[slot];
            const bool loser = (w != lane);
            if (loser) cflag[w] = 1;
            const int vi = idx[i - 1], vj = idx[j - 1];
            __syncwarp();
            const bool conflict = loser || tgt || (cflag[lane] != 0);
            cflag[lane] = 0;
            if (!conflict) { idx[i - 1] = vj; idx[j - 1] = vi; }
            __syncwarp();
            unsigned cm = __ballot_sync(FULL, conflict);
            acc += __popc(cm);
            while (cm) {
                const int l = __ffs(cm) - 1;
                cm &= cm - 1;
                if (lane == l) { const int a = idx[i - 1], b = idx[j - 1]; idx[i - 1] = b; idx[j - 1] = a; }
                __syncwarp();
            }
        }
        i0 -= 32; if (i0 < 1024) i0 = 16384;
    }
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) *out = t1 - t0;
    sink[blockIdx.x * 32 + lane] = acc + (unsigned)x + idx[lane];
}
int main() {
    long long* d; unsigned* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4 * 32 * 148);
    const char* names[] = {"loop + cheap index", "+ temper, u64->f64, mul, f2i", "+ match.any", "+ reduce_or target test", "+ 2 LDS, 2 STS, 2 syncwarp", "full group, claim table (production)", "same with a cheap integer index"};
    for (int mode = 0; mode < 7; ++mode) {
        k<<<148, 32>>>(mode, 20000, d, s);
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-34s %7.1f cycles per 32-step group\n", names[mode], (double)h / 20000.0);
    }
    return 0;
}
