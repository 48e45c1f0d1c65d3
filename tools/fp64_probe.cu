// fp64_probe.cu -- issue-rate microbenchmarks for the instruction mixes the arc scan could use.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/fp64_probe tools/fp64_probe.cu
// Prints lane-instructions per clock per SM and T lane-inst/s for each mix.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

// mix 0: DADD only (8 independent chains)
__global__ void __launch_bounds__(256) k_dadd(double* out, double step) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = a[k] + step;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix 1: DADD + DSETP (what k_scan does today)
__global__ void __launch_bounds__(256) k_dadd_dsetp(double* out, double step) {
    double a[8], th[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = (double)(threadIdx.x + k) * 1e-3; th[k] = 1e300 + k; }
    bool flag = false;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) { a[k] = a[k] + step; flag |= fabs(a[k]) > th[k]; }
    }
    double s = flag ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix 2: DADD + integer compare of the high word: (hi & 0x7fffffff) > thi
__global__ void __launch_bounds__(256) k_dadd_icmp(double* out, double step) {
    double a[8];
    int th[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = (double)(threadIdx.x + k) * 1e-3; th[k] = 0x7fe00000 + k; }
    bool flag = false;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a[k] = a[k] + step;
                flag |= (__double2hiint(a[k]) & 0x7fffffff) > th[k];
            }
    }
    double s = flag ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix 3: DADD + integer max of the masked high word (one IMNMX accumulator per diagonal)
__global__ void __launch_bounds__(256) k_dadd_imax(double* out, double step) {
    double a[8];
    int m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = (double)(threadIdx.x + k) * 1e-3; m[k] = 0; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a[k] = a[k] + step;
                m[k] = max(m[k], __double2hiint(a[k]) & 0x7fffffff);
            }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + (double)m[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix 4: DADD + unsigned max of (hi << 1) -- drops the sign with a shift instead of a mask
__global__ void __launch_bounds__(256) k_dadd_umax(double* out, double step) {
    double a[8];
    unsigned m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = (double)(threadIdx.x + k) * 1e-3; m[k] = 0; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a[k] = a[k] + step;
                m[k] = max(m[k], ((unsigned)__double2hiint(a[k])) << 1);
            }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + (double)m[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix 5: DADD + DMNMX of |.| (fmax accumulators)
__global__ void __launch_bounds__(256) k_dadd_dmax(double* out, double step) {
    double a[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = (double)(threadIdx.x + k) * 1e-3; m[k] = 0; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) { a[k] = a[k] + step; m[k] = fmax(m[k], fabs(a[k])); }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + m[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix 6: FP32 screen: FADD + FSETP on float copies (for a float pre-filter variant)
__global__ void __launch_bounds__(256) k_fadd_fsetp(double* out, float step) {
    float a[8], th[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = (float)(threadIdx.x + k) * 1e-3f; th[k] = 1e30f + k; }
    bool flag = false;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) { a[k] = a[k] + step; flag |= fabsf(a[k]) > th[k]; }
    }
    float s = flag ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix 7: independent DADD (no dependent chain: a[k] = b[k] + c[j]) + integer compare, as in the real scan
__global__ void __launch_bounds__(256) k_dadd_indep_icmp(double* out, double step) {
    double w[8], av[8];
    int th[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { w[k] = (double)(threadIdx.x + k) * 1e-3; av[k] = step * (k + 1); th[k] = 0x7fe00000 + k; }
    bool flag = false;
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double d = w[(s + k) & 7] - av[s];
                flag |= (__double2hiint(d) & 0x7fffffff) > th[k];
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) { w[k] += step; }
    }
    double s = flag ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += w[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// mix 8: FP32 screen with packed adds: FADD2 (2 arcs) + FMNMX3 |m|,|d.lo|,|d.hi| (2 arcs) -> 1 instruction per arc
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
__global__ void __launch_bounds__(256) k_fadd2_fmnmx3(double* out, float step) {
    unsigned long long w[8], av[4];
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { w[k] = pack2((float)(threadIdx.x + k) * 1e-3f, (float)(threadIdx.x + k + 1) * 1e-3f); m[k] = 0.f; }
#pragma unroll
    for (int k = 0; k < 4; ++k) av[k] = pack2(step * (k + 1), step * (k + 2));
    const unsigned long long inc = pack2(step, step);
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                unsigned long long d;
                asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(w[(s + k) & 7]), "l"(av[s]));
                const float dlo = __uint_as_float((unsigned)d), dhi = __uint_as_float((unsigned)(d >> 32));
                asm("max.abs.f32 %0, %1, %2, %3;" : "=f"(m[k]) : "f"(m[k]), "f"(dlo), "f"(dhi));
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) asm("add.rn.f32x2 %0, %1, %2;" : "=l"(w[k]) : "l"(w[k]), "l"(inc));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += m[k] + __uint_as_float((unsigned)w[k]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix 9: scalar FP32 screen: FADD + FMNMX |.| per arc (2 instructions per arc)
__global__ void __launch_bounds__(256) k_fadd_fmnmx(double* out, float step) {
    float w[8], av[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { w[k] = (float)(threadIdx.x + k) * 1e-3f; av[k] = step * (k + 1); m[k] = 0.f; }
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], fabsf(w[(s + k) & 7] - av[s]));
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] += step;
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += m[k] + w[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix 10: independent DADD + VIADDMNMX.U32 on the high word (1 FP64 + 1 INT instruction per arc)
__global__ void __launch_bounds__(256) k_dadd_indep_umax(double* out, double step) {
    double w[8], av[8];
    unsigned m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { w[k] = (double)(threadIdx.x + k) * 1e-3; av[k] = step * (k + 1); m[k] = 0; }
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double d = w[(s + k) & 7] - av[s];
                m[k] = max(m[k], ((unsigned)__double2hiint(d)) << 1);
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) { w[k] += step; }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += w[k] + (double)m[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// latency: one dependent DADD chain per thread, one warp per SM
__global__ void k_dadd_latency(double* out, double step, long long* cycles) {
    double a = threadIdx.x * 1e-3;
    const long long t0 = clock64();
#pragma unroll 64
    for (int it = 0; it < 65536; ++it) a = a + step;
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <class F>
static void run(const char* name, F launch, double fp_inst_per_thread, int sms, double mhz, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double inst = (double)grid * 256 * fp_inst_per_thread;
    const double rate = inst / (best * 1e-3);
    printf("%-28s %8.3f ms  %7.3f T arith-inst/s  %6.2f lanes/clk/SM (at %.0f MHz)\n", name, best, rate / 1e12,
           rate / (sms * mhz * 1e6), mhz);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    const double mhz = khz / 1000.0;
    const int grid = sms * 8;
    double* out; cudaMalloc(&out, sizeof(double) * grid * 256);
    printf("%s, %d SMs, %.0f MHz\n", p.name, sms, mhz);
    const double per = (double)ITERS * 32;  // arithmetic (DADD/FADD) instructions per thread
    run("DADD", [&] { k_dadd<<<grid, 256>>>(out, 1e-9); }, per, sms, mhz, grid);
    run("DADD+DSETP (per DADD)", [&] { k_dadd_dsetp<<<grid, 256>>>(out, 1e-9); }, per, sms, mhz, grid);
    run("DADD+LOP+ISETP (per DADD)", [&] { k_dadd_icmp<<<grid, 256>>>(out, 1e-9); }, per, sms, mhz, grid);
    run("DADD+LOP+IMNMX (per DADD)", [&] { k_dadd_imax<<<grid, 256>>>(out, 1e-9); }, per, sms, mhz, grid);
    run("DADD+SHL+UMNMX (per DADD)", [&] { k_dadd_umax<<<grid, 256>>>(out, 1e-9); }, per, sms, mhz, grid);
    run("DADD+DMNMX (per DADD)", [&] { k_dadd_dmax<<<grid, 256>>>(out, 1e-9); }, per, sms, mhz, grid);
    run("FADD+FSETP (per FADD)", [&] { k_fadd_fsetp<<<grid, 256>>>(out, 1e-9f); }, per, sms, mhz, grid);
    run("indep DADD+LOP+ISETP", [&] { k_dadd_indep_icmp<<<grid, 256>>>(out, 1e-9); }, per, sms, mhz, grid);
    run("FADD2+FMNMX3 (per arc)", [&] { k_fadd2_fmnmx3<<<grid, 256>>>(out, 1e-9f); }, per, sms, mhz, grid);
    run("FADD+FMNMX (per arc)", [&] { k_fadd_fmnmx<<<grid, 256>>>(out, 1e-9f); }, per, sms, mhz, grid);
    run("indep DADD+VIADDMNMX (per arc)", [&] { k_dadd_indep_umax<<<grid, 256>>>(out, 1e-9); }, per, sms, mhz, grid);
    {
        long long* cyc; cudaMalloc(&cyc, 8);
        for (int w = 1; w <= 16; w *= 2) {
            k_dadd_latency<<<1, 32 * w>>>(out, 1e-9, cyc);
            long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("dependent DADD chain, %2d warps on one SM: %.2f cycles per DADD per warp\n", w, (double)h / 65536.0);
        }
        cudaFree(cyc);
    }
    cudaFree(out);
    return 0;
}
