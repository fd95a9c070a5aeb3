// Micro-benchmark behind the instruction budget of the float64 on-chip min-sum kernel (onchip_minsum64.cuh): how many
// warp-instructions per clock and SM do B200's FP64 pipe (DADD, DSETP) and the ALU pipe (FSEL pairs = a 64-bit select)
// sustain, alone and side by side? B200 has no double min / max instruction: the reference's min1 / min2 chain costs two
// DSETP plus six FSEL per edge. (A predicated `add.f64 d, x, 0` as a one-instruction 64-bit move on the FP64 pipe is not
// expressible: ptxas turns it back into an unconditional DADD plus two FSEL.)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_pipe fp64_pipe.cu && ./fp64_pipe
// One CTA of 1024 threads per SM (the kernel's geometry), 8 independent chains per thread.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 8;

// MODE 0: DADD chain. 1: DSETP + 64-bit select (2 FSEL). 3: min1 / min2 update as in the kernel (2 DSETP + 6 FSEL).
// 5: FSEL pairs only.
template <int MODE>
__global__ void __launch_bounds__(1024, 1) pipe(unsigned long long *out_clocks, double *sink, int iters, double seed) {
    double a[kChains], b[kChains], c[kChains];
#pragma unroll
    for (int u = 0; u < kChains; ++u) {
        a[u] = seed + threadIdx.x * 1e-3 + u;
        b[u] = 1e300;
        c[u] = 1e300;
    }
    const double step = seed * 0.999;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kChains; ++u) {
            if (MODE == 0) {
                a[u] = __dadd_rn(a[u], step);
            } else if (MODE == 1) {
                a[u] = __dadd_rn(a[u], step);
                b[u] = (a[u] < b[u]) ? a[u] : b[u];
            } else if (MODE == 3) {
                a[u] = __dadd_rn(a[u], step);
                const bool p1 = a[u] < b[u], p2 = a[u] < c[u];
                c[u] = p1 ? b[u] : (p2 ? a[u] : c[u]);
                b[u] = p1 ? a[u] : b[u];
            } else {
                const bool p = (__double2hiint(a[u]) ^ it) & 1;
                const double t = p ? b[u] : c[u];
                c[u] = p ? c[u] : a[u];
                a[u] = t;
                b[u] = p ? a[u] : b[u];
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int u = 0; u < kChains; ++u) s += a[u] + b[u] + c[u];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0) out_clocks[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
static void run(const char *name, double fp64_per_step, double alu_per_step, int sms) {
    unsigned long long *d_clk;
    double *d_sink;
    cudaMalloc(&d_clk, sms * sizeof(unsigned long long));
    cudaMalloc(&d_sink, 8);
    const int iters = 4096;
    pipe<MODE><<<sms, 1024>>>(d_clk, d_sink, 16, 1.0);
    pipe<MODE><<<sms, 1024>>>(d_clk, d_sink, iters, 1.0);
    cudaDeviceSynchronize();
    unsigned long long clk[256];
    cudaMemcpy(clk, d_clk, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)clk[i];
    mean /= sms;
    const double steps = (double)iters * kChains * 32;   // warp-level steps per SM
    printf("%-46s %8.3f clocks per warp-step per SM | FP64-pipe %.2f, ALU-pipe %.2f warp-instructions per clock and SM\n", name, mean / steps,
           fp64_per_step * steps / mean, alu_per_step * steps / mean);
    cudaFree(d_clk);
    cudaFree(d_sink);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("%s, %d SMs, 1024 threads per SM, %d independent chains per thread\n", prop.name, sms, kChains);
    run<0>("DADD", 1, 0, sms);
    run<1>("DADD + DSETP + 2 FSEL", 2, 2, sms);
    run<3>("DADD + min1/min2: 2 DSETP + 6 FSEL", 3, 6, sms);
    run<5>("6 FSEL (64-bit selects only)", 0, 6, sms);
    return 0;
}
