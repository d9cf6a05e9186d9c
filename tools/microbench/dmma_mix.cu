// dmma_mix.cu -- microbenchmark: is the FP64 tensor path (mma.sync m8n8k4 f64, SASS DMMA) a second
// source of FP64 throughput beside the DFMA pipe on B200, i.e. could the linear stencil of the KS
// right-hand side be moved there as a banded matrix product?  Measures SM cycles per loop iteration
// per warp for (a) 8 DFMA, (b) 8 DMMA, (c) 8 DFMA + 8 DMMA interleaved, at 1..8 warps per SM
// sub-partition, and the implied TFLOP/s (a DMMA m8n8k4 is 512 flops per warp, a DFMA 64).
// Development tool; result recorded in profiles/README.md.
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NM>
__global__ void __launch_bounds__(1024) k(int iters, double *sink, long long *cycles)
{
    double a[NF > 0 ? NF : 1], d0[NM > 0 ? NM : 1], d1[NM > 0 ? NM : 1];
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) { d0[i] = i; d1[i] = -i; }
    const double m = 1.0000000001, c = 1e-12, ma = 1e-3 + threadIdx.x * 1e-9, mb = 1.0 - 1e-9 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
            if (i < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
            if (i < NM)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(d0[i]), "+d"(d1[i]) : "d"(ma), "d"(mb));
        }
    }
    long long t1 = clock64();
    double s = 0.0;
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += a[i];
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) s += d0[i] + d1[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int NF, int NM>
void run(const char *name, int nsm, double clk_ghz, double *sink, long long *cyc_d)
{
    const int iters = 20000;
    for (int wps = 1; wps <= 8; wps *= 2) {
        k<NF, NM><<<nsm, 128 * wps>>>(iters, sink, cyc_d);
        cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, cyc_d, sizeof(cyc), cudaMemcpyDeviceToHost);
        const double per_iter = (double)cyc / iters;                       // SM cycles per iteration (all warps of the SMSP)
        const double flops_per_iter_sm = 4.0 * wps * (NF * 64.0 + NM * 512.0);
        printf("{\"mix\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_iter\": %.1f, \"tflops_at_%.3fGHz\": %.2f}\n", name, wps,
               per_iter, clk_ghz, flops_per_iter_sm / per_iter * clk_ghz * nsm / 1e3);
    }
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate / 1e6;
    double *sink;
    long long *cyc;
    cudaMalloc(&sink, 8);
    cudaMalloc(&cyc, 8);
    run<8, 0>("8 dfma", p.multiProcessorCount, ghz, sink, cyc);
    run<0, 8>("8 dmma", p.multiProcessorCount, ghz, sink, cyc);
    run<8, 8>("8 dfma + 8 dmma", p.multiProcessorCount, ghz, sink, cyc);
    run<8, 2>("8 dfma + 2 dmma", p.multiProcessorCount, ghz, sink, cyc);
    return 0;
}
