// ks_peak.cu -- measurement helper: sustained FP64 FMA throughput of the device, used as the
// self-measured roofline denominator (MEASURED_PEAKS.json carries only HBM and bf16 numbers).
// Not part of the env path.
#include <cuda_runtime.h>

#include "../../include/ks_b200.h"

namespace {

constexpr int kChains = 16;

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double *sink)
{
    double a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
    const double m = 1.0000000001, c = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

__global__ void __launch_bounds__(256) ffma_peak_kernel(int iters, float seed, float *sink)
{
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = seed + threadIdx.x * 1e-6f + i;
    const float m = 1.000001f, c = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}

}  // namespace

extern "C" int ks_bench_fp32_peak(int device, int iters, int repeats, double *tflops_best, double *tflops_mean)
{
    if (iters < 1 || repeats < 1 || !tflops_best) return KS_ERR_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    float *sink = nullptr;
    cudaMalloc(&sink, sizeof(float));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    ffma_peak_kernel<<<blocks, threads>>>(iters / 4 + 1, 1.0f, sink);  // warm-up
    double best = 0.0, sum = 0.0;
    for (int r = 0; r < repeats; ++r) {
        cudaEventRecord(t0);
        ffma_peak_kernel<<<blocks, threads>>>(iters, 1.0f, sink);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flops = 2.0 * kChains * (double)iters * threads * blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
        sum += tf;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(sink);
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) return (int)e;
    *tflops_best = best;
    if (tflops_mean) *tflops_mean = sum / repeats;
    return KS_OK;
}

extern "C" int ks_bench_fp64_peak(int device, int iters, int repeats, double *tflops_best, double *tflops_mean)
{
    if (iters < 1 || repeats < 1 || !tflops_best) return KS_ERR_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    double *sink = nullptr;
    cudaMalloc(&sink, sizeof(double));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    dfma_peak_kernel<<<blocks, threads>>>(iters / 4 + 1, 1.0, sink);  // warm-up
    double best = 0.0, sum = 0.0;
    for (int r = 0; r < repeats; ++r) {
        cudaEventRecord(t0);
        dfma_peak_kernel<<<blocks, threads>>>(iters, 1.0, sink);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flops = 2.0 * kChains * (double)iters * threads * blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
        sum += tf;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(sink);
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) return (int)e;
    *tflops_best = best;
    if (tflops_mean) *tflops_mean = sum / repeats;
    return KS_OK;
}
