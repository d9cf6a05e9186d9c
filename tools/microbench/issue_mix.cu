// issue_mix.cu -- microbenchmark: how do FP64 instructions share issue/dispatch with other pipes on
// B200?  Each variant runs a register-resident loop of NF independent DFMA chains mixed with NO
// "other" instructions of one kind and reports SM cycles per loop iteration per warp for 1/2/4 warps
// per SMSP.  Development tool (results summarised in DESIGN.md), not part of the env path.
#include <cstdio>
#include <cuda_runtime.h>

enum Kind { NONE = 0, SEL, LOP3, IADD, IMAD, FFMA, SHFL, ISETP, FSEL, DADD2, SEL2R, DFMA2R_SEL2R };
static const char *kNames[] = {"none", "sel", "lop3", "iadd", "imad", "ffma", "shfl", "isetp_sel", "fsel", "dfma_more", "sel_2reg", "dfma2reg_sel2reg"};

template <int KIND, int NF, int NO>
__global__ void __launch_bounds__(1024) mix_kernel(int iters, double *sink, long long *cycles, int sel_in)
{
    double a[NF > 0 ? NF : 1];
    int x[NO > 0 ? NO : 1];
    float f[NO > 0 ? NO : 1];
    double e[NO > 0 ? NO : 1];
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < (NO > 0 ? NO : 1); ++i) { x[i] = threadIdx.x + i; f[i] = 1.0f + i; e[i] = 2.0 + i; }
    const double m = 1.0000000001, c = 1e-12;
    int p = sel_in;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > NO ? NF : NO); ++i) {
            if (i < NF) {
                if (KIND == DFMA2R_SEL2R) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(a[(i + 3) % NF]), "d"(m));
                else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
            }
            if (i < NO) {
                if (KIND == SEL) asm volatile("{.reg .pred q; setp.ne.s32 q, %2, 0; selp.b32 %0, %0, %1, q;}" : "+r"(x[i]) : "r"(it), "r"(p));
                if (KIND == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(it), "r"(p));
                if (KIND == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(p));
                if (KIND == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(p), "r"(it));
                if (KIND == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0001f), "f"(0.5f));
                if (KIND == SHFL) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(x[i]) : "r"(p));
                if (KIND == ISETP) asm volatile("{.reg .pred q; setp.lt.s32 q, %0, %2; selp.b32 %0, %0, %1, q;}" : "+r"(x[i]) : "r"(it), "r"(p));
                if (KIND == FSEL) asm volatile("{.reg .pred q; setp.ne.s32 q, %2, 0; selp.f32 %0, %0, %1, q;}" : "+f"(f[i]) : "f"(2.0f), "r"(p));
                if (KIND == SEL2R || KIND == DFMA2R_SEL2R) asm volatile("{.reg .pred q; setp.ne.s32 q, %2, 0; selp.b32 %0, %0, %1, q;}" : "+r"(x[i]) : "r"(x[(i + 5) % NO]), "r"(p));
                if (KIND == DADD2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(e[i]) : "d"(m), "d"(c));
            }
        }
    }
    long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < (NO > 0 ? NO : 1); ++i) s += x[i] + f[i] + e[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int KIND, int NF, int NO>
void run(int nsm, double *sink, long long *cyc_d)
{
    const int iters = 20000;
    for (int wps = 1; wps <= 8; wps *= 2) {   // warps per SMSP: block = 4*wps warps, one block per SM
        mix_kernel<KIND, NF, NO><<<nsm, 128 * wps>>>(iters, sink, cyc_d, 1);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        mix_kernel<KIND, NF, NO><<<nsm, 128 * wps>>>(iters, sink, cyc_d, 1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long cyc; cudaMemcpy(&cyc, cyc_d, 8, cudaMemcpyDeviceToHost);
        // cycles per iteration per SMSP (all warps of the SMSP together complete wps iterations)
        printf("{\"kind\": \"%s\", \"n_dfma\": %d, \"n_other\": %d, \"warps_per_smsp\": %d, \"cycles_per_iter_per_warp\": %.2f, "
               "\"smsp_cycles_per_warp_iter\": %.2f, \"ms\": %.3f}\n", kNames[KIND], NF, NO, wps, (double)cyc / iters,
               (double)cyc / iters / wps, ms);
    }
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int nsm = prop.multiProcessorCount;
    double *sink; long long *cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    run<NONE, 8, 0>(nsm, sink, cyc);
    run<NONE, 16, 0>(nsm, sink, cyc);
    run<DADD2, 8, 8>(nsm, sink, cyc);
    run<SEL, 8, 8>(nsm, sink, cyc);
    run<SEL, 8, 4>(nsm, sink, cyc);
    run<SEL, 8, 16>(nsm, sink, cyc);
    run<LOP3, 8, 8>(nsm, sink, cyc);
    run<IADD, 8, 8>(nsm, sink, cyc);
    run<IMAD, 8, 8>(nsm, sink, cyc);
    run<FFMA, 8, 8>(nsm, sink, cyc);
    run<FFMA, 8, 16>(nsm, sink, cyc);
    run<FSEL, 8, 8>(nsm, sink, cyc);
    run<SHFL, 8, 8>(nsm, sink, cyc);
    run<ISETP, 8, 8>(nsm, sink, cyc);
    run<SEL2R, 8, 8>(nsm, sink, cyc);
    run<DFMA2R_SEL2R, 8, 0>(nsm, sink, cyc);
    run<DFMA2R_SEL2R, 8, 4>(nsm, sink, cyc);
    run<DFMA2R_SEL2R, 8, 8>(nsm, sink, cyc);
    run<DFMA2R_SEL2R, 8, 12>(nsm, sink, cyc);
    run<SEL, 0, 16>(nsm, sink, cyc);
    run<FFMA, 0, 16>(nsm, sink, cyc);
    run<IMAD, 0, 16>(nsm, sink, cyc);
    run<LOP3, 0, 16>(nsm, sink, cyc);
    printf("# cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
