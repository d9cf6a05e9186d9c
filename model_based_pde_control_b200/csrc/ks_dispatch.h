// ks_dispatch.h -- kernel lookup shared by the instantiation units and the C-ABI layer.
#pragma once
#include "ks_kernels.cuh"

namespace ks {

constexpr int kMinP = 2;
constexpr int kMaxP = 16;

// One translation unit per (precision, reward mode) so that the units compile in parallel.
// Each returns the __global__ function for `P` points per lane, or nullptr if P is out of range.
const void *period_kernel_f64_l2(int P);
const void *period_kernel_f64_diss(int P);
const void *period_kernel_f32_l2(int P);
const void *period_kernel_f32_diss(int P);
// Spectral ETDRK4 solver (ks_etd.cuh, N = 64 R): one kernel per precision, R in {1, 2, 4} and reward mode.
const void *etd_kernel_f64(int R, int reward_mode);
const void *etd_kernel_f32(int R, int reward_mode);
// Small-batch layout of the same solver (ks_etd16.cuh, N = 64: 16 lanes x 4 registers per env pair).
const void *etd16_kernel_f64(int reward_mode);
const void *etd16_kernel_f32(int reward_mode);

}  // namespace ks

// Expands to the switch over every supported P for one (T, RMODE) pair.
#define KS_DEFINE_PERIOD_LOOKUP(NAME, T, RMODE)                                              \
    const void *ks::NAME(int P)                                                              \
    {                                                                                        \
        switch (P) {                                                                         \
            case 2: return (const void *)&ks::ks_period_kernel<T, 2, RMODE>;                 \
            case 3: return (const void *)&ks::ks_period_kernel<T, 3, RMODE>;                 \
            case 4: return (const void *)&ks::ks_period_kernel<T, 4, RMODE>;                 \
            case 5: return (const void *)&ks::ks_period_kernel<T, 5, RMODE>;                 \
            case 6: return (const void *)&ks::ks_period_kernel<T, 6, RMODE>;                 \
            case 7: return (const void *)&ks::ks_period_kernel<T, 7, RMODE>;                 \
            case 8: return (const void *)&ks::ks_period_kernel<T, 8, RMODE>;                 \
            case 9: return (const void *)&ks::ks_period_kernel<T, 9, RMODE>;                 \
            case 10: return (const void *)&ks::ks_period_kernel<T, 10, RMODE>;               \
            case 11: return (const void *)&ks::ks_period_kernel<T, 11, RMODE>;               \
            case 12: return (const void *)&ks::ks_period_kernel<T, 12, RMODE>;               \
            case 13: return (const void *)&ks::ks_period_kernel<T, 13, RMODE>;               \
            case 14: return (const void *)&ks::ks_period_kernel<T, 14, RMODE>;               \
            case 15: return (const void *)&ks::ks_period_kernel<T, 15, RMODE>;               \
            case 16: return (const void *)&ks::ks_period_kernel<T, 16, RMODE>;               \
            default: return nullptr;                                                         \
        }                                                                                    \
    }
