// Instantiations of the spectral ETDRK4 control-period kernel (ks_etd.cuh): T = double / float,
// N = 64 R with R = 1, 2, 4, both reward modes.
#include "ks_dispatch.h"
#include "ks_etd.cuh"
#include "ks_etd16.cuh"

#define KS_ETD_LOOKUP(NAME, T)                                                                        \
    const void *ks::NAME(int R, int rmode)                                                            \
    {                                                                                                 \
        const bool l2 = rmode == ks::kRewardL2;                                                       \
        switch (R) {                                                                                  \
            case 1: return l2 ? (const void *)&ks::ks_etd_kernel<T, 1, ks::kRewardL2>                 \
                              : (const void *)&ks::ks_etd_kernel<T, 1, ks::kRewardDissipation>;       \
            case 2: return l2 ? (const void *)&ks::ks_etd_kernel<T, 2, ks::kRewardL2>                 \
                              : (const void *)&ks::ks_etd_kernel<T, 2, ks::kRewardDissipation>;       \
            case 4: return l2 ? (const void *)&ks::ks_etd_kernel<T, 4, ks::kRewardL2>                 \
                              : (const void *)&ks::ks_etd_kernel<T, 4, ks::kRewardDissipation>;       \
            default: return nullptr;                                                                  \
        }                                                                                             \
    }
KS_ETD_LOOKUP(etd_kernel_f64, double)
KS_ETD_LOOKUP(etd_kernel_f32, float)

// Small-batch layout (ks_etd16.cuh): N = 64, 16 lanes x 4 registers per env pair.
const void *ks::etd16_kernel_f64(int rmode)
{
    return rmode == ks::kRewardL2 ? (const void *)&ks::ks_etd16_kernel<double, ks::kRewardL2>
                                  : (const void *)&ks::ks_etd16_kernel<double, ks::kRewardDissipation>;
}
const void *ks::etd16_kernel_f32(int rmode)
{
    return rmode == ks::kRewardL2 ? (const void *)&ks::ks_etd16_kernel<float, ks::kRewardL2>
                                  : (const void *)&ks::ks_etd16_kernel<float, ks::kRewardDissipation>;
}
