// Instantiations of the spectral ETDRK4 control-period kernel (ks_etd.cuh): T = double / float,
// N = 64 R with R = 1, 2, 4.
#include "ks_dispatch.h"
#include "ks_etd.cuh"

const void *ks::etd_kernel_f64(int R)
{
    switch (R) {
        case 1: return (const void *)&ks::ks_etd_kernel<double, 1>;
        case 2: return (const void *)&ks::ks_etd_kernel<double, 2>;
        case 4: return (const void *)&ks::ks_etd_kernel<double, 4>;
        default: return nullptr;
    }
}
const void *ks::etd_kernel_f32(int R)
{
    switch (R) {
        case 1: return (const void *)&ks::ks_etd_kernel<float, 1>;
        case 2: return (const void *)&ks::ks_etd_kernel<float, 2>;
        case 4: return (const void *)&ks::ks_etd_kernel<float, 4>;
        default: return nullptr;
    }
}
