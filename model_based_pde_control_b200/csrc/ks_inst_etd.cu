// Instantiations of the spectral ETDRK4 control-period kernel (ks_etd.cuh): T = double, float.
#include "ks_dispatch.h"
#include "ks_etd.cuh"

const void *ks::etd_kernel_f64() { return (const void *)&ks::ks_etd_kernel<double>; }
const void *ks::etd_kernel_f32() { return (const void *)&ks::ks_etd_kernel<float>; }
