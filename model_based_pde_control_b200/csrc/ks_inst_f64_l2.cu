// Instantiations of the control-period kernel: T = double, reward mode = kRewardL2, P = 4..16.
#include "ks_dispatch.h"
KS_DEFINE_PERIOD_LOOKUP(period_kernel_f64_l2, double, ks::kRewardL2)
