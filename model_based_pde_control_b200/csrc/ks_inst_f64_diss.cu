// Instantiations of the control-period kernel: T = double, reward mode = kRewardDissipation, P = 4..16.
#include "ks_dispatch.h"
KS_DEFINE_PERIOD_LOOKUP(period_kernel_f64_diss, double, ks::kRewardDissipation)
