// Instantiations of the control-period kernel: T = float, reward mode = kRewardDissipation, P = 4..16.
#include "ks_dispatch.h"
KS_DEFINE_PERIOD_LOOKUP(period_kernel_f32_diss, float, ks::kRewardDissipation)
