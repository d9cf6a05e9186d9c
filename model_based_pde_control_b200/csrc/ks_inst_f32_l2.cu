// Instantiations of the control-period kernel: T = float, reward mode = kRewardL2, P = 4..16.
#include "ks_dispatch.h"
KS_DEFINE_PERIOD_LOOKUP(period_kernel_f32_l2, float, ks::kRewardL2)
