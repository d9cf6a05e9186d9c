// ks_kernels.cuh -- device code of the batched Kuramoto-Sivashinsky control-period kernel (sm_100a).
//
// One launch advances every environment of the handle by K control periods (K = 1 for
// ks_step, K = burn-in length for ks_reset).  Layout: an environment's N grid points are
// spread over `lanes` adjacent lanes of ONE warp, P = N / lanes contiguous points per lane (2..16), all
// state in registers for the whole launch.  The +-4 point halo of the periodic finite-difference
// stencils comes from the two neighbouring lanes by warp shuffle; HBM is touched only at
// control-period boundaries (128-bit loads/stores of the state, float32 observation, reward).
//
// Algorithm = the reference's own scheme (NOT a spectral method):
//   rhs(u)  = -uxxxx - uxx - 1/2 d/dx(u^2) + phi          pdegym/kuramoto/kuramoto.py:118-129
//     uxx, uxxxx: 7-/9-tap 6th-order central stencils (kuramoto.py:26-27), merged here into one
//                 symmetric 9-tap stencil W (the two are only ever used as a sum);
//     d/dx(u^2):  5-tap 2nd-order upwind, forward where u < 0, backward where u >= 0
//                 (kuramoto.py:24-25,120-122);
//   classic RK4, reward of the pre-step state every sub-step         kuramoto.py:83-90
//   phi = a @ F in float32 as a sequential FMA chain                 transforms.py:262-265
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ks {

#ifndef KS_BLOCK_THREADS
#define KS_BLOCK_THREADS 128
#endif
constexpr int kBlockThreads = KS_BLOCK_THREADS;
constexpr int kHalo = 4;
constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kMaxRemote = 15;   // peers a launch can mirror its outputs to (world size <= 16)

enum { kRewardL2 = 0, kRewardDissipation = 1 };

// Stencil coefficients pre-divided by powers of dx on the host in float64, then rounded to T.
template <typename T>
struct Coef {
    T W[5];   // W[k] = -(D4[k]/dx^4 + D2[k]/dx^2): centre k=0, symmetric pairs k=1..4
    T A[5];   // A[k] = UPWIND[k] * (-1/2) / dx
    T D2[4];  // D2[k]/dx^2, centre and pairs 1..3 (dissipation reward only)
    T dt_half, dt_full, dt_sixth;
};

struct Params {
    void *u;               // [B,N] T   state, in/out
    int32_t *timestep;     // [B]
    uint8_t *nonfinite;    // [B]       sticky flags
    const float *F;        // [J,N]     forcing matrix
    const float *actions;  // [K,B,J]   or nullptr (no-op periods)
    const float *phi;      // [B,N]     override of a@F (K == 1 only) or nullptr
    float *obs;            // [K,B,N]   or nullptr
    double *reward;        // [K,B]     or nullptr
    uint8_t *truncated;    // [K,B]     or nullptr
    int32_t *step;         // [K,B]     or nullptr
    uint8_t *nonfinite_out;  // [K,B]   or nullptr: per-period copy of the (sticky) non-finite flag
    const uint8_t *mask;   // [B]       or nullptr; envs with mask[b] == 0 are left untouched
    int B, N, J, K, cfg_steps, max_episode_steps;
    int lanes;             // lanes per environment (N = lanes * P)
    int envs_per_warp;     // 32 / lanes
    int reset_timestep;    // 1: timestep = 0 after the launch (burn-in of reset())
    int obs_stride;        // SensorTransform stride s: obs = u[s/2::s] (transforms.py:236-239); 1 = all
    int obs_len;           // observation length, ceil((N - s/2) / s)
    double inv_cfg_steps, inv_N;
    // Fused all-gather (ks_step_gather): every output store of the period epilogue is repeated at
    // `pointer + remote_delta[q]`, q < n_remote -- this rank's slot in peer q's gather buffer, mapped
    // into this process with CUDA IPC, i.e. plain st.global over NVLink.  n_remote = 0 otherwise.
    int n_remote;
    int skip_fence;        // measurement knob (KS_GATHER_DEBUG=nofence): leave out the system-scope fence at kernel end
    // NVLS multicast (ks_gather_attach with a multicast address): `pointer + mc_delta` is the multicast alias of this
    // rank's slot in EVERY rank's gather buffer; the observation rows then leave the GPU once (multimem.st, replicated
    // by the NVSwitch) instead of once per peer.  0 = not available.
    long long mc_delta;
    long long remote_delta[kMaxRemote];
};

template <typename U>
__device__ __forceinline__ U *remote_ptr(U *local, long long delta)
{
    return reinterpret_cast<U *>(reinterpret_cast<char *>(local) + delta);
}

// One store to an NVLS multicast address: the NVSwitch replicates it into every bound GPU's memory.
__device__ __forceinline__ void multimem_st(float *mc, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st(float *mc, float v)
{
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(mc), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T fma_t(T a, T b, T c);
template <>
__device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
template <>
__device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }

// vectorised global access of P contiguous elements (widest naturally aligned type)
template <int P>
__device__ __forceinline__ void load_row(const double *__restrict__ src, double (&dst)[P])
{
    if constexpr (P % 2 == 0) {
#pragma unroll
        for (int i = 0; i < P; i += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(src + i);
            dst[i] = v.x;
            dst[i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) dst[i] = src[i];
    }
}
template <int P>
__device__ __forceinline__ void load_row(const float *__restrict__ src, float (&dst)[P])
{
    if constexpr (P % 4 == 0) {
#pragma unroll
        for (int i = 0; i < P; i += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(src + i);
            dst[i] = v.x; dst[i + 1] = v.y; dst[i + 2] = v.z; dst[i + 3] = v.w;
        }
    } else if constexpr (P % 2 == 0) {
#pragma unroll
        for (int i = 0; i < P; i += 2) {
            const float2 v = *reinterpret_cast<const float2 *>(src + i);
            dst[i] = v.x; dst[i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) dst[i] = src[i];
    }
}
template <int P>
__device__ __forceinline__ void store_row(double *__restrict__ dst, const double (&src)[P])
{
    if constexpr (P % 2 == 0) {
#pragma unroll
        for (int i = 0; i < P; i += 2) *reinterpret_cast<double2 *>(dst + i) = make_double2(src[i], src[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) dst[i] = src[i];
    }
}
template <int P>
__device__ __forceinline__ void store_row(float *__restrict__ dst, const float (&src)[P])
{
    if constexpr (P % 4 == 0) {
#pragma unroll
        for (int i = 0; i < P; i += 4)
            *reinterpret_cast<float4 *>(dst + i) = make_float4(src[i], src[i + 1], src[i + 2], src[i + 3]);
    } else if constexpr (P % 2 == 0) {
#pragma unroll
        for (int i = 0; i < P; i += 2) *reinterpret_cast<float2 *>(dst + i) = make_float2(src[i], src[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) dst[i] = src[i];
    }
}

// t = neg ? fwd : -bwd, done on the integer pipe (sign flip = XOR of the high word, selection =
// SEL) so that the upwind switch costs no FP64-pipe slots.  `nbwd_hi` is the pre-flipped high
// word of bwd.
__device__ __forceinline__ double select_signed(bool neg, int fwd_lo, int fwd_hi, int bwd_lo, int nbwd_hi)
{
    return __hiloint2double(neg ? fwd_hi : nbwd_hi, neg ? fwd_lo : bwd_lo);
}

// Upwind switch `u < 0` (kuramoto.py:122) on the integer pipe: the sign bit of the high word.
// This equals the IEEE compare for every value except -0.0, and -0.0 cannot occur here: the state
// is canonicalised (u + 0.0) when it is loaded and every later value is the result of an FMA whose
// addend is not -0.0.  -DKS_UPWIND_DSETP switches to the FP64-pipe compare for cross-checking.
__device__ __forceinline__ bool is_negative(double x)
{
#ifdef KS_UPWIND_DSETP
    return x < 0.0;
#else
    return __double2hiint(x) < 0;
#endif
}
__device__ __forceinline__ bool is_negative(float x) { return x < 0.0f; }

template <typename T>
struct RewardAcc {
    T a;  // l2: sum u^2          dissipation: sum uxx^2
    T b;  //                      dissipation: sum (ux/2)^2
    T c;  //                      dissipation: sum u*phi
};

// ---------------------------------------------------------------------------------------------
// One RK4 stage for the P points of this lane.
//   STAGE 0: k1 = f(u)            us = u + dt/2 k1   acc = k1          (+ reward of u)
//   STAGE 1: k2 = f(us)           us = u + dt/2 k2   acc += 2 k2
//   STAGE 2: k3 = f(us)           us = u + dt   k3   acc += 2 k3
//   STAGE 3: k4 = f(us)           u  = u + dt/6 (acc + k4)
// ---------------------------------------------------------------------------------------------
template <typename T, int P, int STAGE, int RMODE>
__device__ __forceinline__ void rk4_stage(T (&u)[P], T (&us)[P], T (&acc)[P], const T (&phi)[P],
                                          RewardAcc<T> &racc, const Coef<T> &c, int srcL, int srcR, int srcL2, int srcR2)
{
    constexpr int H = kHalo;
    static_assert(2 * P >= H, "the halo must come from at most two lanes on each side");
    T h[P + 2 * H];
#pragma unroll
    for (int i = 0; i < P; ++i) h[H + i] = (STAGE == 0) ? u[i] : us[i];
    // periodic halo: my left halo = the H points before my first one -- the last H points of the lane to the left
    // (P >= 4), or the points of the two lanes to the left (P = 2, 3: the small-batch layouts) -- and vice versa
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const int hopL = (H - k + P - 1) / P, regL = P * hopL + k - H;      // my point k - H lives there
        const int hopR = k / P + 1, regR = k % P;                            // my point P + k lives there
        h[k] = __shfl_sync(kFullMask, h[H + regL], hopL == 1 ? srcL : srcL2);
        h[P + H + k] = __shfl_sync(kFullMask, h[H + regR], hopR == 1 ? srcR : srcR2);
    }

    T q[P + 2 * H];
#pragma unroll
    for (int i = 0; i < P + 2 * H; ++i) q[i] = h[i] * h[i];

    if constexpr (STAGE == 0 && RMODE == kRewardL2) {
        // reward of the pre-step state: sum of u^2 over my points (tree to keep the chain short)
        T s[P];
#pragma unroll
        for (int i = 0; i < P; ++i) s[i] = q[H + i];
#pragma unroll
        for (int w = 1; w < P; w <<= 1)
#pragma unroll
            for (int i = 0; i + w < P; i += 2 * w) s[i] += s[i + w];
        racc.a += s[0];
    }

    [[maybe_unused]] int qlo[P + 2 * H], qhi[P + 2 * H], nqhi[P + 2 * H];
    [[maybe_unused]] T nq[P + 2 * H];
    if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int i = 0; i < P + 2 * H; ++i) {
            qlo[i] = __double2loint(q[i]);
            qhi[i] = __double2hiint(q[i]);
            nqhi[i] = qhi[i] ^ 0x80000000;
        }
    } else {
#pragma unroll
        for (int i = 0; i < P + 2 * H; ++i) nq[i] = -q[i];
    }

    // merged linear part  -uxxxx - uxx + phi
    T linv[P];
#ifndef KS_LIN_GATHER
    // scatter order (default): every input h[j] is applied to all the outputs it touches back to
    // back, so that consecutive DFMAs share one source operand (operand-reuse cache) -- the register
    // file delivers only ~2 32-bit operands per lane per cycle, which FP64 ops with two register
    // operands already use up (tools/microbench/issue_mix.cu).  Point i accumulates its nine taps
    // in ascending order of the input index whatever P is.  8 % faster at 4096 envs, 4 % at 65 536,
    // than the symmetric-pair form below (4 DADD + 5 DFMA per point, same FP64 instruction count).
#pragma unroll
    for (int i = 0; i < P; ++i) linv[i] = phi[i];
#pragma unroll
    for (int j = 0; j < P + 2 * H; ++j) {
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int d = (i + H) - j;
            if (d >= -4 && d <= 4) linv[i] = fma_t<T>(h[j], c.W[d < 0 ? -d : d], linv[i]);
        }
    }
#else
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int m = i + H;
        T lin = fma_t<T>(c.W[0], h[m], phi[i]);
#pragma unroll
        for (int k = 1; k <= 4; ++k) lin = fma_t<T>(c.W[k], h[m + k] + h[m - k], lin);
        linv[i] = lin;
    }
#endif

#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int m = i + H;
        const T lin = linv[i];
        // upwind part  -1/2 d/dx(u^2): sum_k A[k] * (neg ? q[m+k] : -q[m-k])
        // (in the dissipation-reward stage the upwind sum is needed on its own; otherwise the
        // chain simply continues from `lin`)
        constexpr bool kSplit = (STAGE == 0 && RMODE == kRewardDissipation);
        const bool neg = is_negative(h[m]);
        // (computing BOTH one-sided chains on the FP64 pipe and selecting the result -- 10 DFMA + 2 SEL per
        // point instead of 5 DFMA + 9 SEL, in point order or in scatter order -- was measured in round 2:
        // 0.4108 / 0.4091 ms against 0.4023 ms per period at 4096 envs, 5.60 / 5.59 against 5.42 ms at
        // 65 536; profiles/round2_sweep_upwind_variants.jsonl)
        T t[5];
        if constexpr (sizeof(T) == 8) {
            t[0] = select_signed(neg, qlo[m], qhi[m], qlo[m], nqhi[m]);
#pragma unroll
            for (int k = 1; k <= 4; ++k)
                t[k] = select_signed(neg, qlo[m + k], qhi[m + k], qlo[m - k], nqhi[m - k]);
        } else {
            t[0] = neg ? q[m] : nq[m];
#pragma unroll
            for (int k = 1; k <= 4; ++k) t[k] = neg ? q[m + k] : nq[m - k];
        }
        T kval = kSplit ? c.A[0] * t[0] : fma_t<T>(c.A[0], t[0], lin);
#pragma unroll
        for (int k = 1; k <= 4; ++k) kval = fma_t<T>(c.A[k], t[k], kval);
        if constexpr (kSplit) {
            T uxx = c.D2[0] * h[m];
#pragma unroll
            for (int k = 1; k <= 3; ++k) uxx = fma_t<T>(c.D2[k], h[m + k] + h[m - k], uxx);
            racc.a = fma_t<T>(uxx, uxx, racc.a);
            racc.b = fma_t<T>(kval, kval, racc.b);     // kval = -ux/2 here  ->  ux^2 = 4 kval^2
            racc.c = fma_t<T>(h[m], phi[i], racc.c);
            kval = lin + kval;
        }

        if constexpr (STAGE == 0) {
            acc[i] = kval;
            us[i] = fma_t<T>(c.dt_half, kval, u[i]);
        } else if constexpr (STAGE == 1) {
            acc[i] = fma_t<T>(T(2), kval, acc[i]);
            us[i] = fma_t<T>(c.dt_half, kval, u[i]);
        } else if constexpr (STAGE == 2) {
            acc[i] = fma_t<T>(T(2), kval, acc[i]);
            us[i] = fma_t<T>(c.dt_full, kval, u[i]);
        } else {
            u[i] = fma_t<T>(c.dt_sixth, acc[i] + kval, u[i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The control-period kernel
// ---------------------------------------------------------------------------------------------
template <typename T, int P, int RMODE>
__global__ void __launch_bounds__(kBlockThreads) ks_period_kernel(const Params p, const Coef<T> c)
{
    // gather mode only: the warp's observation rows are staged here so that the stores into the
    // peers' buffers leave as fully coalesced 16-byte-per-lane runs (NVLink packets of 128 B+
    // instead of scattered 16 B pieces)
    __shared__ __align__(16) float s_obs[kBlockThreads / 32][32 * P];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kBlockThreads + threadIdx.x) >> 5;
    if (warp * p.envs_per_warp >= p.B) return;  // warp-uniform

    const int sub = lane / p.lanes;             // environment slot inside the warp
    const int l = lane - sub * p.lanes;         // my position inside the environment
    const int env_raw = warp * p.envs_per_warp + sub;
    bool active = (sub < p.envs_per_warp) && (env_raw < p.B);
    if (active && p.mask != nullptr) active = p.mask[env_raw] != 0;
    if (__ballot_sync(kFullMask, active) == 0u) return;  // nothing to do in this warp
    const int env = active ? env_raw : 0;
    // idle lanes keep shuffling (full-mask collectives) on harmless data and never store
    const int base = sub * p.lanes;
    const int srcL = active ? base + (l + p.lanes - 1) % p.lanes : lane;
    const int srcR = active ? base + (l + 1) % p.lanes : lane;
    const int srcL2 = active ? base + (l + 2 * p.lanes - 2) % p.lanes : lane;    // two lanes away (P < 4 only)
    const int srcR2 = active ? base + (l + 2) % p.lanes : lane;
    const size_t off = (size_t)env * p.N + (size_t)(active ? l : 0) * P;

    T u[P], us[P], acc[P], phi[P];
    T *ug = static_cast<T *>(p.u) + off;
    load_row<P>(ug, u);
#pragma unroll
    for (int i = 0; i < P; ++i) { u[i] += T(0); us[i] = T(0); acc[i] = T(0); }   // -0.0 -> +0.0
    int ts = p.timestep[env];
    bool was_bad = p.nonfinite[env] != 0;

    for (int k = 0; k < p.K; ++k) {
        // ---- jet forcing for this period (float32, transforms.py:262-265) ----
        if (p.phi != nullptr) {
            float f[P];
            load_row<P>(p.phi + off, f);
#pragma unroll
            for (int i = 0; i < P; ++i) phi[i] = T(f[i]);
        } else if (p.actions != nullptr) {
            const float *a = p.actions + ((size_t)k * p.B + env) * p.J;
            const float *Fl = p.F + (size_t)(active ? l : 0) * P;
            float f[P];
#pragma unroll
            for (int i = 0; i < P; ++i) f[i] = 0.0f;
            for (int j = 0; j < p.J; ++j) {
                const float aj = __ldg(a + j);
                float fr[P];
                load_row<P>(Fl + (size_t)j * p.N, fr);
#pragma unroll
                for (int i = 0; i < P; ++i) f[i] = __fmaf_rn(aj, fr[i], f[i]);
            }
#pragma unroll
            for (int i = 0; i < P; ++i) phi[i] = T(f[i]);
        } else {
#pragma unroll
            for (int i = 0; i < P; ++i) phi[i] = T(0);
        }

        // ---- cfg_steps classic RK4 sub-steps, reward of the pre-step state ----
        RewardAcc<T> racc{T(0), T(0), T(0)};
        for (int s = 0; s < p.cfg_steps; ++s) {
            rk4_stage<T, P, 0, RMODE>(u, us, acc, phi, racc, c, srcL, srcR, srcL2, srcR2);
            rk4_stage<T, P, 1, RMODE>(u, us, acc, phi, racc, c, srcL, srcR, srcL2, srcR2);
            rk4_stage<T, P, 2, RMODE>(u, us, acc, phi, racc, c, srcL, srcR, srcL2, srcR2);
            rk4_stage<T, P, 3, RMODE>(u, us, acc, phi, racc, c, srcL, srcR, srcL2, srcR2);
        }

        // ---- period epilogue: reward, flags, observation ----
        double mine;
        if constexpr (RMODE == kRewardL2) mine = (double)racc.a;
        else mine = (double)racc.a + 4.0 * (double)racc.b + (double)racc.c;
        double tot = 0.0;
        for (int j = 0; j < p.lanes; ++j) tot += __shfl_sync(kFullMask, mine, (base + j) & 31);

        bool bad = false;
#pragma unroll
        for (int i = 0; i < P; ++i) bad |= !(fabs((double)u[i]) <= 1.7976931348623157e308);
        const unsigned badmask = __ballot_sync(kFullMask, bad);
        const unsigned grp = (p.lanes == 32 ? kFullMask : ((1u << p.lanes) - 1u)) << (base & 31);

        ts += 1;
        if (active) {
            const size_t kb = (size_t)k * p.B + env;
            if (p.obs != nullptr) {
                if (p.obs_stride <= 1) {
                    float o[P];
#pragma unroll
                    for (int i = 0; i < P; ++i) o[i] = (float)u[i];
                    if (p.n_remote == 0) store_row<P>(p.obs + (size_t)k * p.B * p.N + off, o);
                    else store_row<P>(s_obs[threadIdx.x >> 5] + lane * P, o);     // staged, written out below
                } else {
                    float *orow = p.obs + ((size_t)k * p.B + env) * p.obs_len;
                    const int first = p.obs_stride / 2;
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const int idx = l * P + i - first;
                        if (idx >= 0 && idx % p.obs_stride == 0) {
                            orow[idx / p.obs_stride] = (float)u[i];
                            for (int q = 0; q < p.n_remote; ++q)
                                *remote_ptr(orow + idx / p.obs_stride, p.remote_delta[q]) = (float)u[i];
                        }
                    }
                }
            }
            if (l == 0) {
                if (badmask & grp) {
                    was_bad = true;
                    p.nonfinite[env] = 1;
                }
                const double rv = -(tot * p.inv_N) * p.inv_cfg_steps;
                const uint8_t tv = ts >= p.max_episode_steps ? 1 : 0, bv = was_bad ? 1 : 0;
                if (p.reward != nullptr) p.reward[kb] = rv;
                if (p.truncated != nullptr) p.truncated[kb] = tv;
                if (p.step != nullptr) p.step[kb] = ts;
                if (p.nonfinite_out != nullptr) p.nonfinite_out[kb] = bv;
                for (int q = 0; q < p.n_remote; ++q) {     // gather mode: all four outputs are present
                    const long long d = p.remote_delta[q];
                    *remote_ptr(p.reward + kb, d) = rv;
                    *remote_ptr(p.truncated + kb, d) = tv;
                    *remote_ptr(p.step + kb, d) = ts;
                    *remote_ptr(p.nonfinite_out + kb, d) = bv;
                }
            }
        }
        if (p.n_remote > 0 && p.obs != nullptr && p.obs_stride <= 1) {
            // coalesced write-out of the staged rows: local copy + every peer's copy
            const unsigned actmask = __ballot_sync(kFullMask, active);
            __syncwarp();
            float *wbase = p.obs + ((size_t)k * p.B + (size_t)warp * p.envs_per_warp) * p.N;
            const float *srow = s_obs[threadIdx.x >> 5];
            const int nfl = p.envs_per_warp * p.N;              // floats of this warp's rows
            if constexpr (P % 4 == 0) {
                for (int i4 = lane; i4 * 4 < nfl; i4 += 32) {
                    if (!((actmask >> ((i4 * 4) / P)) & 1u)) continue;
                    const float4 v = *reinterpret_cast<const float4 *>(srow + i4 * 4);
                    *reinterpret_cast<float4 *>(wbase + i4 * 4) = v;      // (local copy: visible here without a trip to the switch)
                    if (p.mc_delta != 0) {
                        multimem_st(remote_ptr(wbase + i4 * 4, p.mc_delta), v);
                    } else {
                        for (int q = 0; q < p.n_remote; ++q)
                            *reinterpret_cast<float4 *>(remote_ptr(wbase + i4 * 4, p.remote_delta[q])) = v;
                    }
                }
            } else {
                for (int i = lane; i < nfl; i += 32) {
                    if (!((actmask >> (i / P)) & 1u)) continue;
                    const float v = srow[i];
                    wbase[i] = v;
                    if (p.mc_delta != 0) {
                        multimem_st(remote_ptr(wbase + i, p.mc_delta), v);
                    } else {
                        for (int q = 0; q < p.n_remote; ++q) *remote_ptr(wbase + i, p.remote_delta[q]) = v;
                    }
                }
            }
            __syncwarp();
        }
    }
    if (p.n_remote > 0 && !p.skip_fence) __threadfence_system();   // peer stores performed before the launch retires

    if (active) {
        store_row<P>(ug, u);
        if (l == 0) p.timestep[env] = p.reset_timestep ? 0 : ts;
    }
}

}  // namespace ks
