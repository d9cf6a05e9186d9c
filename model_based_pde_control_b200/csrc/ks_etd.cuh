// ks_etd.cuh -- spectral ETDRK4 control-period kernel (sm_100a), the solver the north star names.
//
// NOT the reference's scheme: pdegym/kuramoto/kuramoto.py:83-90,118-129 is finite differences +
// classic RK4 (ks_kernels.cuh reproduces that one to 2e-14).  This kernel integrates the same
// equation  u_t = -u_xxxx - u_xx - 1/2 (u^2)_x + phi  (kuramoto.py:127) with the exponential
// time-differencing Runge-Kutta scheme of Cox & Matthews (2002) in Fourier space, coefficients by
// Kassam & Trefethen's (2005) contour integrals precomputed on the host (ks_api.cu).  Everything
// around the time stepper -- float32 jet forcing (transforms.py:262-265), reward of the pre-step
// state averaged over the sub-steps (kuramoto.py:82-84,96), timestep / truncation / float32
// observation (kuramoto.py:92-98) -- is the reference's.
//
// Layout (N = 64 R, R = 1, 2, 4; described for R = 1 first):
//   * TWO environments share one complex transform: z = u_a + i u_b.  Every spectral operation of
//     ETDRK4 is a multiplication by a real-kernel multiplier (E, E2, Q, f1..f3 real and even in k,
//     the derivative factor i*g(k) odd and imaginary), so the packed spectrum Z = U_a + i U_b is
//     evolved as it is and never unpacked; the two fields separate trivially in physical space
//     (real / imaginary part), where the nonlinearity and the reward live.
//   * a pair occupies 8 adjacent lanes x 8 complex registers.  Physical layout: lane l, register r
//     holds x-index n = l + 8 r; spectral layout: lane j, register m holds wavenumber index
//     k = 8 m + j.  A 64-point FFT is the four-step algorithm: radix-8 butterflies in registers,
//     twiddle by W64^(l r), 8x8 transpose inside the 8-lane group, radix-8 butterflies in
//     registers.  The inverse runs the same steps backwards, so no bit reversal is ever needed.
//   * the transpose goes through a warp-private, padded shared-memory tile (128-bit, conflict-free
//     both ways, only __syncwarp); the per-wavenumber tables live in shared memory as well; the
//     state, the stage values, the twiddles and the period's forcing spectrum in registers.
//   * N = 128 / 256 (R = 2 / 4): a pair occupies 8R lanes, x-index n = pl + 8R r.  After the four-step
//     core (whose transpose then runs inside the 8 lanes that share pl mod R) the R lanes that differ
//     in pl mod R are combined by one / two radix-2 butterfly stages with warp shuffles (twiddles
//     W_8R^(m1 s) before, a -i rotation between the stages); the spectral layout becomes lane
//     (m1', k2 = pl / R), register s  <->  k = 64 bitrev(m1') + 8 s + k2, which only the
//     table permutation at kernel start needs to know.
//   * HBM is touched at control-period boundaries only (state in, state / observation / reward out).
#pragma once

#include "ks_kernels.cuh"

namespace ks {

#ifndef KS_ETD_MIN_BLOCKS
#define KS_ETD_MIN_BLOCKS 2
#endif
constexpr int kEtdN = 64;          // grid points of the base layout; the kernel handles N = 64 R, R = 1, 2, 4
// Per-wavenumber tables, each [N] in natural FFT order (host-precomputed, ks_api.cu).  The kernel
// carries the nonlinear terms pre-multiplied by Q (N~ = Q N), which removes Q from the stage
// formulas; the final combination then needs f1/Q, 2 f2/Q, f3/Q (Q = h phi_1(hL/2)/2 > 0).
constexpr int kEtdTables = 9;
enum { kTabE = 0, kTabE2, kTabR1 /* f1/Q */, kTabR22 /* 2 f2/Q */, kTabR3 /* f3/Q */, kTabG /* Q g / N */,
       kTabQN /* Q / N */, kTabK2 /* -k^2 (even-derivative k) */, kTabKN /* k / N (odd-derivative k) */ };

template <typename T>
struct __align__(2 * sizeof(T)) C2 {
    T x, y;
};

// ---------------------------------------------------------------------------------------------
// 8-point DFT in registers, forward sign (W8 = exp(-2 pi i / 8)), natural order in and out.
// 52 floating-point instructions (the two 1/sqrt(2) twiddles are folded into FMAs).
// The inverse transform is the same routine with the roles of x and y exchanged.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void fft8(T (&x)[8], T (&y)[8])
{
    const T c = T(0.70710678118654752440084436210485);
    const T s0x = x[0] + x[4], s0y = y[0] + y[4], d0x = x[0] - x[4], d0y = y[0] - y[4];
    const T s1x = x[1] + x[5], s1y = y[1] + y[5], d1x = x[1] - x[5], d1y = y[1] - y[5];
    const T s2x = x[2] + x[6], s2y = y[2] + y[6], d2x = x[2] - x[6], d2y = y[2] - y[6];
    const T s3x = x[3] + x[7], s3y = y[3] + y[7], d3x = x[3] - x[7], d3y = y[3] - y[7];
    // even outputs: 4-point DFT of s
    {
        const T t0x = s0x + s2x, t0y = s0y + s2y, t1x = s0x - s2x, t1y = s0y - s2y;
        const T t2x = s1x + s3x, t2y = s1y + s3y;
        const T t3x = s1y - s3y, t3y = s3x - s1x;             // (s1 - s3) * (-i)
        x[0] = t0x + t2x; y[0] = t0y + t2y;
        x[4] = t0x - t2x; y[4] = t0y - t2y;
        x[2] = t1x + t3x; y[2] = t1y + t3y;
        x[6] = t1x - t3x; y[6] = t1y - t3y;
    }
    // odd outputs: 4-point DFT of e_j = d_j W8^j
    {
        const T u0x = d0x + d2y, u0y = d0y - d2x;             // e0 + e2,  e2 = d2 * (-i)
        const T u1x = d0x - d2y, u1y = d0y + d2x;             // e0 - e2
        const T p1x = d1x + d1y, p1y = d1y - d1x;             // d1 (1 - i)      = e1 / c
        const T p3x = d3y - d3x, q3y = d3x + d3y;             // d3 (-1 - i)     = (p3x, -q3y) = e3 / c
        const T Px = p1x + p3x, Py = p1y - q3y;               // (e1 + e3) / c
        const T Qx = p1y + q3y, Qy = p3x - p1x;               // (e1 - e3)(-i) / c
        x[1] = fma_t<T>(c, Px, u0x);  y[1] = fma_t<T>(c, Py, u0y);
        x[5] = fma_t<T>(-c, Px, u0x); y[5] = fma_t<T>(-c, Py, u0y);
        x[3] = fma_t<T>(c, Qx, u1x);  y[3] = fma_t<T>(c, Qy, u1y);
        x[7] = fma_t<T>(-c, Qx, u1x); y[7] = fma_t<T>(-c, Qy, u1y);
    }
}

// (x + i y)[r] *= tw[r], r = 1..7 (tw[0] = 1).  Exchanging x and y multiplies by conj(tw).
template <typename T>
__device__ __forceinline__ void twiddle8(T (&x)[8], T (&y)[8], const T (&twx)[8], const T (&twy)[8])
{
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        const T a = x[r], b = y[r];
        x[r] = fma_t<T>(a, twx[r], -(b * twy[r]));
        y[r] = fma_t<T>(a, twy[r], b * twx[r]);
    }
}

// 8x8 transpose inside an 8-lane group: afterwards lane j, register m holds what lane m had in
// register j.  `tile` = this group's 72-element padded tile (row stride 9 keeps both the 128-bit
// stores and the 128-bit loads conflict-free).
template <typename T>
__device__ __forceinline__ void transpose8(T (&x)[8], T (&y)[8], C2<T> *tile, int l)
{
#pragma unroll
    for (int r = 0; r < 8; ++r) tile[r * 9 + l] = C2<T>{x[r], y[r]};
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const C2<T> v = tile[l * 9 + m];
        x[m] = v.x;
        y[m] = v.y;
    }
    __syncwarp();
}

// R-point DFT across the R adjacent lanes that share pl / R (R = 2, 4), radix-2 butterflies with
// warp shuffles, decimation in frequency: the result index ends up bit-reversed in the lane
// position, which the host-side table permutation accounts for.  `inv_core` is the mirrored
// sequence; called with x and y exchanged it is the exact inverse (up to the factor R).
template <typename T>
__device__ __forceinline__ void lane_butterfly(T (&x)[8], T (&y)[8], int xor_mask, T sgn)
{
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const T px = __shfl_xor_sync(kFullMask, x[r], xor_mask), py = __shfl_xor_sync(kFullMask, y[r], xor_mask);
        x[r] = fma_t<T>(sgn, x[r], px);       // lower lane: a + b,  upper lane: a - b (own value is b)
        y[r] = fma_t<T>(sgn, y[r], py);
    }
}
template <typename T>
__device__ __forceinline__ void rotate_minus_i(T (&x)[8], T (&y)[8], bool on)
{
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const T a = x[r], b = y[r];
        x[r] = on ? b : a;                     // (a + i b)(-i) = b - i a
        y[r] = on ? -a : b;
    }
}
template <typename T, int R>
__device__ __forceinline__ void lanes_fwd(T (&x)[8], T (&y)[8], int m1)
{
    if constexpr (R == 2) {
        lane_butterfly<T>(x, y, 1, (m1 & 1) ? T(-1) : T(1));
    } else if constexpr (R == 4) {
        lane_butterfly<T>(x, y, 2, (m1 & 2) ? T(-1) : T(1));
        rotate_minus_i<T>(x, y, m1 == 3);     // W4^(m1 & 1) on the difference lanes
        lane_butterfly<T>(x, y, 1, (m1 & 1) ? T(-1) : T(1));
    }
}
template <typename T, int R>
__device__ __forceinline__ void lanes_inv_core(T (&x)[8], T (&y)[8], int m1)
{
    if constexpr (R == 2) {
        lane_butterfly<T>(x, y, 1, (m1 & 1) ? T(-1) : T(1));
    } else if constexpr (R == 4) {
        lane_butterfly<T>(x, y, 1, (m1 & 1) ? T(-1) : T(1));
        rotate_minus_i<T>(x, y, m1 == 3);
        lane_butterfly<T>(x, y, 2, (m1 & 2) ? T(-1) : T(1));
    }
}
// (x + i y)[s] *= W_8R^(m1 s), s = 1..7, from the small shared table twb[m1][s]
template <typename T>
__device__ __forceinline__ void twiddle_lanes(T (&x)[8], T (&y)[8], const C2<T> *twb_row)
{
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        const C2<T> w = twb_row[r];
        const T a = x[r], b = y[r];
        x[r] = fma_t<T>(a, w.x, -(b * w.y));
        y[r] = fma_t<T>(a, w.y, b * w.x);
    }
}

// What a lane needs for its share of a transform: twiddles W_N^(pl r) (registers), its 8-lane
// transpose group's tile and its position `l` = pl / R in that group, and for R > 1 its row
// W_8R^(m1 s) of the small shared twiddle table with m1 = pl % R.
template <typename T>
struct FftCtx {
    T twx[8], twy[8];
    C2<T> *tile;
    const C2<T> *twb_row;
    int l, m1;
};

// physical (lane pl, reg r: n = pl + 8R r)  ->  spectral, unnormalised
template <typename T, int R>
__device__ __forceinline__ void fft64(T (&x)[8], T (&y)[8], const FftCtx<T> &c)
{
    fft8<T>(x, y);
    twiddle8<T>(x, y, c.twx, c.twy);
    transpose8<T>(x, y, c.tile, c.l);
    fft8<T>(x, y);
    if constexpr (R > 1) {
        twiddle_lanes<T>(x, y, c.twb_row);
        lanes_fwd<T, R>(x, y, c.m1);
    }
}
// spectral -> physical, unnormalised (x/y exchanged = conjugated kernel)
template <typename T, int R>
__device__ __forceinline__ void ifft64(T (&x)[8], T (&y)[8], const FftCtx<T> &c)
{
    if constexpr (R > 1) {
        lanes_inv_core<T, R>(y, x, c.m1);
        twiddle_lanes<T>(y, x, c.twb_row);
    }
    fft8<T>(y, x);
    transpose8<T>(x, y, c.tile, c.l);
    twiddle8<T>(y, x, c.twx, c.twy);
    fft8<T>(y, x);
}

// Table values for this lane's registers m = 2h, 2h+1 (k = 8 m + j), one 2-element vector load.
// Layout in shared memory: [table][h][lane pl of the pair], LP = 8R lanes.
template <typename T, int LP>
__device__ __forceinline__ C2<T> tab2(const C2<T> *tab, int which, int h, int pl)
{
    return tab[(which * 4 + h) * LP + pl];
}

// Spectral right-hand side without the linear part, pre-multiplied by Q, in place:
//   w <- i (Q g / N) * FFT( (Re/Im IFFT(w))^2 ) + Q phi_hat / N        (both packed fields at once)
// With FIRST the squares of the physical values are the pre-step reward terms (kuramoto.py:84).
template <typename T>
__device__ __forceinline__ T sum8(const T (&a)[8])
{
    return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

// Reward terms of the pre-step state, per lane partial sums for the two envs of the pair:
//   L2 mode:           a = sum u^2                                      (kuramoto.py:64-65)
//   dissipation mode:  a = sum uxx^2,  b = sum ux^2,  c = sum u*phi     (kuramoto.py:67-70), where
//     -- literally as in the reference -- ux is the derivative of u^2 (rhs() differentiates u**2,
//     kuramoto.py:120-122); both derivatives are spectral here.
template <typename T>
struct EtdReward {
    T a[2], b[2], c[2];
};

template <typename T, int R, bool FIRST, int RMODE>
__device__ __forceinline__ void nonlinear(T (&wx)[8], T (&wy)[8], const FftCtx<T> &ctx, const C2<T> *tab,
                                          const T (&phx)[8], const T (&phy)[8], int pl, EtdReward<T> &rw,
                                          const float (&pfx)[8], const float (&pfy)[8])
{
    ifft64<T, R>(wx, wy, ctx);
    if constexpr (FIRST && RMODE == kRewardDissipation) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {          // power term u * phi with the float32 jets
            rw.c[0] = fma_t<T>(wx[r], T(pfx[r]), rw.c[0]);
            rw.c[1] = fma_t<T>(wy[r], T(pfy[r]), rw.c[1]);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        wx[r] = wx[r] * wx[r];
        wy[r] = wy[r] * wy[r];
    }
    if constexpr (FIRST && RMODE == kRewardL2) {
        rw.a[0] += sum8<T>(wx);
        rw.a[1] += sum8<T>(wy);
    }
    fft64<T, R>(wx, wy, ctx);
    if constexpr (FIRST && RMODE == kRewardDissipation) {
        // (u^2)_x = IFFT(i k FFT(u^2)) / N for both fields at once
        T dx[8], dy[8];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const C2<T> kn = tab2<T, 8 * R>(tab, kTabKN, h, pl);
            dx[2 * h] = -kn.x * wy[2 * h];         dy[2 * h] = kn.x * wx[2 * h];
            dx[2 * h + 1] = -kn.y * wy[2 * h + 1]; dy[2 * h + 1] = kn.y * wx[2 * h + 1];
        }
        ifft64<T, R>(dx, dy, ctx);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            rw.b[0] = fma_t<T>(dx[r], dx[r], rw.b[0]);
            rw.b[1] = fma_t<T>(dy[r], dy[r], rw.b[1]);
        }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const C2<T> g = tab2<T, 8 * R>(tab, kTabG, h, pl);
        const T re0 = wx[2 * h], im0 = wy[2 * h], re1 = wx[2 * h + 1], im1 = wy[2 * h + 1];
        wx[2 * h] = fma_t<T>(-g.x, im0, phx[2 * h]);
        wy[2 * h] = fma_t<T>(g.x, re0, phy[2 * h]);
        wx[2 * h + 1] = fma_t<T>(-g.y, im1, phx[2 * h + 1]);
        wy[2 * h + 1] = fma_t<T>(g.y, re1, phy[2 * h + 1]);
    }
}

struct EtdParams {
    Params p;              // same buffers / geometry as the finite-difference kernel
    const void *tables;    // [kEtdTables][N] T, natural FFT order (host-precomputed, ks_api.cu)
};

// ---------------------------------------------------------------------------------------------
// The spectral control-period kernel: K periods x cfg_steps ETDRK4 steps, 8/R envs per warp.
// ---------------------------------------------------------------------------------------------
template <typename T, int R, int RMODE>
__global__ void __launch_bounds__(kBlockThreads, KS_ETD_MIN_BLOCKS) ks_etd_kernel(const EtdParams ep)
{
    static_assert(R == 1 || R == 2 || R == 4, "N = 64 R with R = 1, 2, 4");
    const Params &p = ep.p;
    constexpr int N = kEtdN * R;
    constexpr int LP = 8 * R;                       // lanes per env pair
    constexpr int PPW = 4 / R;                      // env pairs per warp
    constexpr int kWarps = kBlockThreads / 32;
    constexpr int kBlk = R == 1 ? 72 : 72 + 8 / R;  // tile stride of an 8-lane transpose group (bank-conflict free)
    __shared__ C2<T> s_tab[kEtdTables * 4 * LP];
    __shared__ C2<T> s_tile[kWarps][4 * kBlk];
    __shared__ C2<T> s_twb[R * 8];                  // W_8R^(m1 s)

    // per-wavenumber tables -> shared memory.  Slot (table, h, pl, q) belongs to lane pl, register
    // s = 2h + q, i.e. wavenumber index k = 64 bitrev_R(pl % R) + 8 s + pl / R.
    {
        const T *src = static_cast<const T *>(ep.tables);
        for (int i = threadIdx.x; i < kEtdTables * 4 * LP * 2; i += kBlockThreads) {
            const int q = i & 1, pl_ = (i >> 1) % LP, h = ((i >> 1) / LP) & 3, t = (i >> 1) / (4 * LP);
            const int m1p = pl_ % R, k2 = pl_ / R;
            const int j1 = R == 4 ? ((m1p & 1) << 1 | (m1p >> 1)) : m1p;
            const int k = 64 * j1 + 8 * (2 * h + q) + k2;
            reinterpret_cast<T *>(&s_tab[(t * 4 + h) * LP + pl_])[q] = src[t * N + k];
        }
        for (int i = threadIdx.x; i < R * 8; i += kBlockThreads) {
            double sn, cs;
            sincospi(-2.0 * (double)((i >> 3) * (i & 7)) / (double)LP, &sn, &cs);
            s_twb[i] = C2<T>{(T)cs, (T)sn};
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * kBlockThreads + threadIdx.x) >> 5;
    const int grp = lane / LP, pl = lane % LP;      // pair slot in the warp, lane inside the pair
    const int envA = (warp * PPW + grp) * 2, envB = envA + 1;
    bool actA = envA < p.B, actB = envB < p.B;
    if (p.mask != nullptr) {
        actA = actA && p.mask[envA] != 0;
        actB = actB && p.mask[envB] != 0;
    }
    if (__ballot_sync(kFullMask, actA || actB) == 0u) return;   // warp-uniform (after the only __syncthreads)

    const C2<T> *tab = s_tab;
    FftCtx<T> ctx;
    ctx.m1 = pl % R;
    ctx.l = pl / R;
    ctx.tile = &s_tile[wib][(grp * R + ctx.m1) * kBlk];
    ctx.twb_row = &s_twb[ctx.m1 * 8];
    // twiddles W_N^(pl r) = exp(-2 pi i pl r / N)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double sn, cs;
        sincospi(-2.0 * (double)(pl * r) / (double)N, &sn, &cs);
        ctx.twx[r] = (T)cs;
        ctx.twy[r] = (T)sn;
    }

    // physical state of the pair: x = env A, y = env B (idle slots integrate zeros, never stored)
    T ux[8], uy[8];
    T *ua = static_cast<T *>(p.u) + (size_t)(actA ? envA : 0) * N + pl;
    T *ub = static_cast<T *>(p.u) + (size_t)(actB ? envB : 0) * N + pl;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        ux[r] = actA ? ua[LP * r] : T(0);
        uy[r] = actB ? ub[LP * r] : T(0);
    }
    int tsA = actA ? p.timestep[envA] : 0, tsB = actB ? p.timestep[envB] : 0;
    bool badA = actA ? p.nonfinite[envA] != 0 : false, badB = actB ? p.nonfinite[envB] != 0 : false;
    const T invN = T(1.0 / N);

    for (int k = 0; k < p.K; ++k) {
        // ---- jet forcing of this period: float32 FMA chain (transforms.py:262-265), then its spectrum
        T phx[8], phy[8];      // Q phi_hat / N, kept in registers for the whole period
        [[maybe_unused]] float pfx[8], pfy[8];     // the jets in physical space (dissipation reward: u * phi)
        {
            T (&fx)[8] = phx, (&fy)[8] = phy;
            if (p.phi != nullptr) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    fx[r] = actA ? T(p.phi[(size_t)envA * N + pl + LP * r]) : T(0);
                    fy[r] = actB ? T(p.phi[(size_t)envB * N + pl + LP * r]) : T(0);
                }
            } else if (p.actions != nullptr) {
                const float *aA = p.actions + ((size_t)k * p.B + (actA ? envA : 0)) * p.J;
                const float *aB = p.actions + ((size_t)k * p.B + (actB ? envB : 0)) * p.J;
                float fa[8], fb[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) fa[r] = fb[r] = 0.0f;
                for (int j = 0; j < p.J; ++j) {
                    const float a = __ldg(aA + j), b = __ldg(aB + j);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const float f = __ldg(p.F + (size_t)j * N + pl + LP * r);
                        fa[r] = __fmaf_rn(a, f, fa[r]);
                        fb[r] = __fmaf_rn(b, f, fb[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    fx[r] = actA ? T(fa[r]) : T(0);
                    fy[r] = actB ? T(fb[r]) : T(0);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 8; ++r) fx[r] = fy[r] = T(0);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) { pfx[r] = (float)fx[r]; pfy[r] = (float)fy[r]; }
            fft64<T, R>(fx, fy, ctx);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> qn = tab2<T, LP>(tab, kTabQN, h, pl);
                fx[2 * h] *= qn.x; fy[2 * h] *= qn.x;
                fx[2 * h + 1] *= qn.y; fy[2 * h + 1] *= qn.y;
            }
        }

        // ---- spectrum of the state, normalised: v = FFT(u) / N.  An idle slot (odd batch, masked
        // env) restarts from exact zeros every period, so that its rounding-level leakage into the
        // partner does not depend on how many periods a launch covers (rollout == repeated steps).
        T vx[8], vy[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { vx[r] = actA ? ux[r] : T(0); vy[r] = actB ? uy[r] : T(0); }
        fft64<T, R>(vx, vy, ctx);
#pragma unroll
        for (int m = 0; m < 8; ++m) { vx[m] *= invN; vy[m] *= invN; }

        // ---- cfg_steps ETDRK4 steps (Cox & Matthews 2002, eqs. 26-29), with N~ = Q N:
        //   a = E2 v + N~v          b = E2 v + N~a = (a - N~v) + N~a          c = E2 a + 2 N~b - N~v
        //   v' = E v + (f1/Q) N~v + (2 f2/Q) (N~a + N~b) + (f3/Q) N~c
        EtdReward<T> rw{{T(0), T(0)}, {T(0), T(0)}, {T(0), T(0)}};
        for (int s = 0; s < p.cfg_steps; ++s) {
            T nx[8], ny[8], ax[8], ay[8], wx[8], wy[8];
            if constexpr (RMODE == kRewardDissipation) {
                // uxx = IFFT(-k^2 v) of the pre-step state, both fields at once
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const C2<T> k2 = tab2<T, LP>(tab, kTabK2, h, pl);
                    wx[2 * h] = k2.x * vx[2 * h];         wy[2 * h] = k2.x * vy[2 * h];
                    wx[2 * h + 1] = k2.y * vx[2 * h + 1]; wy[2 * h + 1] = k2.y * vy[2 * h + 1];
                }
                ifft64<T, R>(wx, wy, ctx);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    rw.a[0] = fma_t<T>(wx[r], wx[r], rw.a[0]);
                    rw.a[1] = fma_t<T>(wy[r], wy[r], rw.a[1]);
                }
            }
            // N~v
#pragma unroll
            for (int m = 0; m < 8; ++m) { nx[m] = vx[m]; ny[m] = vy[m]; }
            nonlinear<T, R, true, RMODE>(nx, ny, ctx, tab, phx, phy, pl, rw, pfx, pfy);
            // a = E2 v + N~v
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> e2 = tab2<T, LP>(tab, kTabE2, h, pl);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int m = 2 * h + q;
                    const T c = q ? e2.y : e2.x;
                    wx[m] = ax[m] = fma_t<T>(c, vx[m], nx[m]);
                    wy[m] = ay[m] = fma_t<T>(c, vy[m], ny[m]);
                }
            }
            // N~a
            nonlinear<T, R, false, RMODE>(wx, wy, ctx, tab, phx, phy, pl, rw, pfx, pfy);
            // b = (a - N~v) + N~a (-> w);   a <- E2 a - N~v;   v <- E v + (f1/Q) N~v + (2 f2/Q) N~a
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> e2 = tab2<T, LP>(tab, kTabE2, h, pl), e = tab2<T, LP>(tab, kTabE, h, pl);
                const C2<T> r1 = tab2<T, LP>(tab, kTabR1, h, pl), r22 = tab2<T, LP>(tab, kTabR22, h, pl);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int m = 2 * h + q;
                    const T ce2 = q ? e2.y : e2.x, ce = q ? e.y : e.x, c1 = q ? r1.y : r1.x, c22 = q ? r22.y : r22.x;
                    const T nax = wx[m], nay = wy[m];
                    wx[m] = (ax[m] - nx[m]) + nax;
                    wy[m] = (ay[m] - ny[m]) + nay;
                    ax[m] = fma_t<T>(ce2, ax[m], -nx[m]);
                    ay[m] = fma_t<T>(ce2, ay[m], -ny[m]);
                    vx[m] = fma_t<T>(c22, nax, fma_t<T>(c1, nx[m], ce * vx[m]));
                    vy[m] = fma_t<T>(c22, nay, fma_t<T>(c1, ny[m], ce * vy[m]));
                }
            }
            // N~b
            nonlinear<T, R, false, RMODE>(wx, wy, ctx, tab, phx, phy, pl, rw, pfx, pfy);
            // v += (2 f2/Q) N~b;   c = (E2 a - N~v) + 2 N~b (-> w)
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> r22 = tab2<T, LP>(tab, kTabR22, h, pl);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int m = 2 * h + q;
                    const T c22 = q ? r22.y : r22.x;
                    vx[m] = fma_t<T>(c22, wx[m], vx[m]);
                    vy[m] = fma_t<T>(c22, wy[m], vy[m]);
                    wx[m] = fma_t<T>(T(2), wx[m], ax[m]);
                    wy[m] = fma_t<T>(T(2), wy[m], ay[m]);
                }
            }
            // N~c;   v += (f3/Q) N~c
            nonlinear<T, R, false, RMODE>(wx, wy, ctx, tab, phx, phy, pl, rw, pfx, pfy);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> r3 = tab2<T, LP>(tab, kTabR3, h, pl);
                vx[2 * h] = fma_t<T>(r3.x, wx[2 * h], vx[2 * h]);
                vy[2 * h] = fma_t<T>(r3.x, wy[2 * h], vy[2 * h]);
                vx[2 * h + 1] = fma_t<T>(r3.y, wx[2 * h + 1], vx[2 * h + 1]);
                vy[2 * h + 1] = fma_t<T>(r3.y, wy[2 * h + 1], vy[2 * h + 1]);
            }
        }

        // ---- back to physical space
#pragma unroll
        for (int m = 0; m < 8; ++m) { ux[m] = vx[m]; uy[m] = vy[m]; }
        ifft64<T, R>(ux, uy, ctx);

        // ---- period epilogue: reward, flags, observation (kuramoto.py:92-98)
        double ra = (double)rw.a[0] + (double)rw.b[0] + (double)rw.c[0];
        double rb = (double)rw.a[1] + (double)rw.b[1] + (double)rw.c[1];
        bool bA = false, bB = false;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            bA |= !(fabs((double)ux[r]) <= 1.7976931348623157e308);
            bB |= !(fabs((double)uy[r]) <= 1.7976931348623157e308);
        }
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) {  // fixed-order reduction over the lanes of the pair
            ra += __shfl_xor_sync(kFullMask, ra, o);
            rb += __shfl_xor_sync(kFullMask, rb, o);
        }
        const unsigned gmask = (LP == 32 ? kFullMask : ((1u << LP) - 1u)) << (grp * LP);
        const bool anyA = (__ballot_sync(kFullMask, bA) & gmask) != 0u;
        const bool anyB = (__ballot_sync(kFullMask, bB) & gmask) != 0u;
        tsA += 1;
        tsB += 1;
        if (p.obs != nullptr) {
            if (p.obs_stride <= 1) {
                float *oa = p.obs + ((size_t)k * p.B + envA) * N + pl;
                float *ob = p.obs + ((size_t)k * p.B + envB) * N + pl;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (actA) oa[LP * r] = (float)ux[r];
                    if (actB) ob[LP * r] = (float)uy[r];
                }
                if (p.n_remote > 0) {
                    // gather mode: stage the warp's rows (512 floats, contiguous in the batch) in the
                    // transpose tile and mirror them to the peers as coalesced 16-byte-per-lane stores
                    float *stg = reinterpret_cast<float *>(&s_tile[wib][0]);
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        stg[(2 * grp) * N + pl + LP * r] = (float)ux[r];
                        stg[(2 * grp + 1) * N + pl + LP * r] = (float)uy[r];
                    }
                    __syncwarp();
                    float *wbase = p.obs + ((size_t)k * p.B + (size_t)warp * (2 * PPW)) * N;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i4 = j * 32 + lane, e = warp * (2 * PPW) + (i4 * 4) / N;     // env that owns this float4
                        if (e >= p.B || (p.mask != nullptr && p.mask[e] == 0)) continue;
                        const float4 v = *reinterpret_cast<const float4 *>(stg + i4 * 4);
                        for (int q = 0; q < p.n_remote; ++q)
                            *reinterpret_cast<float4 *>(remote_ptr(wbase + i4 * 4, p.remote_delta[q])) = v;
                    }
                    __syncwarp();
                }
            } else {
                const int first = p.obs_stride / 2;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int idx = pl + LP * r - first;
                    if (idx >= 0 && idx % p.obs_stride == 0) {
                        float *da = p.obs + ((size_t)k * p.B + envA) * p.obs_len + idx / p.obs_stride;
                        float *db = p.obs + ((size_t)k * p.B + envB) * p.obs_len + idx / p.obs_stride;
                        if (actA) *da = (float)ux[r];
                        if (actB) *db = (float)uy[r];
                        for (int q = 0; q < p.n_remote; ++q) {
                            if (actA) *remote_ptr(da, p.remote_delta[q]) = (float)ux[r];
                            if (actB) *remote_ptr(db, p.remote_delta[q]) = (float)uy[r];
                        }
                    }
                }
            }
        }
        if (pl == 0) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool act = e ? actB : actA;
                if (!act) continue;
                const int env = e ? envB : envA, ts = e ? tsB : tsA;
                bool &bad = e ? badB : badA;
                if (e ? anyB : anyA) { bad = true; p.nonfinite[env] = 1; }
                const size_t kb = (size_t)k * p.B + env;
                const double rv = -((e ? rb : ra) * p.inv_N) * p.inv_cfg_steps;
                const uint8_t tv = ts >= p.max_episode_steps ? 1 : 0, bv = bad ? 1 : 0;
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
    }
    if (p.n_remote > 0 && !p.skip_fence) __threadfence_system();   // peer stores performed before the launch retires

#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (actA) ua[LP * r] = ux[r];
        if (actB) ub[LP * r] = uy[r];
    }
    if (pl == 0) {
        if (actA) p.timestep[envA] = p.reset_timestep ? 0 : tsA;
        if (actB) p.timestep[envB] = p.reset_timestep ? 0 : tsB;
    }
}

}  // namespace ks
