// ks_etd16.cuh -- small-batch layout of the spectral ETDRK4 control-period kernel (N = 64).
//
// Same algorithm, tables, reward and env semantics as ks_etd.cuh (read that header first); what
// differs is how a pair of envs is spread over the warp.  ks_etd.cuh gives a pair 8 lanes x 8
// complex registers: 8 envs per warp, 255 registers, so a batch of 4096 envs is 512 warps -- one
// warp on 512 of the B200's 592 SM sub-partitions, nothing to overlap its dependent FFT stages
// with.  Here a pair occupies 16 lanes x 4 complex registers: 4 envs per warp, 168 registers, twice
// the warps, and every warp's dependent chain is half as long: 23.3 us instead of 29.3 us per
// control period for a warp that has its sub-partition to itself.  ks_create therefore picks this
// layout for batches of up to 4 x 592 = 2368 envs (ks_api.cu, choose_etd_regs_per_lane).  Beyond
// that the 8-lane layout wins (measured, profiles/round2_sweep_etd_layouts.jsonl: 4096 envs 29.3
// against 34.9 us, 65 536 envs 0.271 against 0.382 ms): this layout needs two exchanges per transform
// instead of one, i.e. twice the shared-memory crossbar traffic per env, and the crossbar
// (128 B/clk/SM) is the co-bottleneck of the spectral solver.
//
// 64-point FFT on 16 lanes x 4 registers: three radix-4 stages in registers with two 4x4
// exchanges between them.  With n = n0 + 4 n1 + 16 n2 and k = k2 + 4 k1 + 16 k0,
//     W64^(nk) = W4^(n2 k2) . W64^((n0 + 4 n1) k2) . W4^(n1 k1) . W16^(n0 k1) . W4^(n0 k0):
//   physical:  lane a = n0 + 4 n1, register n2          (x-index a + 16 r: coalesced rows)
//   S1  radix-4 over the registers (n2 -> k2);   T1  register k2 *= W64^(a k2)
//   X1  exchange inside the 4 lanes sharing n0:  lane (n0,n1) reg k2 -> lane (n0,k2) reg n1
//   S2  radix-4 (n1 -> k1);                      T2  register k1 *= W16^(n0 k1)
//   X2  exchange inside the 4 lanes sharing k2:  lane (n0,k2) reg k1 -> lane (k1,k2) reg n0
//   S3  radix-4 (n0 -> k0)
//   spectral:  lane c = k1 + 4 k2, register k0  <->  k = k2 + 4 k1 + 16 k0
// The inverse runs the mirrored sequence with the real / imaginary arrays exchanged (conjugated
// kernels), so no bit reversal is ever needed; only the table permutation at kernel start knows
// the spectral order.  The exchanges go through a warp-private padded shared-memory tile whose
// strides make the 128-bit stores and loads of both directions bank-conflict free.
#pragma once

#include "ks_etd.cuh"

namespace ks {

#ifndef KS_ETD16_MIN_BLOCKS
#define KS_ETD16_MIN_BLOCKS 3
#endif

// 4-point DFT in registers, forward sign (W4 = -i), natural order in and out: 16 additions.
template <typename T>
__device__ __forceinline__ void fft4(T (&x)[4], T (&y)[4])
{
    const T s0x = x[0] + x[2], s0y = y[0] + y[2], d0x = x[0] - x[2], d0y = y[0] - y[2];
    const T s1x = x[1] + x[3], s1y = y[1] + y[3], d1x = x[1] - x[3], d1y = y[1] - y[3];
    x[0] = s0x + s1x; y[0] = s0y + s1y;
    x[2] = s0x - s1x; y[2] = s0y - s1y;
    x[1] = d0x + d1y; y[1] = d0y - d1x;       // d0 - i d1
    x[3] = d0x - d1y; y[3] = d0y + d1x;       // d0 + i d1
}

// (x + i y)[r] *= tw[r], r = 1..3 (tw[0] = 1).  Exchanging x and y multiplies by conj(tw).
template <typename T>
__device__ __forceinline__ void twiddle4(T (&x)[4], T (&y)[4], const T (&twx)[4], const T (&twy)[4])
{
#pragma unroll
    for (int r = 1; r < 4; ++r) {
        const T a = x[r], b = y[r];
        x[r] = fma_t<T>(a, twx[r], -(b * twy[r]));
        y[r] = fma_t<T>(a, twy[r], b * twx[r]);
    }
}

// Tile geometry (units of one complex element).  X1 addresses slot [n0][k2][n1] as
// n0 * 26 + k2 * 5 + n1, X2 addresses slot [k2][k1][n0] as k2 * 20 + k1 * 5 + n0: within every
// quarter-warp (the unit a 128-bit shared-memory access is served in) the eight lanes then touch
// eight different 16-byte bank groups, whether they store with the register index in the middle
// position or load with it in the last one.
constexpr int kTile16 = 104;     // 4 * 26 elements per pair

// `lo` = lane & 3, `hi` = (lane >> 2) & 3 of the pair-relative lane.  FWD: the transforms' forward
// direction (register index moves from the middle digit to the last one), else its inverse.
template <typename T, bool FWD>
__device__ __forceinline__ void exchange1(T (&x)[4], T (&y)[4], C2<T> *tile, int lo, int hi)
{
#pragma unroll
    for (int r = 0; r < 4; ++r) tile[FWD ? lo * 26 + r * 5 + hi : lo * 26 + hi * 5 + r] = C2<T>{x[r], y[r]};
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const C2<T> v = tile[FWD ? lo * 26 + hi * 5 + r : lo * 26 + r * 5 + hi];
        x[r] = v.x;
        y[r] = v.y;
    }
    __syncwarp();
}
template <typename T, bool FWD>
__device__ __forceinline__ void exchange2(T (&x)[4], T (&y)[4], C2<T> *tile, int lo, int hi)
{
#pragma unroll
    for (int r = 0; r < 4; ++r) tile[FWD ? hi * 20 + r * 5 + lo : hi * 20 + lo * 5 + r] = C2<T>{x[r], y[r]};
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const C2<T> v = tile[FWD ? hi * 20 + lo * 5 + r : hi * 20 + r * 5 + lo];
        x[r] = v.x;
        y[r] = v.y;
    }
    __syncwarp();
}

template <typename T>
struct Fft16Ctx {
    T t1x[4], t1y[4];     // W64^(a k2), a = pair-relative lane
    T t2x[4], t2y[4];     // W16^(n0 k1), n0 = a & 3
    C2<T> *tile;
    int lo, hi;
};

// physical (lane a, reg r: n = a + 16 r)  ->  spectral (lane c, reg k0: k = (c >> 2) + 4 (c & 3) + 16 k0), unnormalised
template <typename T>
__device__ __forceinline__ void fft64x16(T (&x)[4], T (&y)[4], const Fft16Ctx<T> &c)
{
    fft4<T>(x, y);
    twiddle4<T>(x, y, c.t1x, c.t1y);
    exchange1<T, true>(x, y, c.tile, c.lo, c.hi);
    fft4<T>(x, y);
    twiddle4<T>(x, y, c.t2x, c.t2y);
    exchange2<T, true>(x, y, c.tile, c.lo, c.hi);
    fft4<T>(x, y);
}
// spectral -> physical, unnormalised (x / y exchanged = conjugated kernels, mirrored order)
template <typename T>
__device__ __forceinline__ void ifft64x16(T (&x)[4], T (&y)[4], const Fft16Ctx<T> &c)
{
    fft4<T>(y, x);
    exchange2<T, false>(x, y, c.tile, c.lo, c.hi);
    twiddle4<T>(y, x, c.t2x, c.t2y);
    fft4<T>(y, x);
    exchange1<T, false>(x, y, c.tile, c.lo, c.hi);
    twiddle4<T>(y, x, c.t1x, c.t1y);
    fft4<T>(y, x);
}

// Table values for this lane's registers k0 = 2h, 2h+1: shared layout [table][h][lane c].
template <typename T>
__device__ __forceinline__ C2<T> tab16(const C2<T> *tab, int which, int h, int c)
{
    return tab[(which * 2 + h) * 16 + c];
}

template <typename T>
__device__ __forceinline__ T sum4(const T (&a)[4])
{
    return (a[0] + a[1]) + (a[2] + a[3]);
}

// w <- i (Q g / N) * FFT( (Re/Im IFFT(w))^2 ) + Q phi_hat / N   (see nonlinear<> in ks_etd.cuh)
template <typename T, bool FIRST, int RMODE>
__device__ __forceinline__ void nonlinear16(T (&wx)[4], T (&wy)[4], const Fft16Ctx<T> &ctx, const C2<T> *tab,
                                            const T (&phx)[4], const T (&phy)[4], int c, EtdReward<T> &rw,
                                            const float (&pfx)[4], const float (&pfy)[4])
{
    ifft64x16<T>(wx, wy, ctx);
    if constexpr (FIRST && RMODE == kRewardDissipation) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {          // power term u * phi with the float32 jets
            rw.c[0] = fma_t<T>(wx[r], T(pfx[r]), rw.c[0]);
            rw.c[1] = fma_t<T>(wy[r], T(pfy[r]), rw.c[1]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        wx[r] = wx[r] * wx[r];
        wy[r] = wy[r] * wy[r];
    }
    if constexpr (FIRST && RMODE == kRewardL2) {
        rw.a[0] += sum4<T>(wx);
        rw.a[1] += sum4<T>(wy);
    }
    fft64x16<T>(wx, wy, ctx);
    if constexpr (FIRST && RMODE == kRewardDissipation) {
        T dx[4], dy[4];                        // (u^2)_x = IFFT(i k FFT(u^2)) / N for both fields at once
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const C2<T> kn = tab16<T>(tab, kTabKN, h, c);
            dx[2 * h] = -kn.x * wy[2 * h];         dy[2 * h] = kn.x * wx[2 * h];
            dx[2 * h + 1] = -kn.y * wy[2 * h + 1]; dy[2 * h + 1] = kn.y * wx[2 * h + 1];
        }
        ifft64x16<T>(dx, dy, ctx);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            rw.b[0] = fma_t<T>(dx[r], dx[r], rw.b[0]);
            rw.b[1] = fma_t<T>(dy[r], dy[r], rw.b[1]);
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const C2<T> g = tab16<T>(tab, kTabG, h, c);
        const T re0 = wx[2 * h], im0 = wy[2 * h], re1 = wx[2 * h + 1], im1 = wy[2 * h + 1];
        wx[2 * h] = fma_t<T>(-g.x, im0, phx[2 * h]);
        wy[2 * h] = fma_t<T>(g.x, re0, phy[2 * h]);
        wx[2 * h + 1] = fma_t<T>(-g.y, im1, phx[2 * h + 1]);
        wy[2 * h + 1] = fma_t<T>(g.y, re1, phy[2 * h + 1]);
    }
}

// ---------------------------------------------------------------------------------------------
// The spectral control-period kernel, 16 lanes x 4 registers per env pair: 4 envs per warp.
// ---------------------------------------------------------------------------------------------
template <typename T, int RMODE>
__global__ void __launch_bounds__(kBlockThreads, KS_ETD16_MIN_BLOCKS) ks_etd16_kernel(const EtdParams ep)
{
    const Params &p = ep.p;
    constexpr int N = kEtdN;
    constexpr int LP = 16;                          // lanes per env pair
    constexpr int kWarps = kBlockThreads / 32;
    __shared__ C2<T> s_tab[kEtdTables * 2 * LP];
    __shared__ C2<T> s_tile[kWarps][2 * kTile16];

    // per-wavenumber tables -> shared memory: slot (table, h, c, q) = lane c, register k0 = 2h + q,
    // wavenumber index k = (c >> 2) + 4 (c & 3) + 16 k0
    {
        const T *src = static_cast<const T *>(ep.tables);
        for (int i = threadIdx.x; i < kEtdTables * 2 * LP * 2; i += kBlockThreads) {
            const int q = i & 1, c_ = (i >> 1) % LP, h = ((i >> 1) / LP) & 1, t = (i >> 1) / (2 * LP);
            const int k = (c_ >> 2) + 4 * (c_ & 3) + 16 * (2 * h + q);
            reinterpret_cast<T *>(&s_tab[(t * 2 + h) * LP + c_])[q] = src[t * N + k];
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * kBlockThreads + threadIdx.x) >> 5;
    const int grp = lane >> 4, a = lane & 15;       // pair slot in the warp, lane inside the pair
    const int envA = (warp * 2 + grp) * 2, envB = envA + 1;
    bool actA = envA < p.B, actB = envB < p.B;
    if (p.mask != nullptr) {
        actA = actA && p.mask[envA] != 0;
        actB = actB && p.mask[envB] != 0;
    }
    if (__ballot_sync(kFullMask, actA || actB) == 0u) return;   // warp-uniform (after the only __syncthreads)

    const C2<T> *tab = s_tab;
    Fft16Ctx<T> ctx;
    ctx.lo = a & 3;
    ctx.hi = a >> 2;
    ctx.tile = &s_tile[wib][grp * kTile16];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double sn, cs;
        sincospi(-2.0 * (double)(a * r) / 64.0, &sn, &cs);
        ctx.t1x[r] = (T)cs;
        ctx.t1y[r] = (T)sn;
        sincospi(-2.0 * (double)(ctx.lo * r) / 16.0, &sn, &cs);
        ctx.t2x[r] = (T)cs;
        ctx.t2y[r] = (T)sn;
    }

    // physical state of the pair: x = env A, y = env B (idle slots integrate zeros, never stored)
    T ux[4], uy[4];
    T *ua = static_cast<T *>(p.u) + (size_t)(actA ? envA : 0) * N + a;
    T *ub = static_cast<T *>(p.u) + (size_t)(actB ? envB : 0) * N + a;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        ux[r] = actA ? ua[LP * r] : T(0);
        uy[r] = actB ? ub[LP * r] : T(0);
    }
    int tsA = actA ? p.timestep[envA] : 0, tsB = actB ? p.timestep[envB] : 0;
    bool badA = actA ? p.nonfinite[envA] != 0 : false, badB = actB ? p.nonfinite[envB] != 0 : false;
    const T invN = T(1.0 / N);

    for (int k = 0; k < p.K; ++k) {
        // ---- jet forcing of this period: float32 FMA chain (transforms.py:262-265), then its spectrum
        T phx[4], phy[4];      // Q phi_hat / N, kept in registers for the whole period
        [[maybe_unused]] float pfx[4], pfy[4];     // the jets in physical space (dissipation reward: u * phi)
        {
            T (&fx)[4] = phx, (&fy)[4] = phy;
            if (p.phi != nullptr) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    fx[r] = actA ? T(p.phi[(size_t)envA * N + a + LP * r]) : T(0);
                    fy[r] = actB ? T(p.phi[(size_t)envB * N + a + LP * r]) : T(0);
                }
            } else if (p.actions != nullptr) {
                const float *aA = p.actions + ((size_t)k * p.B + (actA ? envA : 0)) * p.J;
                const float *aB = p.actions + ((size_t)k * p.B + (actB ? envB : 0)) * p.J;
                float fa[4], fb[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) fa[r] = fb[r] = 0.0f;
                for (int j = 0; j < p.J; ++j) {
                    const float va = __ldg(aA + j), vb = __ldg(aB + j);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float f = __ldg(p.F + (size_t)j * N + a + LP * r);
                        fa[r] = __fmaf_rn(va, f, fa[r]);
                        fb[r] = __fmaf_rn(vb, f, fb[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    fx[r] = actA ? T(fa[r]) : T(0);
                    fy[r] = actB ? T(fb[r]) : T(0);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r) fx[r] = fy[r] = T(0);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) { pfx[r] = (float)fx[r]; pfy[r] = (float)fy[r]; }
            fft64x16<T>(fx, fy, ctx);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const C2<T> qn = tab16<T>(tab, kTabQN, h, a);
                fx[2 * h] *= qn.x; fy[2 * h] *= qn.x;
                fx[2 * h + 1] *= qn.y; fy[2 * h + 1] *= qn.y;
            }
        }

        // ---- spectrum of the state, normalised: v = FFT(u) / N (idle slots restart from exact zeros
        // every period, see ks_etd.cuh)
        T vx[4], vy[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { vx[r] = actA ? ux[r] : T(0); vy[r] = actB ? uy[r] : T(0); }
        fft64x16<T>(vx, vy, ctx);
#pragma unroll
        for (int m = 0; m < 4; ++m) { vx[m] *= invN; vy[m] *= invN; }

        // ---- cfg_steps ETDRK4 steps (Cox & Matthews 2002, eqs. 26-29), with N~ = Q N:
        //   a = E2 v + N~v          b = (a - N~v) + N~a          c = (E2 a - N~v) + 2 N~b
        //   v' = E v + (f1/Q) N~v + (2 f2/Q) (N~a + N~b) + (f3/Q) N~c
        EtdReward<T> rw{{T(0), T(0)}, {T(0), T(0)}, {T(0), T(0)}};
        for (int s = 0; s < p.cfg_steps; ++s) {
            T nx[4], ny[4], ax[4], ay[4], wx[4], wy[4];
            if constexpr (RMODE == kRewardDissipation) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {      // uxx = IFFT(-k^2 v) of the pre-step state
                    const C2<T> k2 = tab16<T>(tab, kTabK2, h, a);
                    wx[2 * h] = k2.x * vx[2 * h];         wy[2 * h] = k2.x * vy[2 * h];
                    wx[2 * h + 1] = k2.y * vx[2 * h + 1]; wy[2 * h + 1] = k2.y * vy[2 * h + 1];
                }
                ifft64x16<T>(wx, wy, ctx);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    rw.a[0] = fma_t<T>(wx[r], wx[r], rw.a[0]);
                    rw.a[1] = fma_t<T>(wy[r], wy[r], rw.a[1]);
                }
            }
            // N~v
#pragma unroll
            for (int m = 0; m < 4; ++m) { nx[m] = vx[m]; ny[m] = vy[m]; }
            nonlinear16<T, true, RMODE>(nx, ny, ctx, tab, phx, phy, a, rw, pfx, pfy);
            // a = E2 v + N~v
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const C2<T> e2 = tab16<T>(tab, kTabE2, h, a);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int m = 2 * h + q;
                    const T c = q ? e2.y : e2.x;
                    wx[m] = ax[m] = fma_t<T>(c, vx[m], nx[m]);
                    wy[m] = ay[m] = fma_t<T>(c, vy[m], ny[m]);
                }
            }
            // N~a
            nonlinear16<T, false, RMODE>(wx, wy, ctx, tab, phx, phy, a, rw, pfx, pfy);
            // b = (a - N~v) + N~a (-> w);   a <- E2 a - N~v;   v <- E v + (f1/Q) N~v + (2 f2/Q) N~a
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const C2<T> e2 = tab16<T>(tab, kTabE2, h, a), e = tab16<T>(tab, kTabE, h, a);
                const C2<T> r1 = tab16<T>(tab, kTabR1, h, a), r22 = tab16<T>(tab, kTabR22, h, a);
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
            nonlinear16<T, false, RMODE>(wx, wy, ctx, tab, phx, phy, a, rw, pfx, pfy);
            // v += (2 f2/Q) N~b;   c = (E2 a - N~v) + 2 N~b (-> w)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const C2<T> r22 = tab16<T>(tab, kTabR22, h, a);
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
            nonlinear16<T, false, RMODE>(wx, wy, ctx, tab, phx, phy, a, rw, pfx, pfy);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const C2<T> r3 = tab16<T>(tab, kTabR3, h, a);
                vx[2 * h] = fma_t<T>(r3.x, wx[2 * h], vx[2 * h]);
                vy[2 * h] = fma_t<T>(r3.x, wy[2 * h], vy[2 * h]);
                vx[2 * h + 1] = fma_t<T>(r3.y, wx[2 * h + 1], vx[2 * h + 1]);
                vy[2 * h + 1] = fma_t<T>(r3.y, wy[2 * h + 1], vy[2 * h + 1]);
            }
        }

        // ---- back to physical space
#pragma unroll
        for (int m = 0; m < 4; ++m) { ux[m] = vx[m]; uy[m] = vy[m]; }
        ifft64x16<T>(ux, uy, ctx);

        // ---- period epilogue: reward, flags, observation (kuramoto.py:92-98)
        double ra = (double)rw.a[0] + (double)rw.b[0] + (double)rw.c[0];
        double rb = (double)rw.a[1] + (double)rw.b[1] + (double)rw.c[1];
        bool bA = false, bB = false;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            bA |= !(fabs((double)ux[r]) <= 1.7976931348623157e308);
            bB |= !(fabs((double)uy[r]) <= 1.7976931348623157e308);
        }
#pragma unroll
        for (int o = LP / 2; o > 0; o >>= 1) {  // fixed-order reduction over the lanes of the pair
            ra += __shfl_xor_sync(kFullMask, ra, o);
            rb += __shfl_xor_sync(kFullMask, rb, o);
        }
        const unsigned gmask = 0xffffu << (grp * LP);
        const bool anyA = (__ballot_sync(kFullMask, bA) & gmask) != 0u;
        const bool anyB = (__ballot_sync(kFullMask, bB) & gmask) != 0u;
        tsA += 1;
        tsB += 1;
        if (p.obs != nullptr) {
            if (p.obs_stride <= 1) {
                // for every register the 16 lanes of a pair store 16 consecutive floats (64-byte runs),
                // locally and -- gather mode -- into every peer's buffer
                float *oa = p.obs + ((size_t)k * p.B + envA) * N + a;
                float *ob = p.obs + ((size_t)k * p.B + envB) * N + a;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float va = (float)ux[r], vb = (float)uy[r];
                    if (actA) oa[LP * r] = va;
                    if (actB) ob[LP * r] = vb;
                    for (int q = 0; q < p.n_remote; ++q) {
                        if (actA) *remote_ptr(oa + LP * r, p.remote_delta[q]) = va;
                        if (actB) *remote_ptr(ob + LP * r, p.remote_delta[q]) = vb;
                    }
                }
            } else {
                const int first = p.obs_stride / 2;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int idx = a + LP * r - first;
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
        if (a == 0) {
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
    for (int r = 0; r < 4; ++r) {
        if (actA) ua[LP * r] = ux[r];
        if (actB) ub[LP * r] = uy[r];
    }
    if (a == 0) {
        if (actA) p.timestep[envA] = p.reset_timestep ? 0 : tsA;
        if (actB) p.timestep[envB] = p.reset_timestep ? 0 : tsB;
    }
}

}  // namespace ks
