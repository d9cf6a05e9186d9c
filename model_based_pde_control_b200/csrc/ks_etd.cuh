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
// Layout (N = 64):
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
//   * HBM is touched at control-period boundaries only (state in, state / observation / reward out).
#pragma once

#include "ks_kernels.cuh"

namespace ks {

constexpr int kEtdN = 64;          // grid points handled by this kernel
// Per-wavenumber tables, each [N] in natural FFT order (host-precomputed, ks_api.cu).  The kernel
// carries the nonlinear terms pre-multiplied by Q (N~ = Q N), which removes Q from the stage
// formulas; the final combination then needs f1/Q, 2 f2/Q, f3/Q (Q = h phi_1(hL/2)/2 > 0).
constexpr int kEtdTables = 7;
enum { kTabE = 0, kTabE2, kTabR1 /* f1/Q */, kTabR22 /* 2 f2/Q */, kTabR3 /* f3/Q */, kTabG /* Q g / N */,
       kTabQN /* Q / N */ };

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

// physical (lane n1, reg n2: n = n1 + 8 n2)  ->  spectral (lane k2, reg k1: k = 8 k1 + k2), unnormalised
template <typename T>
__device__ __forceinline__ void fft64(T (&x)[8], T (&y)[8], const T (&twx)[8], const T (&twy)[8], C2<T> *tile, int l)
{
    fft8<T>(x, y);
    twiddle8<T>(x, y, twx, twy);
    transpose8<T>(x, y, tile, l);
    fft8<T>(x, y);
}
// spectral -> physical, unnormalised (x/y exchanged = conjugated kernel)
template <typename T>
__device__ __forceinline__ void ifft64(T (&x)[8], T (&y)[8], const T (&twx)[8], const T (&twy)[8], C2<T> *tile, int l)
{
    fft8<T>(y, x);
    transpose8<T>(x, y, tile, l);
    twiddle8<T>(y, x, twx, twy);
    fft8<T>(y, x);
}

// Table values for this lane's registers m = 2h, 2h+1 (k = 8 m + j), one 2-element vector load.
// Layout in shared memory: [table][h][lane j].
template <typename T>
__device__ __forceinline__ C2<T> tab2(const C2<T> *tab, int which, int h, int l)
{
    return tab[(which * 4 + h) * 8 + l];
}

// Spectral right-hand side without the linear part, pre-multiplied by Q, in place:
//   w <- i (Q g / N) * FFT( (Re/Im IFFT(w))^2 ) + Q phi_hat / N        (both packed fields at once)
// With FIRST the squares of the physical values are the pre-step reward terms (kuramoto.py:84).
template <typename T, bool FIRST>
__device__ __forceinline__ void nonlinear(T (&wx)[8], T (&wy)[8], const T (&twx)[8], const T (&twy)[8], C2<T> *tile,
                                          const C2<T> *tab, const T (&phx)[8], const T (&phy)[8], int l, T &racc_a,
                                          T &racc_b)
{
    ifft64<T>(wx, wy, twx, twy, tile, l);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        wx[r] = wx[r] * wx[r];
        wy[r] = wy[r] * wy[r];
    }
    if constexpr (FIRST) {
        racc_a += ((wx[0] + wx[1]) + (wx[2] + wx[3])) + ((wx[4] + wx[5]) + (wx[6] + wx[7]));
        racc_b += ((wy[0] + wy[1]) + (wy[2] + wy[3])) + ((wy[4] + wy[5]) + (wy[6] + wy[7]));
    }
    fft64<T>(wx, wy, twx, twy, tile, l);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const C2<T> g = tab2<T>(tab, kTabG, h, l);
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
// The spectral control-period kernel: K periods x cfg_steps ETDRK4 steps, 8 envs per warp.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBlockThreads) ks_etd_kernel(const EtdParams ep)
{
    const Params &p = ep.p;
    constexpr int N = kEtdN;
    constexpr int kWarps = kBlockThreads / 32;
    __shared__ C2<T> s_tab[kEtdTables * 4 * 8];
    __shared__ C2<T> s_tile[kWarps][4 * 72];

    // per-wavenumber tables -> shared memory in register-pair layout (k = 8 m + j)
    {
        const T *src = static_cast<const T *>(ep.tables);
        for (int i = threadIdx.x; i < kEtdTables * N; i += kBlockThreads) {
            const int t = i / N, k = i % N, m = k >> 3, j = k & 7;
            T *dst = reinterpret_cast<T *>(&s_tab[(t * 4 + (m >> 1)) * 8 + j]);
            dst[m & 1] = src[i];
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * kBlockThreads + threadIdx.x) >> 5;
    const int grp = lane >> 3, l = lane & 7;
    const int envA = (warp * 4 + grp) * 2, envB = envA + 1;
    bool actA = envA < p.B, actB = envB < p.B;
    if (p.mask != nullptr) {
        actA = actA && p.mask[envA] != 0;
        actB = actB && p.mask[envB] != 0;
    }
    if (__ballot_sync(kFullMask, actA || actB) == 0u) return;   // warp-uniform (after the only __syncthreads)

    C2<T> *tile = &s_tile[wib][grp * 72];
    const C2<T> *tab = s_tab;

    // twiddles W64^(l r) = exp(-2 pi i l r / 64)
    T twx[8], twy[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double sn, cs;
        sincospi(-(double)(l * r) / 32.0, &sn, &cs);
        twx[r] = (T)cs;
        twy[r] = (T)sn;
    }

    // physical state of the pair: x = env A, y = env B (idle slots integrate zeros, never stored)
    T ux[8], uy[8];
    T *ua = static_cast<T *>(p.u) + (size_t)(actA ? envA : 0) * N + l;
    T *ub = static_cast<T *>(p.u) + (size_t)(actB ? envB : 0) * N + l;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        ux[r] = actA ? ua[8 * r] : T(0);
        uy[r] = actB ? ub[8 * r] : T(0);
    }
    int tsA = actA ? p.timestep[envA] : 0, tsB = actB ? p.timestep[envB] : 0;
    bool badA = actA ? p.nonfinite[envA] != 0 : false, badB = actB ? p.nonfinite[envB] != 0 : false;
    const T invN = T(1.0 / N);

    for (int k = 0; k < p.K; ++k) {
        // ---- jet forcing of this period: float32 FMA chain (transforms.py:262-265), then its spectrum
        T phx[8], phy[8];      // Q phi_hat / N, kept in registers for the whole period
        {
            T (&fx)[8] = phx, (&fy)[8] = phy;
            if (p.phi != nullptr) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    fx[r] = actA ? T(p.phi[(size_t)envA * N + l + 8 * r]) : T(0);
                    fy[r] = actB ? T(p.phi[(size_t)envB * N + l + 8 * r]) : T(0);
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
                        const float f = __ldg(p.F + (size_t)j * N + l + 8 * r);
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
            fft64<T>(fx, fy, twx, twy, tile, l);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> qn = tab2<T>(tab, kTabQN, h, l);
                fx[2 * h] *= qn.x; fy[2 * h] *= qn.x;
                fx[2 * h + 1] *= qn.y; fy[2 * h + 1] *= qn.y;
            }
        }

        // ---- spectrum of the state, normalised: v = FFT(u) / N
        T vx[8], vy[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { vx[r] = ux[r]; vy[r] = uy[r]; }
        fft64<T>(vx, vy, twx, twy, tile, l);
#pragma unroll
        for (int m = 0; m < 8; ++m) { vx[m] *= invN; vy[m] *= invN; }

        // ---- cfg_steps ETDRK4 steps (Cox & Matthews 2002, eqs. 26-29), with N~ = Q N:
        //   a = E2 v + N~v          b = E2 v + N~a = (a - N~v) + N~a          c = E2 a + 2 N~b - N~v
        //   v' = E v + (f1/Q) N~v + (2 f2/Q) (N~a + N~b) + (f3/Q) N~c
        T racc_a = T(0), racc_b = T(0), dummy = T(0);
        for (int s = 0; s < p.cfg_steps; ++s) {
            T nx[8], ny[8], ax[8], ay[8], wx[8], wy[8];
            // N~v
#pragma unroll
            for (int m = 0; m < 8; ++m) { nx[m] = vx[m]; ny[m] = vy[m]; }
            nonlinear<T, true>(nx, ny, twx, twy, tile, tab, phx, phy, l, racc_a, racc_b);
            // a = E2 v + N~v
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> e2 = tab2<T>(tab, kTabE2, h, l);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int m = 2 * h + q;
                    const T c = q ? e2.y : e2.x;
                    wx[m] = ax[m] = fma_t<T>(c, vx[m], nx[m]);
                    wy[m] = ay[m] = fma_t<T>(c, vy[m], ny[m]);
                }
            }
            // N~a
            nonlinear<T, false>(wx, wy, twx, twy, tile, tab, phx, phy, l, dummy, dummy);
            // b = (a - N~v) + N~a (-> w);   a <- E2 a - N~v;   v <- E v + (f1/Q) N~v + (2 f2/Q) N~a
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> e2 = tab2<T>(tab, kTabE2, h, l), e = tab2<T>(tab, kTabE, h, l);
                const C2<T> r1 = tab2<T>(tab, kTabR1, h, l), r22 = tab2<T>(tab, kTabR22, h, l);
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
            nonlinear<T, false>(wx, wy, twx, twy, tile, tab, phx, phy, l, dummy, dummy);
            // v += (2 f2/Q) N~b;   c = (E2 a - N~v) + 2 N~b (-> w)
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> r22 = tab2<T>(tab, kTabR22, h, l);
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
            nonlinear<T, false>(wx, wy, twx, twy, tile, tab, phx, phy, l, dummy, dummy);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const C2<T> r3 = tab2<T>(tab, kTabR3, h, l);
                vx[2 * h] = fma_t<T>(r3.x, wx[2 * h], vx[2 * h]);
                vy[2 * h] = fma_t<T>(r3.x, wy[2 * h], vy[2 * h]);
                vx[2 * h + 1] = fma_t<T>(r3.y, wx[2 * h + 1], vx[2 * h + 1]);
                vy[2 * h + 1] = fma_t<T>(r3.y, wy[2 * h + 1], vy[2 * h + 1]);
            }
        }

        // ---- back to physical space
#pragma unroll
        for (int m = 0; m < 8; ++m) { ux[m] = vx[m]; uy[m] = vy[m]; }
        ifft64<T>(ux, uy, twx, twy, tile, l);

        // ---- period epilogue: reward, flags, observation (kuramoto.py:92-98)
        double ra = (double)racc_a, rb = (double)racc_b;
        bool bA = false, bB = false;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            bA |= !(fabs((double)ux[r]) <= 1.7976931348623157e308);
            bB |= !(fabs((double)uy[r]) <= 1.7976931348623157e308);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {      // fixed-order reduction over the 8 lanes of the pair
            ra += __shfl_xor_sync(kFullMask, ra, o);
            rb += __shfl_xor_sync(kFullMask, rb, o);
        }
        const unsigned gmask = 0xffu << (grp * 8);
        const bool anyA = (__ballot_sync(kFullMask, bA) & gmask) != 0u;
        const bool anyB = (__ballot_sync(kFullMask, bB) & gmask) != 0u;
        tsA += 1;
        tsB += 1;
        if (p.obs != nullptr) {
            if (p.obs_stride <= 1) {
                float *oa = p.obs + ((size_t)k * p.B + envA) * N + l;
                float *ob = p.obs + ((size_t)k * p.B + envB) * N + l;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (actA) oa[8 * r] = (float)ux[r];
                    if (actB) ob[8 * r] = (float)uy[r];
                }
                if (p.n_remote > 0) {
                    // gather mode: stage the warp's 8 rows (512 floats, contiguous in the batch) in the
                    // transpose tile and mirror them to the peers as coalesced 16-byte-per-lane stores
                    float *stg = reinterpret_cast<float *>(&s_tile[wib][0]);
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        stg[(2 * grp) * N + l + 8 * r] = (float)ux[r];
                        stg[(2 * grp + 1) * N + l + 8 * r] = (float)uy[r];
                    }
                    __syncwarp();
                    float *wbase = p.obs + ((size_t)k * p.B + (size_t)warp * 8) * N;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i4 = j * 32 + lane, e = warp * 8 + (i4 * 4) / N;     // env that owns this float4
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
                    const int idx = l + 8 * r - first;
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
        if (l == 0) {
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
    if (p.n_remote > 0) __threadfence_system();   // peer stores performed before the launch retires

#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (actA) ua[8 * r] = ux[r];
        if (actB) ub[8 * r] = uy[r];
    }
    if (l == 0) {
        if (actA) p.timestep[envA] = p.reset_timestep ? 0 : tsA;
        if (actB) p.timestep[envB] = p.reset_timestep ? 0 : tsB;
    }
}

}  // namespace ks
