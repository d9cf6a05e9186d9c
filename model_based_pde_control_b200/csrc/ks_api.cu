// ks_api.cu -- C-ABI layer of libks_b200.so (see include/ks_b200.h for the contract and the
// reference lines each entry point replaces).  Host-side bookkeeping only: every numerical
// operation runs in the CUDA kernels of ks_kernels.cuh / this file; there is no CPU fallback.
#include "../../include/ks_b200.h"

#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ks_dispatch.h"
#include "ks_etd.cuh"

namespace {

// Finite-difference weights, stencil order (kuramoto.py:24-27 after un-flipping convolve1d).
constexpr double kUpwind[5] = {-25.0 / 12.0, 4.0, -3.0, 4.0 / 3.0, -1.0 / 4.0};
constexpr double kD2[4] = {-49.0 / 18.0, 3.0 / 2.0, -3.0 / 20.0, 1.0 / 90.0};            // centre, +-1..3
constexpr double kD4[5] = {91.0 / 8.0, -122.0 / 15.0, 169.0 / 60.0, -2.0 / 5.0, 7.0 / 240.0};  // centre, +-1..4

thread_local char g_create_error[256] = "";

inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

struct ks_handle {
    ks_config cfg;
    int P = 0, lanes = 0, envs_per_warp = 0, grid = 0, regs = 0, obs_len = 0;
    const void *kernel = nullptr;
    ks::Coef<double> c64;
    ks::Coef<float> c32;
    // device memory
    void *u = nullptr;          // [B,N] double or float
    int32_t *timestep = nullptr;
    uint8_t *nonfinite = nullptr;
    float *F = nullptr;         // [J,N]
    float *actions = nullptr;   // [B,J] staging for ks_step_host
    uint8_t *out = nullptr;     // packed output block
    double *scratch64 = nullptr;  // [B,N] f64 staging for set/get_state in F32 mode and host ICs
    uint8_t *mask = nullptr;      // [B] staging for host reset masks
    void *etd_tables = nullptr;   // [kEtdTables][N] T: ETDRK4 coefficient tables (spectral solver only)
    // fused all-gather over NVLink peer memory (ks_gather_*)
    int g_world = 0, g_rank = 0;
    uint8_t *g_buf = nullptr;       // local: [2 parities][world][out_total] + flags [world] u32 (own cudaMalloc)
    uint8_t *g_peer[KS_MAX_WORLD] = {nullptr};   // every rank's buffer as mapped into this process (own = g_buf)
    bool g_connected = false;
    bool g_external = false;        // buffers attached by the caller (ks_gather_attach): never freed / unmapped here
    uint8_t *g_mc = nullptr;        // NVLS multicast alias of the gather buffers (nullable)
    uint32_t g_epoch = 0;
    uint32_t g_barrier_epoch = 0;   // ks_gather_barrier's own epochs (second flag array of the flag block)
    int32_t *g_timeout = nullptr;   // device: [2] spare words + the table of every rank's flag array
    int32_t *g_timeout_host = nullptr;   // mapped pinned host word: 1 + rank of a peer that never signalled (sticky)
    long long g_timeout_cycles = 0;
    size_t g_slot = 0, g_flags_off = 0, g_total = 0;
    int32_t *collect_keys = nullptr;   // [2 parities][2]: ordered-int min / max keys of ks_collect
    uint32_t collect_calls = 0;
    size_t out_off[5] = {0, 0, 0, 0, 0}, out_total = 0;
    uint64_t launches = 0;
    char err[256] = "";
};

namespace {

int fail(ks_handle *h, int code, const char *fmt, ...)
{
    char *dst = h ? h->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(ks_handle *h, cudaError_t e, const char *what)
{
    cudaGetLastError();  // clear the sticky-less error state
    return fail(h, (int)e, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

#define KS_CUDA(h, call)                                      \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(h, e__, #call); \
    } while (0)

template <typename T>
void fill_coef(ks::Coef<T> &c, double dx, double dt)
{
    const double dx2 = dx * dx, dx4 = dx2 * dx2;
    for (int k = 0; k < 5; ++k) {
        const double d2 = k < 4 ? kD2[k] : 0.0;
        c.W[k] = (T)(-(kD4[k] / dx4 + d2 / dx2));
        c.A[k] = (T)(kUpwind[k] * (-0.5) / dx);
    }
    for (int k = 0; k < 4; ++k) c.D2[k] = (T)(kD2[k] / dx2);
    c.dt_half = (T)(dt / 2.0);
    c.dt_full = (T)dt;
    c.dt_sixth = (T)(dt / 6.0);
}

// ETDRK4 coefficient tables for the spectral solver, natural FFT order, float64 on the host.
//   L(k) = k^2 - k^4;  E = exp(hL), E2 = exp(hL/2);  Q, f1, f2, f3 = the phi-function combinations
//   of Cox & Matthews (2002) eqs. 26-29, evaluated as means over M = 32 points of the upper unit
//   half-circle around hL (Kassam & Trefethen 2005, section 3 -- avoids the cancellation of the
//   closed forms near L = 0);  g = -k/2 * dealias multiplies i*FFT(u^2).
// The kernel carries the nonlinear terms pre-multiplied by Q, so the tables it gets are
//   E, E2, f1/Q, 2 f2/Q, f3/Q, Q g / N, Q / N   (ks::kTabE ... kTabQN; the 1/N makes the kernel's
//   unnormalised transform pair an identity), and for the dissipation reward -k^2 and k / N.
void etd_tables_host(int N, double L, double h, bool dealias, std::vector<double> &out)
{
    constexpr int M = 32;
    const double pi = 3.14159265358979323846;
    out.assign((size_t)ks::kEtdTables * N, 0.0);
    for (int i = 0; i < N; ++i) {
        const int m = i < N / 2 ? i : i - N;                 // -N/2 .. N/2-1 (fftfreq order)
        const double k_even = 2.0 * pi / L * m;
        const double k_odd = (N % 2 == 0 && i == N / 2) ? 0.0 : k_even;   // odd derivative of the Nyquist mode = 0
        const double k2 = k_even * k_even, lin = k2 - k2 * k2;
        std::complex<double> sq(0, 0), s1(0, 0), s2(0, 0), s3(0, 0);
        for (int j = 1; j <= M; ++j) {
            const std::complex<double> r = std::polar(1.0, pi * (j - 0.5) / M);
            const std::complex<double> z = h * lin + r, ez = std::exp(z), z3 = z * z * z;
            sq += (std::exp(z / 2.0) - 1.0) / z;
            s1 += (-4.0 - z + ez * (4.0 - 3.0 * z + z * z)) / z3;
            s2 += (2.0 + z + ez * (-2.0 + z)) / z3;
            s3 += (-4.0 - 3.0 * z - z * z + ez * (4.0 - z)) / z3;
        }
        const double Q = h * sq.real() / M, f1 = h * s1.real() / M, f2 = h * s2.real() / M, f3 = h * s3.real() / M;
        const bool keep = !dealias || (m < 0 ? -m : m) <= N / 3;           // 2/3 rule
        out[(size_t)ks::kTabE * N + i] = std::exp(h * lin);
        out[(size_t)ks::kTabE2 * N + i] = std::exp(h * lin / 2.0);
        out[(size_t)ks::kTabR1 * N + i] = f1 / Q;
        out[(size_t)ks::kTabR22 * N + i] = 2.0 * f2 / Q;
        out[(size_t)ks::kTabR3 * N + i] = f3 / Q;
        out[(size_t)ks::kTabG * N + i] = keep ? Q * (-0.5 * k_odd) / N : 0.0;
        out[(size_t)ks::kTabQN * N + i] = Q / N;
        out[(size_t)ks::kTabK2 * N + i] = -k2;
        out[(size_t)ks::kTabKN * N + i] = k_odd / N;
    }
}

// Points per lane P for a grid of N points and a batch of B envs (lanes = N / P must fit one warp).
// Static cost model fitted to B200 measurements (profiles/README.md sections 10, 12 and R2-7), cycles per RK4
// sub-step of the fullest SM sub-partition (SMSP), which is what the launch takes:
//     T(P) = max( L(P), w * c(P) ),   w = ceil(warps(P) / #SMSP),   warps(P) = ceil(B / (32 / lanes))
//     c(P) = 182 P + 130   a warp's share of a saturated SMSP (859 / 1640 / 3048 cycles measured for
//                          P = 4 / 8 / 16 at 65 536 envs);  P < 4: 140 P + 280 (565-590 measured for P = 2)
//     L(P) = 169 P + 430   one warp alone: its dependent-issue latency (1105 / 1915 / 3135 cycles
//                          measured for P = 4 / 8 / 16);  P < 4: 140 P + 544 (824 measured for P = 2)
// Small batches therefore run with few points per lane (up to 592 envs at N = 64: 0.105 ms per period with
// P = 2, where the halo comes from two lanes on each side, against 0.138 ms with P = 4 and 0.40 ms with
// P = 16), large ones with many (less halo overhead per point).  Smallest T wins; a larger P wins ties
// within 0.5 %.  The choice is deterministic in (N, B, #SM) so that every rank of a sharded run picks
// the same layout.
int choose_points_per_lane(int N, long long B, int sm_count)
{
    int best = 0;
    double best_t = 0.0;
    const long long smsp = 4LL * (sm_count > 0 ? sm_count : 148);
    for (int P = ks::kMinP; P <= ks::kMaxP; ++P) {
        if (N % P) continue;
        const int lanes = N / P;
        if (lanes < 1 || lanes > 32) continue;
        const long long epw = 32 / lanes;
        const long long warps = (B + epw - 1) / epw;
        const long long per_smsp = (warps + smsp - 1) / smsp;
        const double share = P < 4 ? 140.0 * P + 280.0 : 182.0 * P + 130.0;
        double t = (double)per_smsp * share;
        const double alone = P < 4 ? 140.0 * P + 544.0 : 169.0 * P + 430.0;
        if (t < alone) t = alone;
        if (best == 0 || t <= best_t * 1.005) { best_t = t < best_t || best == 0 ? t : best_t; best = P; }
    }
    return best;
}

// Registers per lane of the spectral kernel at N = 64: 8 (ks_etd.cuh: 8 lanes x 8 complex registers per
// env pair, 8 envs per warp) or 4 (ks_etd16.cuh: 16 lanes x 4, 4 envs per warp).  Measured on B200
// (profiles/round2_sweep_etd_layouts.jsonl): a warp alone on its SM sub-partition needs 23.3 us per control
// period (10 ETDRK4 steps) in the 16-lane layout and 29.3 us in the 8-lane layout -- its dependent chain
// is half as long -- so the 16-lane layout is 20 % faster as long as every warp has a sub-partition to
// itself (B <= 4 x 592 = 2368 envs).  Beyond that it loses: it moves every value through the
// shared-memory crossbar twice per transform instead of once (600 against 688 wavefronts per warp-step for
// half the envs), and the crossbar (128 B/clk/SM) is what bounds this solver at large batches.
int choose_etd_regs_per_lane(long long B, int sm_count)
{
    const long long smsp = 4LL * (sm_count > 0 ? sm_count : 148);
    return (B + 3) / 4 <= smsp ? 4 : 8;
}

// ---------------------------------------------------------------------------------------------
// auxiliary kernels
// ---------------------------------------------------------------------------------------------
__global__ void convert_f64_to_f32(const double *__restrict__ src, float *__restrict__ dst, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}
__global__ void convert_f32_to_f64(const float *__restrict__ src, double *__restrict__ dst, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (double)src[i];
}

// Philox-4x32-10 (Salmon et al. 2011), counter = (point index, env), key = seed.
__device__ inline void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Initial condition of reset() for the envs selected by `mask` (all if nullptr):
//   u0 != nullptr : u[b,:] = u0[b,:]                     (host-drawn NumPy ICs, parity path)
//   u0 == nullptr : u[b,i] ~ U(-0.4, 0.4) (kuramoto.py:106), 53-bit uniform from Philox(seed; i, b)
// also clears the env's non-finite flag and, when there is no burn-in launch, its timestep.
template <typename T>
__global__ void reset_rows(T *__restrict__ u, const double *__restrict__ u0, const uint8_t *__restrict__ mask,
                           uint8_t *__restrict__ nonfinite, int32_t *__restrict__ timestep, int zero_timestep, int B,
                           int N, uint64_t seed, uint32_t env_base)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * N) return;
    const uint32_t env = (uint32_t)(idx / N), pt = (uint32_t)(idx % N);
    if (mask != nullptr && mask[env] == 0) return;
    if (u0 != nullptr) {
        u[idx] = (T)u0[idx];
    } else {
        // counter = (point, GLOBAL env index): a sharded run draws what the single-GPU run draws
        uint32_t c[4] = {pt, env + env_base, 0x4b53u /* 'KS' */, 0u};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            philox_round(c, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        const uint64_t bits = (((uint64_t)c[0] << 32) | c[1]) >> 11;
        const double r01 = (double)bits * (1.0 / 9007199254740992.0);
        u[idx] = (T)(-0.4 + 0.8 * r01);
    }
    if (pt == 0) {
        nonfinite[env] = 0;
        if (zero_timestep) timestep[env] = 0;
    }
}

// rhs() and reward on arbitrary states: one warp per row, unmerged stencils in the reference's
// own order of operations (kuramoto.py:118-129, 64-70).  Memory-bound helper, not the hot path.
__global__ void eval_rows(int M, int N, double dx, int reward_mode, const double *__restrict__ u,
                          const float *__restrict__ phi, double *__restrict__ rhs, double *__restrict__ ux_o,
                          double *__restrict__ uxx_o, double *__restrict__ uxxxx_o, double *__restrict__ reward)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const double *ur = u + (size_t)row * N;
    const float *pr = phi ? phi + (size_t)row * N : nullptr;
    const double dx2 = dx * dx, dx4 = dx2 * dx2;
    const double kUpwind[5] = {-25.0 / 12.0, 4.0, -3.0, 4.0 / 3.0, -1.0 / 4.0};
    const double kD2[4] = {-49.0 / 18.0, 3.0 / 2.0, -3.0 / 20.0, 1.0 / 90.0};
    const double kD4[5] = {91.0 / 8.0, -122.0 / 15.0, 169.0 / 60.0, -2.0 / 5.0, 7.0 / 240.0};
    double s_l2 = 0.0, s_uxx = 0.0, s_ux = 0.0, s_up = 0.0;
    for (int i = lane; i < N; i += 32) {
        double v[9];
#pragma unroll
        for (int k = -4; k <= 4; ++k) v[k + 4] = ur[(i + k + N) % N];
        double fwd = kUpwind[0] * (v[4] * v[4]), bwd = -kUpwind[0] * (v[4] * v[4]);
#pragma unroll
        for (int k = 1; k <= 4; ++k) {
            fwd = fwd + kUpwind[k] * (v[4 + k] * v[4 + k]);
            bwd = bwd - kUpwind[k] * (v[4 - k] * v[4 - k]);
        }
        const double ux = (v[4] < 0.0 ? fwd : bwd) / dx;
        double s2 = kD2[0] * v[4], s4 = kD4[0] * v[4];
#pragma unroll
        for (int k = 1; k <= 3; ++k) s2 += kD2[k] * (v[4 + k] + v[4 - k]);
#pragma unroll
        for (int k = 1; k <= 4; ++k) s4 += kD4[k] * (v[4 + k] + v[4 - k]);
        const double uxx = s2 / dx2, uxxxx = s4 / dx4;
        const double ph = pr ? (double)pr[i] : 0.0;
        const size_t o = (size_t)row * N + i;
        if (rhs) rhs[o] = -uxxxx - uxx - 0.5 * ux + ph;
        if (ux_o) ux_o[o] = ux;
        if (uxx_o) uxx_o[o] = uxx;
        if (uxxxx_o) uxxxx_o[o] = uxxxx;
        s_l2 += v[4] * v[4];
        s_uxx += uxx * uxx;
        s_ux += ux * ux;
        s_up += v[4] * ph;
    }
    if (reward) {
        double t = reward_mode == KS_REWARD_L2 ? s_l2 : (s_uxx + s_ux + s_up);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) reward[row] = -(t / N);
    }
}

// Completion handshake of the fused all-gather: one thread per peer.  Thread r publishes this
// rank's epoch into peer r's flag word (a peer store, after a system-scope fence: the period
// kernel that wrote the payload retired earlier on the same stream) and then waits until peer r's
// epoch shows up in the local flag word.  Only the GPUs of DIFFERENT ranks ever wait on each other
// here.  The spin is bounded by `max_cycles` (KS_GATHER_TIMEOUT_S, default 120 s -- ranks of a real
// training job drift apart by seconds: logging, checkpoints, evaluation passes); when the bound is
// hit the launch does NOT pretend the block is complete: it sets the sticky *timeout word, which
// lives in mapped host memory so that the next ks_step_gather / ks_gather_status sees it without a
// synchronise and fails, and it poisons the missing peer's slot (non-finite flags = 0xFF) so that a
// consumer that validates flags rejects the stale data even before the host has looked.
// `word0` selects the flag array inside the 256-byte flag block: 0 = the exchange's epochs, KS_MAX_WORLD = the
// stand-alone rendezvous of ks_gather_barrier (its own epoch counter; poison_len = 0 there).
__global__ void gather_signal_wait(uint32_t *const *peer_flags, volatile uint32_t *local_flags, int world, int rank,
                                   uint32_t epoch, volatile int32_t *timeout, long long max_cycles, uint8_t *poison_base,
                                   size_t slot_bytes, size_t poison_off, int poison_len, int word0)
{
    const int r = threadIdx.x;
    if (r >= world || r == rank) return;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(peer_flags[r] + word0 + rank) = epoch;
    const long long t0 = clock64();
    while ((int32_t)(local_flags[word0 + r] - epoch) < 0) {
        if (clock64() - t0 > max_cycles) {
            *timeout = 1 + r;
            uint8_t *bad = poison_base + (size_t)r * slot_bytes + poison_off;
            for (int i = 0; i < poison_len; ++i) bad[i] = 0xFF;
            break;
        }
        __nanosleep(64);
    }
    __threadfence_system();
}

// ---- ks_collect: fused wrapper plumbing --------------------------------------------------------------
__device__ __forceinline__ int float_key(float f)       // monotonic float -> int map (for atomicMin / Max)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void collect_minmax(const float *__restrict__ obs, size_t n, int32_t *__restrict__ keys)
{
    float lo = INFINITY, hi = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = obs[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&keys[0], float_key(lo));
        atomicMax(&keys[1], float_key(hi));
    }
}

__global__ void collect_apply(const ks_collect_args a, int B, int No, int J, const int32_t *__restrict__ keys,
                              int32_t *__restrict__ next_keys)
{
    float vmin = a.vminmax[0], vmax = a.vminmax[1];
    if (!a.frozen) {
        vmin = fminf(vmin, key_float(keys[0]));
        vmax = fmaxf(vmax, key_float(keys[1]));
    }
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    if (gid == 0) {
        // every thread merges the same two pairs, so whether it read the old or the merged running
        // bounds does not matter; the other parity's keys are re-armed for the next call
        a.vminmax[0] = vmin;
        a.vminmax[1] = vmax;
        next_keys[0] = 0x7fffffff;
        next_keys[1] = (int32_t)0x80000000;
    }
    const float width = __fsub_rn(vmax, vmin);
    const int first = a.agent_stride / 2;
    const int Na = (No - first + a.agent_stride - 1) / a.agent_stride;
    const size_t n = (size_t)B * No;
    const size_t t = a.slot_index ? (size_t)*a.slot_index : 0;      // incremented by collect_advance afterwards
    float *rec_obs = a.rec_obs + t * n, *rec_nxtobs = a.rec_nxtobs + t * n, *rec_actions = a.rec_actions + t * (size_t)B * J;
    double *rec_reward = a.rec_reward + t * B;
    uint8_t *rec_truncated = a.rec_truncated + t * B;
    int64_t *rec_step = a.rec_step + t * B;
    for (size_t i = gid; i < n; i += stride) {
        const float x = a.obs[i];
        rec_obs[i] = a.obs_store[i];
        rec_nxtobs[i] = x;
        a.obs_store[i] = x;
        const int b = (int)(i / No), c = (int)(i % No) - first;
        if (c >= 0 && c % a.agent_stride == 0) {
            const float s = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x, vmin), width), a.scale_width), a.lower);
            a.agent_obs[(size_t)b * Na + c / a.agent_stride] = s;
        }
    }
    for (size_t i = gid; i < (size_t)B * J; i += stride) {
        const float v = a.actions[i];
        rec_actions[i] = v;
        a.act_store[i] = v;
    }
    for (size_t i = gid; i < (size_t)B; i += stride) {
        rec_reward[i] = a.reward[i];
        rec_truncated[i] = a.truncated[i] != 0 ? 1 : 0;
        rec_step[i] = (int64_t)a.step[i];
    }
}

__global__ void collect_advance(int64_t *slot_index) { *slot_index += 1; }

// KS_GATHER_DEBUG (measurement only, profiles/README.md "exchange cost breakdown"): "nofence" leaves out the
// system-scope fence at the end of the period kernel, "nosignal" the signal / wait kernel.  Results are then
// NOT guaranteed to be complete when the stream reaches the consumer; bench.py refuses to verify such a run.
int gather_debug()
{
    static const int flags = []() {
        const char *e = getenv("KS_GATHER_DEBUG");
        int f = 0;
        if (e && strstr(e, "nofence")) f |= 1;
        if (e && strstr(e, "nosignal")) f |= 2;
        return f;
    }();
    return flags;
}

int launch_period(ks_handle *h, int K, const float *actions, const float *phi, float *obs, double *reward,
                  uint8_t *truncated, int32_t *step, uint8_t *nonfinite_out, int reset_timestep, const uint8_t *mask,
                  cudaStream_t stream, int n_remote = 0, const long long *remote_delta = nullptr, long long mc_delta = 0)
{
    ks::Params p;
    p.n_remote = n_remote;
    p.mc_delta = mc_delta;
    p.skip_fence = (n_remote > 0 && (gather_debug() & 1)) ? 1 : 0;
    for (int q = 0; q < ks::kMaxRemote; ++q) p.remote_delta[q] = q < n_remote ? remote_delta[q] : 0;
    p.u = h->u;
    p.timestep = h->timestep;
    p.nonfinite = h->nonfinite;
    p.F = h->F;
    p.actions = actions;
    p.phi = phi;
    p.obs = obs;
    p.reward = reward;
    p.truncated = truncated;
    p.step = step;
    p.nonfinite_out = nonfinite_out;
    p.mask = mask;
    p.B = h->cfg.num_envs;
    p.N = h->cfg.N;
    p.J = h->cfg.J;
    p.K = K;
    p.cfg_steps = h->cfg.cfg_steps;
    p.max_episode_steps = h->cfg.max_episode_steps;
    p.lanes = h->lanes;
    p.envs_per_warp = h->envs_per_warp;
    p.reset_timestep = reset_timestep;
    p.obs_stride = h->cfg.obs_stride < 1 ? 1 : h->cfg.obs_stride;
    p.obs_len = h->obs_len;
    p.inv_cfg_steps = 1.0 / h->cfg.cfg_steps;
    p.inv_N = 1.0 / h->cfg.N;
    if (h->cfg.solver == KS_SOLVER_ETDRK4) {
        ks::EtdParams ep;
        ep.p = p;
        ep.tables = h->etd_tables;
        void *eargs[1] = {&ep};
        KS_CUDA(h, cudaLaunchKernel(h->kernel, dim3(h->grid), dim3(ks::kBlockThreads), eargs, 0, stream));
        h->launches += 1;
        return KS_OK;
    }
    void *args[2] = {&p, h->cfg.precision == KS_F64 ? (void *)&h->c64 : (void *)&h->c32};
    KS_CUDA(h, cudaLaunchKernel(h->kernel, dim3(h->grid), dim3(ks::kBlockThreads), args, 0, stream));
    h->launches += 1;
    return KS_OK;
}

}  // namespace

// =============================================================================================
// exported C ABI
// =============================================================================================
extern "C" {

int ks_abi_version(void) { return KS_ABI_VERSION; }

const char *ks_last_error(const ks_handle *h) { return h ? h->err : g_create_error; }

int ks_create(const ks_config *cfg, ks_handle **out)
{
    if (!cfg || !out) return fail(nullptr, KS_ERR_ARG, "ks_create: NULL argument");
    *out = nullptr;
    if (cfg->abi_version != KS_ABI_VERSION)
        return fail(nullptr, KS_ERR_ARG, "ks_create: abi_version %d, library is %d", cfg->abi_version, KS_ABI_VERSION);
    if (cfg->num_envs < 1 || cfg->N < 9 || cfg->J < 1 || cfg->J > 32 || cfg->cfg_steps < 1 ||
        cfg->max_episode_steps < 1 || cfg->burnin_periods < 0 || !(cfg->L > 0.0) || !(cfg->dt > 0.0) || !cfg->forcing)
        return fail(nullptr, KS_ERR_ARG, "ks_create: config out of range (num_envs=%d N=%d J=%d cfg_steps=%d L=%g dt=%g)",
                    cfg->num_envs, cfg->N, cfg->J, cfg->cfg_steps, cfg->L, cfg->dt);
    if (cfg->precision != KS_F64 && cfg->precision != KS_F32) return fail(nullptr, KS_ERR_ARG, "ks_create: bad precision");
    if (cfg->obs_stride < 0 || cfg->obs_stride > cfg->N) return fail(nullptr, KS_ERR_ARG, "ks_create: bad obs_stride");
    if (cfg->env_index_base < 0) return fail(nullptr, KS_ERR_ARG, "ks_create: bad env_index_base");
    if (cfg->reward_mode != KS_REWARD_L2 && cfg->reward_mode != KS_REWARD_DISSIPATION)
        return fail(nullptr, KS_ERR_ARG, "ks_create: bad reward_mode");

    if (cfg->solver != KS_SOLVER_FD_RK4 && cfg->solver != KS_SOLVER_ETDRK4)
        return fail(nullptr, KS_ERR_ARG, "ks_create: bad solver %d", cfg->solver);
    const bool etd = cfg->solver == KS_SOLVER_ETDRK4;
    if (etd && cfg->N != 64 && cfg->N != 128 && cfg->N != 256)
        return fail(nullptr, KS_ERR_UNSUPPORTED, "ks_create: the spectral ETDRK4 solver supports N = 64, 128, 256 (N=%d)", cfg->N);

    // spectral solver: points_per_lane = complex registers per lane, 8 (all N) or 4 (N = 64 only), 0 = automatic
    if (etd && cfg->points_per_lane != 0 && cfg->points_per_lane != 8 && !(cfg->points_per_lane == 4 && cfg->N == 64))
        return fail(nullptr, KS_ERR_UNSUPPORTED,
                    "ks_create: the spectral solver takes points_per_lane = 0, 8, or (N = 64 only) 4 (got %d)", cfg->points_per_lane);
    int P = etd ? (cfg->N == 64 ? cfg->points_per_lane : 8) : cfg->points_per_lane;
    if (!etd && P != 0 && (P < ks::kMinP || P > ks::kMaxP || cfg->N % P || cfg->N / P > 32))
        return fail(nullptr, KS_ERR_UNSUPPORTED,
                    "ks_create: N=%d needs N = lanes*P with 2<=P<=16, lanes<=32 (points_per_lane=%d)", cfg->N,
                    cfg->points_per_lane);
    if (!etd && P == 0 && choose_points_per_lane(cfg->N, cfg->num_envs, 0) == 0)
        return fail(nullptr, KS_ERR_UNSUPPORTED, "ks_create: N=%d cannot be split as lanes*P with 2<=P<=16, lanes<=32",
                    cfg->N);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(nullptr, KS_ERR_NO_DEVICE, "ks_create: no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, KS_ERR_ARG, "ks_create: device %d of %d", cfg->device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, KS_ERR_NO_DEVICE, "ks_create: device %d is sm_%d%d; kernels are built for sm_100a only",
                    cfg->device, prop.major, prop.minor);
    if (P == 0)
        P = etd ? choose_etd_regs_per_lane(cfg->num_envs, prop.multiProcessorCount)
                : choose_points_per_lane(cfg->N, cfg->num_envs, prop.multiProcessorCount);

    ks_handle *h = new (std::nothrow) ks_handle();
    if (!h) return fail(nullptr, KS_ERR_ARG, "ks_create: out of host memory");
    h->cfg = *cfg;
    h->cfg.forcing = nullptr;
    h->cfg.points_per_lane = P;
    h->P = P;
    const int stride = cfg->obs_stride < 1 ? 1 : cfg->obs_stride;
    h->cfg.obs_stride = stride;
    h->obs_len = (cfg->N - stride / 2 + stride - 1) / stride;
    h->lanes = cfg->N / P;
    h->envs_per_warp = 32 / h->lanes;
    const long long warps = ((long long)cfg->num_envs + h->envs_per_warp - 1) / h->envs_per_warp;
    h->grid = (int)((warps * 32 + ks::kBlockThreads - 1) / ks::kBlockThreads);
    const bool f64 = cfg->precision == KS_F64, l2 = cfg->reward_mode == KS_REWARD_L2;
    h->kernel = f64 ? (l2 ? ks::period_kernel_f64_l2(P) : ks::period_kernel_f64_diss(P))
                    : (l2 ? ks::period_kernel_f32_l2(P) : ks::period_kernel_f32_diss(P));
    if (etd) {
        // a pair of envs per 8R lanes x 8 complex registers (N = 64 R), 8/R envs per warp (ks_etd.cuh)
        const int R = cfg->N / ks::kEtdN;
        h->lanes = 8 * R;
        h->envs_per_warp = 8 / R;
        const int per_cta = h->envs_per_warp * (ks::kBlockThreads / 32);
        h->grid = (int)((cfg->num_envs + per_cta - 1) / per_cta);
        h->kernel = f64 ? ks::etd_kernel_f64(R, cfg->reward_mode) : ks::etd_kernel_f32(R, cfg->reward_mode);
        if (P == 4) {      // small-batch layout: 16 lanes x 4 registers per pair, 4 envs per warp (ks_etd16.cuh)
            h->lanes = 16;
            h->envs_per_warp = 4;
            const int per_cta16 = 4 * (ks::kBlockThreads / 32);
            h->grid = (int)((cfg->num_envs + per_cta16 - 1) / per_cta16);
            h->kernel = f64 ? ks::etd16_kernel_f64(cfg->reward_mode) : ks::etd16_kernel_f32(cfg->reward_mode);
        }
    }
    const double dx = cfg->L / cfg->N;  // kuramoto.py:55
    fill_coef(h->c64, dx, cfg->dt);
    fill_coef(h->c32, dx, cfg->dt);

    DeviceGuard guard(cfg->device);
    const size_t B = cfg->num_envs, N = cfg->N, J = cfg->J;
    const size_t esz = f64 ? sizeof(double) : sizeof(float);
    h->out_off[0] = 0;
    h->out_off[1] = align16(B * sizeof(double));
    h->out_off[2] = h->out_off[1] + align16(B * (size_t)h->obs_len * sizeof(float));
    h->out_off[3] = h->out_off[2] + align16(B * sizeof(int32_t));
    h->out_off[4] = h->out_off[3] + align16(B);
    h->out_total = h->out_off[4] + align16(B);
    cudaFuncAttributes attr;
    int rc = KS_OK;
    do {
        if (!guard.ok) { rc = fail(nullptr, KS_ERR_NO_DEVICE, "ks_create: cudaSetDevice(%d) failed", cfg->device); break; }
#define KS_TRY(call)                                                                  \
    if ((e = (call)) != cudaSuccess) { rc = cuda_fail(nullptr, e, #call); break; }
        KS_TRY(cudaFuncGetAttributes(&attr, h->kernel));
        h->regs = attr.numRegs;
        KS_TRY(cudaMalloc(&h->u, B * N * esz));
        KS_TRY(cudaMalloc(&h->timestep, B * sizeof(int32_t)));
        KS_TRY(cudaMalloc(&h->nonfinite, B));
        KS_TRY(cudaMalloc(&h->F, J * N * sizeof(float)));
        KS_TRY(cudaMalloc(&h->actions, B * J * sizeof(float)));
        KS_TRY(cudaMalloc(&h->out, h->out_total));
        KS_TRY(cudaMalloc(&h->scratch64, B * N * sizeof(double)));
        KS_TRY(cudaMalloc(&h->mask, B));
        KS_TRY(cudaMalloc(&h->collect_keys, 4 * sizeof(int32_t)));
        {
            const int32_t init[4] = {0x7fffffff, (int32_t)0x80000000, 0x7fffffff, (int32_t)0x80000000};
            KS_TRY(cudaMemcpy(h->collect_keys, init, sizeof(init), cudaMemcpyHostToDevice));
        }
        KS_TRY(cudaMemset(h->u, 0, B * N * esz));
        KS_TRY(cudaMemset(h->timestep, 0, B * sizeof(int32_t)));
        KS_TRY(cudaMemset(h->nonfinite, 0, B));
        KS_TRY(cudaMemset(h->out, 0, h->out_total));
        KS_TRY(cudaMemcpy(h->F, cfg->forcing, J * N * sizeof(float), cudaMemcpyHostToDevice));
        if (etd) {
            std::vector<double> tab;
            etd_tables_host(cfg->N, cfg->L, cfg->dt, cfg->dealias != 0, tab);
            KS_TRY(cudaMalloc(&h->etd_tables, tab.size() * esz));
            if (f64) {
                KS_TRY(cudaMemcpy(h->etd_tables, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
            } else {
                std::vector<float> tab32(tab.begin(), tab.end());
                KS_TRY(cudaMemcpy(h->etd_tables, tab32.data(), tab32.size() * sizeof(float), cudaMemcpyHostToDevice));
            }
        }
#undef KS_TRY
    } while (0);
    if (rc != KS_OK) {
        ks_destroy(h);
        return rc;
    }
    *out = h;
    return KS_OK;
}

int ks_destroy(ks_handle *h)
{
    if (!h) return KS_OK;
    {
        DeviceGuard guard(h->cfg.device);
        cudaFree(h->u);
        cudaFree(h->timestep);
        cudaFree(h->nonfinite);
        cudaFree(h->F);
        cudaFree(h->actions);
        cudaFree(h->out);
        cudaFree(h->scratch64);
        cudaFree(h->mask);
        cudaFree(h->collect_keys);
        cudaFree(h->etd_tables);
        if (h->g_connected && !h->g_external)
            for (int r = 0; r < h->g_world; ++r)
                if (r != h->g_rank && h->g_peer[r]) cudaIpcCloseMemHandle(h->g_peer[r]);
        if (!h->g_external) cudaFree(h->g_buf);
        cudaFree(h->g_timeout);
        if (h->g_timeout_host) cudaFreeHost(h->g_timeout_host);
        cudaGetLastError();
    }
    delete h;
    return KS_OK;
}

int ks_get_config(const ks_handle *h, ks_config *out)
{
    if (!h || !out) return KS_ERR_ARG;
    *out = h->cfg;
    return KS_OK;
}

int ks_launch_info(const ks_handle *h, int32_t *P, int32_t *lanes, int32_t *block, int32_t *grid, int32_t *regs)
{
    if (!h) return KS_ERR_ARG;
    if (P) *P = h->P;
    if (lanes) *lanes = h->lanes;
    if (block) *block = ks::kBlockThreads;
    if (grid) *grid = h->grid;
    if (regs) *regs = h->regs;
    return KS_OK;
}

uint64_t ks_launch_count(const ks_handle *h) { return h ? h->launches : 0; }

int ks_out_layout(const ks_handle *h, size_t offsets[5], size_t *total)
{
    if (!h) return KS_ERR_ARG;
    if (offsets) memcpy(offsets, h->out_off, sizeof(h->out_off));
    if (total) *total = h->out_total;
    return KS_OK;
}

int ks_set_state(ks_handle *h, const double *u, const int32_t *timestep, int where, void *stream_)
{
    if (!h) return KS_ERR_ARG;
    if (where != KS_HOST && where != KS_DEVICE) return fail(h, KS_ERR_ARG, "ks_set_state: bad `where`");
    cudaStream_t stream = (cudaStream_t)stream_;
    DeviceGuard guard(h->cfg.device);
    const size_t B = h->cfg.num_envs, n = B * h->cfg.N;
    const cudaMemcpyKind kind = where == KS_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (u) {
        if (h->cfg.precision == KS_F64) {
            KS_CUDA(h, cudaMemcpyAsync(h->u, u, n * sizeof(double), kind, stream));
        } else {
            const double *src = u;
            if (where == KS_HOST) {
                KS_CUDA(h, cudaMemcpyAsync(h->scratch64, u, n * sizeof(double), kind, stream));
                src = h->scratch64;
            }
            convert_f64_to_f32<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, (float *)h->u, n);
            KS_CUDA(h, cudaGetLastError());
            h->launches += 1;
        }
    }
    if (timestep) KS_CUDA(h, cudaMemcpyAsync(h->timestep, timestep, B * sizeof(int32_t), kind, stream));
    KS_CUDA(h, cudaMemsetAsync(h->nonfinite, 0, B, stream));
    if (where == KS_HOST) KS_CUDA(h, cudaStreamSynchronize(stream));  // caller may reuse its host buffers
    return KS_OK;
}

int ks_get_state(ks_handle *h, double *u, int32_t *timestep, int where, void *stream_)
{
    if (!h) return KS_ERR_ARG;
    if (where != KS_HOST && where != KS_DEVICE) return fail(h, KS_ERR_ARG, "ks_get_state: bad `where`");
    cudaStream_t stream = (cudaStream_t)stream_;
    DeviceGuard guard(h->cfg.device);
    const size_t B = h->cfg.num_envs, n = B * h->cfg.N;
    const cudaMemcpyKind kind = where == KS_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (u) {
        if (h->cfg.precision == KS_F64) {
            KS_CUDA(h, cudaMemcpyAsync(u, h->u, n * sizeof(double), kind, stream));
        } else {
            double *dst = where == KS_HOST ? h->scratch64 : u;
            convert_f32_to_f64<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const float *)h->u, dst, n);
            KS_CUDA(h, cudaGetLastError());
            h->launches += 1;
            if (where == KS_HOST) KS_CUDA(h, cudaMemcpyAsync(u, h->scratch64, n * sizeof(double), kind, stream));
        }
    }
    if (timestep) KS_CUDA(h, cudaMemcpyAsync(timestep, h->timestep, B * sizeof(int32_t), kind, stream));
    if (where == KS_HOST) KS_CUDA(h, cudaStreamSynchronize(stream));
    return KS_OK;
}

int ks_reset(ks_handle *h, const double *u0, const uint8_t *mask, int where, uint64_t seed,
             int32_t burnin_periods, void *stream_)
{
    if (!h) return KS_ERR_ARG;
    if (where != KS_HOST && where != KS_DEVICE) return fail(h, KS_ERR_ARG, "ks_reset: bad `where`");
    cudaStream_t stream = (cudaStream_t)stream_;
    DeviceGuard guard(h->cfg.device);
    const size_t B = h->cfg.num_envs, n = B * h->cfg.N;
    const double *u0_dev = u0;
    const uint8_t *mask_dev = mask;
    if (where == KS_HOST) {
        if (u0) {
            KS_CUDA(h, cudaMemcpyAsync(h->scratch64, u0, n * sizeof(double), cudaMemcpyHostToDevice, stream));
            u0_dev = h->scratch64;
        }
        if (mask) {
            KS_CUDA(h, cudaMemcpyAsync(h->mask, mask, B, cudaMemcpyHostToDevice, stream));
            mask_dev = h->mask;
        }
    }
    const int K = burnin_periods < 0 ? h->cfg.burnin_periods : burnin_periods;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (h->cfg.precision == KS_F64)
        reset_rows<double><<<blocks, 256, 0, stream>>>((double *)h->u, u0_dev, mask_dev, h->nonfinite, h->timestep,
                                                       K == 0, (int)B, h->cfg.N, seed, (uint32_t)h->cfg.env_index_base);
    else
        reset_rows<float><<<blocks, 256, 0, stream>>>((float *)h->u, u0_dev, mask_dev, h->nonfinite, h->timestep,
                                                      K == 0, (int)B, h->cfg.N, seed, (uint32_t)h->cfg.env_index_base);
    KS_CUDA(h, cudaGetLastError());
    h->launches += 1;
    int rc = KS_OK;
    if (K > 0) rc = launch_period(h, K, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1, mask_dev, stream);
    // host buffers (u0 / mask) may be reused by the caller as soon as we return
    if (rc == KS_OK && where == KS_HOST && (u0 || mask)) KS_CUDA(h, cudaStreamSynchronize(stream));
    return rc;
}

int ks_step(ks_handle *h, const float *actions, const float *phi, float *obs, double *reward, uint8_t *truncated,
            int32_t *step, uint8_t *nonfinite, void *stream)
{
    if (!h) return KS_ERR_ARG;
    if (!actions && !phi) return fail(h, KS_ERR_ARG, "ks_step: need actions or phi");
    DeviceGuard guard(h->cfg.device);
    return launch_period(h, 1, actions, phi, obs, reward, truncated, step, nonfinite, 0, nullptr, (cudaStream_t)stream);
}

int ks_rollout(ks_handle *h, int32_t K, const float *actions, float *obs, double *reward, uint8_t *truncated,
               int32_t *step, uint8_t *nonfinite, void *stream)
{
    if (!h) return KS_ERR_ARG;
    if (K < 1) return fail(h, KS_ERR_ARG, "ks_rollout: K = %d", K);
    DeviceGuard guard(h->cfg.device);
    return launch_period(h, K, actions, nullptr, obs, reward, truncated, step, nonfinite, 0, nullptr, (cudaStream_t)stream);
}

namespace {
// Is `p` pinned host memory that kernels of the current device can address as it is (UVA)?
bool device_addressable_host(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost && a.devicePointer == p;
}
}  // namespace

int ks_step_host(ks_handle *h, const float *actions_host, void *out_host, void *stream_)
{
    if (!h || !actions_host || !out_host) return h ? fail(h, KS_ERR_ARG, "ks_step_host: NULL argument") : KS_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    DeviceGuard guard(h->cfg.device);
    const size_t B = h->cfg.num_envs;
    float *obs = (float *)(h->out + h->out_off[1]);
    double *reward = (double *)(h->out + h->out_off[0]);
    uint8_t *trunc = h->out + h->out_off[3], *bad = h->out + h->out_off[4];
    int32_t *step = (int32_t *)(h->out + h->out_off[2]);
    // Zero-copy form (both buffers pinned, the normal case: KSVecEnv allocates them so): the kernel
    // reads the actions straight from host memory and its epilogue mirrors every output into the host
    // block -- the same mirrored, shared-memory-staged stores the multi-GPU exchange uses, with the
    // host block as the only "peer" -- so the step is ONE launch and a synchronise, no copy calls.
    static const bool allow_zero_copy = []() { const char *m = getenv("KS_HOST_IO"); return !(m && strcmp(m, "copy") == 0); }();
    // (sensor-strided observations are written element by element: those go through the copy form)
    if (allow_zero_copy && h->cfg.obs_stride <= 1 && device_addressable_host(actions_host) && device_addressable_host(out_host)) {
        const long long delta = (long long)((uint8_t *)out_host - h->out);
        int rc = launch_period(h, 1, actions_host, nullptr, obs, reward, trunc, step, bad, 0, nullptr, stream, 1, &delta);
        if (rc != KS_OK) return rc;
        KS_CUDA(h, cudaStreamSynchronize(stream));
        return KS_OK;
    }
    KS_CUDA(h, cudaMemcpyAsync(h->actions, actions_host, B * h->cfg.J * sizeof(float), cudaMemcpyHostToDevice, stream));
    int rc = launch_period(h, 1, h->actions, nullptr, obs, reward, trunc, step, bad, 0, nullptr, stream);
    if (rc != KS_OK) return rc;
    KS_CUDA(h, cudaMemcpyAsync(out_host, h->out, h->out_total, cudaMemcpyDeviceToHost, stream));
    KS_CUDA(h, cudaStreamSynchronize(stream));
    return KS_OK;
}

int ks_status(ks_handle *h, uint8_t *nonfinite_host, int32_t *any, void *stream_)
{
    if (!h) return KS_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    DeviceGuard guard(h->cfg.device);
    const size_t B = h->cfg.num_envs;
    uint8_t *tmp = nonfinite_host;
    if (!tmp) {
        tmp = new (std::nothrow) uint8_t[B];
        if (!tmp) return fail(h, KS_ERR_ARG, "ks_status: out of host memory");
    }
    cudaError_t e = cudaMemcpyAsync(tmp, h->nonfinite, B, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    int32_t acc = 0;
    if (e == cudaSuccess)
        for (size_t i = 0; i < B; ++i) acc |= tmp[i];
    if (!nonfinite_host) delete[] tmp;
    if (e != cudaSuccess) return cuda_fail(h, e, "ks_status copy");
    if (any) *any = acc ? 1 : 0;
    return KS_OK;
}

int ks_eval(ks_handle *h, int32_t M, const double *u, const float *phi, double *rhs, double *ux, double *uxx,
            double *uxxxx, double *reward, void *stream)
{
    if (!h || !u || M < 0) return h ? fail(h, KS_ERR_ARG, "ks_eval: bad argument") : KS_ERR_ARG;
    if (M == 0) return KS_OK;
    DeviceGuard guard(h->cfg.device);
    const double dx = h->cfg.L / h->cfg.N;
    const unsigned blocks = (unsigned)(((size_t)M * 32 + 127) / 128);
    eval_rows<<<blocks, 128, 0, (cudaStream_t)stream>>>(M, h->cfg.N, dx, h->cfg.reward_mode, u, phi, rhs, ux, uxx,
                                                        uxxxx, reward);
    KS_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return KS_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused all-gather of the packed output block over NVLink peer memory
// ---------------------------------------------------------------------------------------------
int ks_gather_init(ks_handle *h, int32_t world, int32_t rank, void *ipc_handle_out, size_t *slot_bytes)
{
    if (!h || !ipc_handle_out) return h ? fail(h, KS_ERR_ARG, "ks_gather_init: NULL argument") : KS_ERR_ARG;
    if (world < 1 || world > KS_MAX_WORLD || rank < 0 || rank >= world)
        return fail(h, KS_ERR_ARG, "ks_gather_init: world=%d rank=%d (max world %d)", world, rank, KS_MAX_WORLD);
    if (h->g_buf) return fail(h, KS_ERR_STATE, "ks_gather_init: already initialised");
    static_assert(sizeof(cudaIpcMemHandle_t) == KS_IPC_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(h->cfg.device);
    h->g_world = world;
    h->g_rank = rank;
    h->g_slot = (h->out_total + 255) & ~size_t(255);
    h->g_flags_off = 2 * (size_t)world * h->g_slot;
    h->g_total = h->g_flags_off + 256;
    KS_CUDA(h, cudaMalloc(&h->g_buf, h->g_total));
    KS_CUDA(h, cudaMemset(h->g_buf, 0, h->g_total));
    KS_CUDA(h, cudaMalloc(&h->g_timeout, 2 * sizeof(int32_t) + KS_MAX_WORLD * sizeof(void *)));
    KS_CUDA(h, cudaMemset(h->g_timeout, 0, 2 * sizeof(int32_t) + KS_MAX_WORLD * sizeof(void *)));
    KS_CUDA(h, cudaHostAlloc((void **)&h->g_timeout_host, sizeof(int32_t), cudaHostAllocMapped));
    *h->g_timeout_host = 0;
    {
        double seconds = 120.0;
        if (const char *e = getenv("KS_GATHER_TIMEOUT_S")) {
            const double v = atof(e);
            if (v > 0.0) seconds = v;
        }
        int khz = 1965000;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->cfg.device);
        h->g_timeout_cycles = (long long)(seconds * 1e3 * (double)khz);
    }
    KS_CUDA(h, cudaDeviceSynchronize());
    cudaIpcMemHandle_t hd;
    KS_CUDA(h, cudaIpcGetMemHandle(&hd, h->g_buf));
    memcpy(ipc_handle_out, &hd, sizeof(hd));
    if (slot_bytes) *slot_bytes = h->g_slot;
    h->g_peer[rank] = h->g_buf;
    return KS_OK;
}

int ks_gather_connect(ks_handle *h, const void *all_handles)
{
    if (!h || !all_handles) return h ? fail(h, KS_ERR_ARG, "ks_gather_connect: NULL argument") : KS_ERR_ARG;
    if (!h->g_buf) return fail(h, KS_ERR_STATE, "ks_gather_connect: call ks_gather_init first");
    if (h->g_connected) return fail(h, KS_ERR_STATE, "ks_gather_connect: already connected");
    DeviceGuard guard(h->cfg.device);
    const uint8_t *hs = static_cast<const uint8_t *>(all_handles);
    for (int r = 0; r < h->g_world; ++r) {
        if (r == h->g_rank) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, hs + (size_t)r * KS_IPC_HANDLE_BYTES, sizeof(hd));
        void *ptr = nullptr;
        KS_CUDA(h, cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
        h->g_peer[r] = static_cast<uint8_t *>(ptr);
    }
    // table of every rank's flag array for the signal kernel (lives behind the timeout word)
    uint32_t *flags[KS_MAX_WORLD] = {nullptr};
    for (int r = 0; r < h->g_world; ++r) flags[r] = reinterpret_cast<uint32_t *>(h->g_peer[r] + h->g_flags_off);
    KS_CUDA(h, cudaMemcpy(h->g_timeout + 2, flags, sizeof(flags), cudaMemcpyHostToDevice));
    *h->g_timeout_host = 0;      // (re-)arm
    h->g_connected = true;
    return KS_OK;
}

int ks_step_gather(ks_handle *h, const float *actions, void **gathered, void *stream_)
{
    if (!h || !actions) return h ? fail(h, KS_ERR_ARG, "ks_step_gather: NULL argument") : KS_ERR_ARG;
    if (!h->g_buf || (h->g_world > 1 && !h->g_connected))
        return fail(h, KS_ERR_STATE, "ks_step_gather: gather not initialised / connected");
    if (h->g_timeout_host && *(volatile int32_t *)h->g_timeout_host != 0)
        return fail(h, KS_ERR_STATE,
                    "ks_step_gather: rank %d never signalled an earlier exchange within the handshake bound "
                    "(KS_GATHER_TIMEOUT_S); the gathered blocks since then are incomplete -- ks_gather_clear re-arms",
                    *(volatile int32_t *)h->g_timeout_host - 1);
    cudaStream_t stream = (cudaStream_t)stream_;
    DeviceGuard guard(h->cfg.device);
    h->g_epoch += 1;
    const size_t parity_off = (size_t)(h->g_epoch & 1u) * h->g_world * h->g_slot;
    uint8_t *mine = h->g_buf + parity_off + (size_t)h->g_rank * h->g_slot;
    long long delta[ks::kMaxRemote];
    int n = 0;
    for (int r = 0; r < h->g_world; ++r) {
        if (r == h->g_rank) continue;
        uint8_t *theirs = h->g_peer[r] + parity_off + (size_t)h->g_rank * h->g_slot;
        delta[n++] = (long long)(theirs - mine);
    }
    // multicast alias of this rank's slot (FD-RK4 kernel only; the spectral kernels use the unicast stores)
    const long long mc_delta = (h->g_mc && h->cfg.solver == KS_SOLVER_FD_RK4 && h->g_world > 1)
                                   ? (long long)((h->g_mc + parity_off + (size_t)h->g_rank * h->g_slot) - mine) : 0;
    int rc = launch_period(h, 1, actions, nullptr, (float *)(mine + h->out_off[1]), (double *)(mine + h->out_off[0]),
                           mine + h->out_off[3], (int32_t *)(mine + h->out_off[2]), mine + h->out_off[4], 0, nullptr, stream,
                           n, delta, mc_delta);
    if (rc != KS_OK) return rc;
    if (h->g_world > 1 && !(gather_debug() & 2)) {
        gather_signal_wait<<<1, 32, 0, stream>>>(reinterpret_cast<uint32_t *const *>(h->g_timeout + 2),
                                                 reinterpret_cast<volatile uint32_t *>(h->g_buf + h->g_flags_off), h->g_world,
                                                 h->g_rank, h->g_epoch, h->g_timeout_host, h->g_timeout_cycles,
                                                 h->g_buf + parity_off, h->g_slot, h->out_off[4], h->cfg.num_envs, 0);
        KS_CUDA(h, cudaGetLastError());
        h->launches += 1;
    }
    if (gathered) *gathered = h->g_buf + parity_off;
    return KS_OK;
}

int ks_gather_barrier(ks_handle *h, void *stream_)
{
    if (!h) return KS_ERR_ARG;
    if (!h->g_buf || (h->g_world > 1 && !h->g_connected))
        return fail(h, KS_ERR_STATE, "ks_gather_barrier: gather not initialised / connected");
    if (h->g_world < 2) return KS_OK;
    if (h->g_timeout_host && *(volatile int32_t *)h->g_timeout_host != 0)
        return fail(h, KS_ERR_STATE, "ks_gather_barrier: rank %d never signalled earlier (ks_gather_clear re-arms)",
                    *(volatile int32_t *)h->g_timeout_host - 1);
    DeviceGuard guard(h->cfg.device);
    h->g_barrier_epoch += 1;
    gather_signal_wait<<<1, 32, 0, (cudaStream_t)stream_>>>(reinterpret_cast<uint32_t *const *>(h->g_timeout + 2),
                                                            reinterpret_cast<volatile uint32_t *>(h->g_buf + h->g_flags_off),
                                                            h->g_world, h->g_rank, h->g_barrier_epoch, h->g_timeout_host,
                                                            h->g_timeout_cycles, h->g_buf, h->g_slot, h->out_off[4], 0,
                                                            KS_MAX_WORLD);
    KS_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return KS_OK;
}

int ks_gather_layout(const ks_handle *h, int32_t world, size_t *slot_bytes, size_t *total_bytes)
{
    if (!h || world < 1 || world > KS_MAX_WORLD) return KS_ERR_ARG;
    const size_t slot = (h->out_total + 255) & ~size_t(255);
    if (slot_bytes) *slot_bytes = slot;
    if (total_bytes) *total_bytes = 2 * (size_t)world * slot + 256;
    return KS_OK;
}

int ks_gather_attach(ks_handle *h, int32_t world, int32_t rank, void *const *peer_bufs, void *multicast, size_t bytes)
{
    if (!h || !peer_bufs) return h ? fail(h, KS_ERR_ARG, "ks_gather_attach: NULL argument") : KS_ERR_ARG;
    if (world < 1 || world > KS_MAX_WORLD || rank < 0 || rank >= world)
        return fail(h, KS_ERR_ARG, "ks_gather_attach: world=%d rank=%d (max world %d)", world, rank, KS_MAX_WORLD);
    if (h->g_buf) return fail(h, KS_ERR_STATE, "ks_gather_attach: gather already initialised");
    size_t slot = 0, total = 0;
    ks_gather_layout(h, world, &slot, &total);
    if (bytes < total) return fail(h, KS_ERR_ARG, "ks_gather_attach: buffers of %zu bytes, %zu needed", bytes, total);
    for (int r = 0; r < world; ++r)
        if (!peer_bufs[r]) return fail(h, KS_ERR_ARG, "ks_gather_attach: peer_bufs[%d] is NULL", r);
    DeviceGuard guard(h->cfg.device);
    h->g_world = world;
    h->g_rank = rank;
    h->g_slot = slot;
    h->g_flags_off = 2 * (size_t)world * slot;
    h->g_total = total;
    h->g_external = true;
    for (int r = 0; r < world; ++r) h->g_peer[r] = static_cast<uint8_t *>(peer_bufs[r]);
    h->g_buf = h->g_peer[rank];
    h->g_mc = static_cast<uint8_t *>(multicast);
    KS_CUDA(h, cudaMemset(h->g_buf, 0, h->g_total));
    KS_CUDA(h, cudaMalloc(&h->g_timeout, 2 * sizeof(int32_t) + KS_MAX_WORLD * sizeof(void *)));
    KS_CUDA(h, cudaMemset(h->g_timeout, 0, 2 * sizeof(int32_t) + KS_MAX_WORLD * sizeof(void *)));
    KS_CUDA(h, cudaHostAlloc((void **)&h->g_timeout_host, sizeof(int32_t), cudaHostAllocMapped));
    *h->g_timeout_host = 0;
    {
        double seconds = 120.0;
        if (const char *e = getenv("KS_GATHER_TIMEOUT_S")) {
            const double v = atof(e);
            if (v > 0.0) seconds = v;
        }
        int khz = 1965000;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->cfg.device);
        h->g_timeout_cycles = (long long)(seconds * 1e3 * (double)khz);
    }
    uint32_t *flags[KS_MAX_WORLD] = {nullptr};
    for (int r = 0; r < world; ++r) flags[r] = reinterpret_cast<uint32_t *>(h->g_peer[r] + h->g_flags_off);
    KS_CUDA(h, cudaMemcpy(h->g_timeout + 2, flags, sizeof(flags), cudaMemcpyHostToDevice));
    KS_CUDA(h, cudaDeviceSynchronize());
    h->g_connected = true;
    return KS_OK;
}

int ks_gather_status(ks_handle *h, int32_t *timed_out, void *stream_)
{
    if (!h || !timed_out) return KS_ERR_ARG;
    if (!h->g_timeout_host) return fail(h, KS_ERR_STATE, "ks_gather_status: gather not initialised");
    DeviceGuard guard(h->cfg.device);
    KS_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream_));
    *timed_out = *(volatile int32_t *)h->g_timeout_host != 0 ? 1 : 0;
    return KS_OK;
}

int ks_gather_clear(ks_handle *h, void *stream_)
{
    if (!h) return KS_ERR_ARG;
    if (!h->g_timeout_host) return fail(h, KS_ERR_STATE, "ks_gather_clear: gather not initialised");
    DeviceGuard guard(h->cfg.device);
    KS_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream_));
    *h->g_timeout_host = 0;
    return KS_OK;
}

int ks_collect(ks_handle *h, const ks_collect_args *a, void *stream_)
{
    if (!h || !a) return h ? fail(h, KS_ERR_ARG, "ks_collect: NULL argument") : KS_ERR_ARG;
    if (!a->actions || !a->obs || !a->reward || !a->truncated || !a->step || !a->obs_store || !a->act_store || !a->vminmax ||
        !a->agent_obs || !a->rec_obs || !a->rec_actions || !a->rec_nxtobs || !a->rec_reward || !a->rec_truncated || !a->rec_step)
        return fail(h, KS_ERR_ARG, "ks_collect: NULL buffer");
    if (a->agent_stride < 1 || a->agent_stride > h->obs_len) return fail(h, KS_ERR_ARG, "ks_collect: bad agent_stride");
    cudaStream_t stream = (cudaStream_t)stream_;
    DeviceGuard guard(h->cfg.device);
    const int B = h->cfg.num_envs, No = h->obs_len, J = h->cfg.J;
    const size_t n = (size_t)B * No;
    // The min / max keys of one call are re-armed by the NEXT call's apply kernel (parity slots).  A
    // call captured in a CUDA graph replays with the parity it was captured with; its keys then simply
    // keep accumulating, which is the running minimum / maximum the caller merges them into anyway.
    int32_t *keys = h->collect_keys + 2 * (h->collect_calls & 1u), *next = h->collect_keys + 2 * ((h->collect_calls + 1) & 1u);
    h->collect_calls += 1;
    const unsigned blocks = (unsigned)((n + 1023) / 1024 < 592 ? (n + 1023) / 1024 : 592);
    if (!a->frozen) {
        collect_minmax<<<blocks, 256, 0, stream>>>(a->obs, n, keys);
        KS_CUDA(h, cudaGetLastError());
        h->launches += 1;
    }
    collect_apply<<<blocks, 256, 0, stream>>>(*a, B, No, J, keys, next);
    KS_CUDA(h, cudaGetLastError());
    h->launches += 1;
    if (a->slot_index) {
        collect_advance<<<1, 1, 0, stream>>>(a->slot_index);
        KS_CUDA(h, cudaGetLastError());
        h->launches += 1;
    }
    return KS_OK;
}

}  // extern "C"
