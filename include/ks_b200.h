/* ks_b200.h -- C ABI of libks_b200.so: batched Kuramoto-Sivashinsky control environment,
 * hand-written CUDA for NVIDIA B200 (sm_100a).
 *
 * The reference (stwerner97/model-based-pde-control) is pure Python and has no FFI; the
 * interface this library replaces is the gym 0.25.2 Env / VectorEnv protocol as exercised by
 *   pdegym/kuramoto/kuramoto.py:78-116          KuramotoSivashinskyEnv.step / reset
 *   pdegym/common/transforms.py:250-265         GaussianForcing (jet actuation, float32)
 *   pdegym/kuramoto/__init__.py:8-12            make(): TimeLimit(env, 400)
 *   pdecontrol/mbrl/mbrl.py:81-86               gym.vector.make(env_id, num_envs=cpus)
 *   pdecontrol/mbrl/worker.py:48-66             envs.reset() / envs.step(actions)
 * Each entry point below cites the reference lines it stands in for.  The Python mirror of the
 * gym surface (KSVecEnv) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns int: 0 = OK, < 0 = argument / state error (the device is not
 *     touched), > 0 = a cudaError_t value.  ks_last_error() gives a message for the last
 *     non-zero return on that handle (or on creation, with a NULL handle).
 *   - no C++ exceptions, callbacks or global mutable state cross the boundary; no allocations
 *     after ks_create.
 *   - the caller owns every buffer it passes.  "dev" pointers are device memory on the
 *     handle's device; "host" pointers are ordinary (ideally pinned) host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Device
 *     entry points are asynchronous and stream-ordered; *_host entry points synchronise the
 *     stream before returning.
 *   - a handle is bound to one device and is not thread-safe.  Multi-GPU = one process and one
 *     handle per GPU; envs are independent, so no collective is needed inside the library.
 *   - there is no CPU fallback: without a usable CUDA device ks_create fails.
 */
#ifndef KS_B200_H
#define KS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KS_ABI_VERSION 3

enum ks_precision { KS_F64 = 0, KS_F32 = 1 };
/* KS_REWARD_L2: what the reference executes, -(1/N)*||u||^2 (kuramoto.py:64-65,72).
 * KS_REWARD_DISSIPATION: the intended -(mean(uxx^2)+mean(ux^2)+mean(u*phi)) (kuramoto.py:67-70). */
enum ks_reward_mode { KS_REWARD_L2 = 0, KS_REWARD_DISSIPATION = 1 };
enum ks_where { KS_HOST = 0, KS_DEVICE = 1 };
/* KS_SOLVER_FD_RK4: the reference's scheme -- periodic finite differences + classic RK4
 *   (kuramoto.py:83-90,118-129); the parity path (<= 1e-10 against the reference).
 * KS_SOLVER_ETDRK4: pseudo-spectral exponential integrator (Cox & Matthews 2002; contour-integral
 *   coefficients after Kassam & Trefethen 2005, precomputed on the host in ks_create), hand-written
 *   in-register / warp-level FFT.  NOT in the reference: it integrates the same equation
 *   (kuramoto.py:127) with a different discretisation, `dt` and `cfg_steps` are then the ETDRK4
 *   step and the steps per control period (e.g. dt = 0.025, cfg_steps = 10 for the reference's
 *   0.25 time units).  N = 64, 128 or 256.  KS_REWARD_DISSIPATION uses spectral derivatives
 *   (uxx of u, and -- literally as kuramoto.py:67-70,120-122 -- the derivative of u^2 for "ux"). */
enum ks_solver { KS_SOLVER_FD_RK4 = 0, KS_SOLVER_ETDRK4 = 1 };

enum ks_error {
    KS_OK = 0,
    KS_ERR_ARG = -1,         /* NULL / out-of-range argument */
    KS_ERR_UNSUPPORTED = -2, /* grid size N cannot be mapped to the register kernel */
    KS_ERR_NO_DEVICE = -3,   /* no CUDA device / wrong architecture */
    KS_ERR_STATE = -4        /* call not valid in the handle's current state */
};

/* Constants of KuramotoSivashinskyEnv.__init__ (kuramoto.py:29-57) plus batch geometry. */
typedef struct ks_config {
    int32_t abi_version;       /* = KS_ABI_VERSION */
    int32_t num_envs;          /* B: independent environments owned by this handle */
    int32_t N;                 /* grid points (reference default 64) */
    int32_t J;                 /* jets = len(Xi) (reference default 4), 0 < J <= 32 */
    int32_t cfg_steps;         /* RK4 sub-steps per control period (250) */
    int32_t max_episode_steps; /* ceil(Tmax/(dt*cfg_steps)) = 400 (kuramoto.py:57) */
    int32_t burnin_periods;    /* int(200/dt/cfg_steps) = 800 (kuramoto.py:103) */
    int32_t precision;         /* enum ks_precision */
    int32_t reward_mode;       /* enum ks_reward_mode */
    int32_t device;            /* CUDA device ordinal */
    int32_t points_per_lane;   /* 0 = choose automatically; else P with N % P == 0, 2<=P<=16 */
    int32_t obs_stride;        /* SensorTransform stride s: obs = u[s/2::s] (transforms.py:236-239);
                                  0 or 1 = full state (what the MBRL loop uses, mbrl.py:171,174) */
    int32_t solver;            /* enum ks_solver; 0 = the reference's FD-RK4 scheme */
    int32_t dealias;           /* KS_SOLVER_ETDRK4 only: 1 = 2/3-rule dealiasing of (u^2)_x, 0 = none */
    int32_t env_index_base;    /* global index of this handle's env 0 (a sharded run: the shard's first env);
                                  enters the Philox counter of device-drawn initial conditions so that a
                                  sharded reset draws exactly what the single-GPU reset draws.  0 otherwise */
    int32_t reserved0;         /* = 0 (keeps the doubles 8-byte aligned) */
    double L;                  /* domain length (22.0) */
    double dt;                 /* RK4 step (1e-3) */
    const float *forcing;      /* host, [J*N] row-major float32: GaussianForcing.forcing
                                  (transforms.py:258-260), built by the caller with torch so that
                                  it is bit-identical to the reference's matrix */
} ks_config;

typedef struct ks_handle ks_handle;

/* Replaces KuramotoSivashinskyEnv(**config) x num_envs (kuramoto.py:29-76; mbrl.py:81-86).
 * Allocates u [B,N], timestep [B], flags [B] and the packed output block on `device`; state
 * starts at u = 0, timestep = 0. */
int ks_create(const ks_config *cfg, ks_handle **out);
int ks_destroy(ks_handle *h); /* idempotent on NULL */

/* env.u = ...; env.timestep = ... (plain attributes in the reference; kuramoto.py:105-106).
 * u is [B,N] float64 whatever the precision (converted on device in KS_F32 mode); either
 * pointer may be NULL to leave that part unchanged.  Clears the non-finite flags. */
int ks_set_state(ks_handle *h, const double *u, const int32_t *timestep, int where, void *stream);
int ks_get_state(ks_handle *h, double *u, int32_t *timestep, int where, void *stream);

/* KuramotoSivashinskyEnv.reset (kuramoto.py:100-116): initial condition, then `burnin_periods`
 * no-op control periods in ONE launch, then timestep = 0 -- for every env, or only for the envs
 * with mask[b] != 0 when `mask` is given (gym's vector env resets sub-envs individually when they
 * truncate).  `where` says where u0 and mask live (KS_HOST / KS_DEVICE).
 *   u0 != NULL: [B,N] float64 initial condition used verbatim (rows of unmasked envs ignored);
 *               host NumPy draws give parity with np.random.seed(seed); np.random.uniform(-0.4,0.4,N);
 *   u0 == NULL: U(-0.4,0.4) from a counter-based Philox generator keyed by (seed; point, env).
 * burnin_periods < 0 uses the configured value; 0 skips the burn-in. */
int ks_reset(ks_handle *h, const double *u0, const uint8_t *mask, int where, uint64_t seed,
             int32_t burnin_periods, void *stream);

/* KuramotoSivashinskyEnv.step for all B envs (kuramoto.py:78-98), one launch, device buffers.
 *   actions  dev [B,J] float32 (np.array(action, float32), kuramoto.py:79)
 *   phi      dev [B,N] float32 or NULL; non-NULL overrides the in-kernel `a @ F` FMA chain
 *   obs      dev [B,No] float32: float32 cast of the new state (what gym's vector env hands on),
 *                               No = N, or ceil((N - s/2)/s) sensors when obs_stride = s > 1
 *   reward   dev [B]   float64: mean over sub-steps of the pre-step reward (kuramoto.py:84,96)
 *   truncated dev [B]  uint8:   timestep >= max_episode_steps (kuramoto.py:93)
 *   step     dev [B]   int32:   info["step"] (kuramoto.py:98)
 *   nonfinite dev [B]  uint8:   1 once the env's state left the finite range (np.seterr(over=
 *                               "raise"), kuramoto.py:12, raises FloatingPointError at that point)
 * Any output pointer may be NULL.  No auto-reset here: the caller (KSVecEnv) does it. */
int ks_step(ks_handle *h, const float *actions, const float *phi, float *obs, double *reward,
            uint8_t *truncated, int32_t *step, uint8_t *nonfinite, void *stream);

/* K consecutive control periods in one persistent launch (open loop, no auto-reset):
 * actions dev [K,B,J] (NULL = no-op periods); outputs dev [K,B,...] or NULL. */
int ks_rollout(ks_handle *h, int32_t K, const float *actions, float *obs, double *reward,
               uint8_t *truncated, int32_t *step, uint8_t *nonfinite, void *stream);

/* Host-buffer form of ks_step -- the call the gym-facing wrapper makes when the policy lives on
 * the host (worker.py:60-66).  With pinned (device-addressable) buffers the step is ONE launch and a
 * synchronise: the kernel reads the actions straight from host memory and its epilogue mirrors
 * every output into the host block over PCIe (the mirrored-store path of the multi-GPU exchange).
 * With pageable buffers, or KS_HOST_IO=copy in the environment, it copies actions host->device,
 * runs the period and copies ONE packed output block device->host.  Either way it synchronises
 * `stream` before returning.  Layout of the block (see ks_out_layout):
 *   reward f64 [B] | obs f32 [B*No] | step i32 [B] | truncated u8 [B] | nonfinite u8 [B]
 * (each part 16-byte aligned). */
int ks_step_host(ks_handle *h, const float *actions_host, void *out_host, void *stream);
/* offsets[5] = byte offsets of {reward, obs, step, truncated, nonfinite}; *total = block bytes. */
int ks_out_layout(const ks_handle *h, size_t offsets[5], size_t *total);

/* Fused all-gather of the packed output block (ks_out_layout) across the ranks of a sharded run
 * -- the exchange that lets every learner rank see the whole batch (the reference's counterpart is
 * gym's AsyncVectorEnv collecting the sub-env results in the parent, mbrl.py:81-86; worker.py:60-66).
 * Instead of a separate collective after the kernel, the period kernel's epilogue stores every
 * output both locally and, with plain st.global over NVLink, into this rank's slot of every peer's
 * gather buffer (CUDA-IPC mapped), followed by a one-warp epoch handshake.  One process per GPU:
 *   1. every rank: ks_gather_init(h, world, rank, handle[64], &slot_bytes)   allocates its buffer
 *   2. exchange the 64-byte handles by any means (e.g. torch.distributed.all_gather_object)
 *   3. every rank: ks_gather_connect(h, all_handles = [world][64] in rank order)
 *   4. per period: ks_step_gather(h, actions_dev, &gathered, stream); `gathered` is a device
 *      pointer to [world][slot_bytes]: rank r's packed block at r*slot_bytes, valid (for work
 *      enqueued on `stream`) until the SECOND next ks_step_gather (double-buffered).
 * All ranks must call ks_step_gather the same number of times.  The handshake waits for every peer
 * for up to KS_GATHER_TIMEOUT_S seconds (environment, default 120; ranks of a training job drift apart
 * by seconds).  If a peer still has not signalled, the launch sets a sticky error word and poisons
 * that peer's slot (its non-finite flags read 0xFF); every later ks_step_gather then FAILS with
 * KS_ERR_STATE instead of handing out incomplete blocks, ks_gather_status reports timed_out = 1, and
 * ks_gather_clear (after the application has re-synchronised its ranks) re-arms the exchange. */
#define KS_MAX_WORLD 16
#define KS_IPC_HANDLE_BYTES 64
int ks_gather_init(ks_handle *h, int32_t world, int32_t rank, void *ipc_handle_out, size_t *slot_bytes);
int ks_gather_connect(ks_handle *h, const void *all_handles);
int ks_step_gather(ks_handle *h, const float *actions, void **gathered, void *stream);
int ks_gather_status(ks_handle *h, int32_t *timed_out, void *stream);
int ks_gather_clear(ks_handle *h, void *stream);
/* Stream-ordered rendezvous of all ranks on the device (the exchange's one-warp signal / wait kernel with its own
 * flag words and epoch counter, no payload): work enqueued after it starts within about one NVLink flag flight on
 * every rank.  bench.py uses it to re-align the ranks between timed steps; all ranks must call it equally often. */
int ks_gather_barrier(ks_handle *h, void *stream);
/* Alternative set-up on buffers the CALLER has allocated and mapped -- any symmetric-memory mechanism
 * (torch.distributed._symmetric_memory, NVSHMEM, cuMem* with fabric handles) instead of steps 1-3 above:
 *   ks_gather_layout(h, world, &slot_bytes, &total_bytes)   size every rank's buffer must have
 *   ks_gather_attach(h, world, rank, peer_bufs, multicast, bytes)
 * peer_bufs[r] = rank r's buffer as mapped into THIS process (peer_bufs[rank] = the local one, which this call
 * zero-fills; the caller synchronises the ranks before the first ks_step_gather).  `multicast` (nullable) = an NVLS
 * multicast address bound to all of those buffers: the FD-RK4 period kernel then sends its observation rows ONCE
 * with multimem.st (replicated by the NVSwitch) instead of once per peer; the small per-env outputs and the
 * handshake stay unicast.  The library neither frees nor unmaps attached buffers. */
int ks_gather_layout(const ks_handle *h, int32_t world, size_t *slot_bytes, size_t *total_bytes);
int ks_gather_attach(ks_handle *h, int32_t world, int32_t rank, void *const *peer_bufs, void *multicast, size_t bytes);

/* One step of the reference's data-collection plumbing, fused (two small kernels), for stores of
 * length 1 as the MBRL loop uses them (mbrl.py:257-275): what StoreNObsVecWrapper.step_wait
 * (vec_wrappers.py:21-37), StoreNActionsVecWrapper.step_async (:65-72), TransformObsWrapper.step_wait with
 * a running-min/max ScaleTransform and a SensorTransform (:152-171; transforms.py:141-247) and the
 * bookkeeping of Worker.rollout (worker.py:68-88) do per env step, after ks_step has written its outputs:
 *   vminmax    <- min / max over everything seen so far (unless frozen)
 *   rec_obs    <- obs_store (the stored observation before the step);  rec_nxtobs, obs_store <- obs
 *   agent_obs  <- ((obs - vmin) / (vmax - vmin)) * scale_width + lower, sampled at [stride/2::stride]
 *   rec_actions, act_store <- actions;  rec_reward / rec_truncated / rec_step <- reward / truncated / step
 * All pointers are device memory; float32 arithmetic in exactly this order (no contraction), so the
 * result equals the element-wise torch / NumPy evaluation bit for bit.  Episode-end steps (final
 * observation, auto-reset) are handled by the caller. */
typedef struct ks_collect_args {
    const float *actions;     /* [B,J]   env-scale actions of this step */
    const float *obs;         /* [B,No]  ks_step's observation output */
    const double *reward;     /* [B] */
    const uint8_t *truncated; /* [B] */
    const int32_t *step;      /* [B] */
    float *obs_store;         /* [B,No]  in / out */
    float *act_store;         /* [B,J]   out */
    float *vminmax;           /* [2]     running min, max: in / out */
    float *agent_obs;         /* [B,No'] out, No' = len(range(stride/2, No, stride)) */
    float *rec_obs, *rec_actions, *rec_nxtobs; /* slot of the transition buffers: [B,No], [B,J], [B,No] */
    double *rec_reward;       /* [B] */
    uint8_t *rec_truncated;   /* [B] (bool) */
    int64_t *rec_step;        /* [B] */
    float lower, scale_width; /* target range: lower bound and (upper - lower) */
    int32_t frozen;           /* != 0: vminmax is not updated */
    int32_t agent_stride;     /* SensorTransform stride on the agent's observation (1 = all) */
    int64_t *slot_index;      /* NULL: rec_* point at the slot itself.  Else a device counter t: rec_* are
                                 the bases of [T,B,...] buffers, slot t is written and t incremented by a
                                 third one-thread kernel -- this form can be captured in a CUDA graph */
} ks_collect_args;
int ks_collect(ks_handle *h, const ks_collect_args *args, void *stream);

/* Mirrors np.seterr(over="raise") (kuramoto.py:12): flags[b] != 0 once env b's state went
 * non-finite (sticky until ks_reset / ks_set_state).  Synchronises `stream`.  *any (nullable)
 * receives the OR of all flags; nonfinite_host (nullable) the [B] flags. */
int ks_status(ks_handle *h, uint8_t *nonfinite_host, int32_t *any, void *stream);

/* Batched evaluation of rhs() and the reward on arbitrary states, for callers that use
 * env.rhs / env.reward_func offline (training.py:211-236; mbrl/world/world.py:170):
 * u dev [M,N] f64, phi dev [M,N] f32; outputs (each nullable) dev [M,N] f64 / reward dev [M] f64
 * with the handle's reward_mode.  Always float64. */
int ks_eval(ks_handle *h, int32_t M, const double *u, const float *phi, double *rhs, double *ux,
            double *uxx, double *uxxxx, double *reward, void *stream);

/* Measurement helper (not part of the env path): FP64 FMA throughput of `device` in TFLOP/s from
 * a register-resident DFMA loop, best and mean over `repeats` launches of `iters` iterations x 16
 * independent chains per thread.  Used by bench.py as the self-measured FP64 roofline peak. */
int ks_bench_fp64_peak(int device, int iters, int repeats, double *tflops_best, double *tflops_mean);
/* The same with FFMA: the roofline denominator of the optional fp32 mode. */
int ks_bench_fp32_peak(int device, int iters, int repeats, double *tflops_best, double *tflops_mean);

/* Introspection */
int ks_get_config(const ks_handle *h, ks_config *out); /* forcing pointer is returned NULL */
int ks_launch_info(const ks_handle *h, int32_t *points_per_lane, int32_t *lanes_per_env,
                   int32_t *block_threads, int32_t *grid_blocks, int32_t *regs_per_thread);
uint64_t ks_launch_count(const ks_handle *h); /* kernels launched through this handle so far */
const char *ks_last_error(const ks_handle *h);
int ks_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* KS_B200_H */
