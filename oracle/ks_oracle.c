/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference KS control period.
 *
 * Checker / CPU baseline, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the library built from this file.
 *
 * Parity pinning: validated against the reference's own outputs (tests/golden/ *.npz, made by
 * executing /root/reference through oracle/ref_loader.py) and against oracle/ks_numpy.py in
 * tests/test_oracle.py.  Compiled by oracle/Makefile with -ffp-contract=off so that fp64
 * arithmetic is not fused (NumPy, which the reference runs on, never fuses); the jet forcing
 * uses fmaf() explicitly because that is what torch's CPU sgemm does.
 *
 * Reference lines followed (relative to the reference root):
 *   kso_forcing      pdegym/common/transforms.py:262-265   phi = a @ F, float32
 *   kso_rhs          pdegym/kuramoto/kuramoto.py:24-27,118-129
 *   kso_step         pdegym/kuramoto/kuramoto.py:82-96      250 x RK4, reward of pre-step state
 *   reward modes     pdegym/kuramoto/kuramoto.py:64-73
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define KSO_MAX_N 4096

static const double UPWIND[5] = {-25.0 / 12.0, 4.0, -3.0, 4.0 / 3.0, -1.0 / 4.0};
static const double D2[7] = {1.0 / 90.0, -3.0 / 20.0, 3.0 / 2.0, -49.0 / 18.0, 3.0 / 2.0, -3.0 / 20.0,
                             1.0 / 90.0};
static const double D4[9] = {7.0 / 240.0, -2.0 / 5.0, 169.0 / 60.0, -122.0 / 15.0, 91.0 / 8.0,
                             -122.0 / 15.0, 169.0 / 60.0, -2.0 / 5.0, 7.0 / 240.0};

/* phi[b,i] = sequential-k fp32 FMA chain over the J jets, acc0 = 0 */
void kso_forcing(int B, int J, int N, const float *actions, const float *F, float *phi)
{
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            float acc = 0.0f;
            for (int k = 0; k < J; ++k) acc = fmaf(actions[b * J + k], F[k * N + i], acc);
            phi[b * N + i] = acc;
        }
}

/* derivative triple of one env.  q = work array of N doubles (receives u*u).
 * Periodic indexing goes through two halo-padded copies so that the stencil loops have fixed
 * offsets (no modulo) and vectorise; the order of floating-point operations per output point is
 * exactly the one of the straightforward loops (the compiler may not re-associate: no -ffast-math,
 * -ffp-contract=off). */
static void derivs(int N, double dx, const double *u, double *q, double *ux, double *uxx, double *uxxxx)
{
    const double dx2 = dx * dx, dx4 = dx2 * dx2;
    double up[KSO_MAX_N + 8], qp[KSO_MAX_N + 8];
    for (int i = 0; i < N; ++i) {
        q[i] = u[i] * u[i];
        up[i + 4] = u[i];
        qp[i + 4] = q[i];
    }
    for (int k = 0; k < 4; ++k) {
        up[k] = u[N - 4 + k];
        qp[k] = q[N - 4 + k];
        up[N + 4 + k] = u[k];
        qp[N + 4 + k] = q[k];
    }
    for (int i = 0; i < N; ++i) {
        const double *qc = qp + i + 4, *uc = up + i + 4;
        double fwd = UPWIND[0] * qc[0], bwd = -UPWIND[0] * qc[0];
        for (int k = 1; k < 5; ++k) {
            fwd = fwd + UPWIND[k] * qc[k];
            bwd = bwd - UPWIND[k] * qc[-k];
        }
        ux[i] = (uc[0] < 0.0 ? fwd : bwd) / dx;
        double s2 = 0.0, s4 = 0.0;
        for (int k = -3; k <= 3; ++k) s2 = s2 + D2[k + 3] * uc[k];
        for (int k = -4; k <= 4; ++k) s4 = s4 + D4[k + 4] * uc[k];
        uxx[i] = s2 / dx2;
        uxxxx[i] = s4 / dx4;
    }
}

static void rhs(int N, double dx, const double *u, const float *phi, double *out, double *w)
{
    double *q = w, *ux = w + N, *uxx = w + 2 * N, *uxxxx = w + 3 * N;
    derivs(N, dx, u, q, ux, uxx, uxxxx);
    for (int i = 0; i < N; ++i) out[i] = -uxxxx[i] - uxx[i] - 0.5 * ux[i] + (double)phi[i];
}

typedef struct {
    int b0, b1, N, cfg_steps, reward_mode;
    double dt, dx;
    double *u;
    const float *phi;
    double *reward;
} kso_job;

static void *kso_worker(void *arg)
{
    const kso_job *j = (const kso_job *)arg;
    const int N = j->N;
    const double dt = j->dt, dx = j->dx;
    double *w = (double *)malloc(sizeof(double) * (size_t)N * 9);
    double *k1 = w + 4 * N, *k2 = w + 5 * N, *k3 = w + 6 * N, *k4 = w + 7 * N, *us = w + 8 * N;
    for (int b = j->b0; b < j->b1; ++b) {
        double *ub = j->u + (size_t)b * N;
        const float *pb = j->phi + (size_t)b * N;
        double racc = 0.0;
        for (int s = 0; s < j->cfg_steps; ++s) {
            if (j->reward_mode == 0) {
                double ss = 0.0;
                for (int i = 0; i < N; ++i) ss += ub[i] * ub[i];
                double nrm = sqrt(ss);
                racc += -(1.0 / N) * (nrm * nrm);
            } else {
                derivs(N, dx, ub, w, w + N, w + 2 * N, w + 3 * N);
                double a = 0.0, c = 0.0, d = 0.0;
                for (int i = 0; i < N; ++i) {
                    a += w[2 * N + i] * w[2 * N + i];
                    c += w[N + i] * w[N + i];
                    d += ub[i] * (double)pb[i];
                }
                racc += -(a / N + c / N + d / N);
            }
            rhs(N, dx, ub, pb, k1, w);
            for (int i = 0; i < N; ++i) us[i] = ub[i] + dt * k1[i] / 2.0;
            rhs(N, dx, us, pb, k2, w);
            for (int i = 0; i < N; ++i) us[i] = ub[i] + dt * k2[i] / 2.0;
            rhs(N, dx, us, pb, k3, w);
            for (int i = 0; i < N; ++i) us[i] = ub[i] + dt * k3[i];
            rhs(N, dx, us, pb, k4, w);
            for (int i = 0; i < N; ++i)
                ub[i] = ub[i] + dt * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) / 6.0;
        }
        j->reward[b] = racc / j->cfg_steps;
    }
    free(w);
    return NULL;
}

static int g_threads = 0; /* 0 = all online cores */

void kso_set_threads(int n) { g_threads = n; }

int kso_num_threads(void)
{
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* One control period for B independent envs (split over pthreads by env index).
 * u [B,N] in/out, phi [B,N] float32, reward [B].
 * reward_mode 0: -(1/N)*||u||^2 ; 1: -(mean(uxx^2)+mean(ux^2)+mean(u*phi)).
 * Returns 0, or -1 on bad arguments. */
int kso_step(int B, int N, int cfg_steps, double dt, double dx, int reward_mode, double *u,
             const float *phi, double *reward)
{
    if (N < 9 || N > KSO_MAX_N || B < 0 || cfg_steps < 1) return -1;
    int T = kso_num_threads();
    if (T > B) T = B > 0 ? B : 1;
    if (T > 256) T = 256;
    pthread_t tid[256];
    kso_job job[256];
    for (int t = 0; t < T; ++t) {
        kso_job j = {(int)((long long)B * t / T), (int)((long long)B * (t + 1) / T), N, cfg_steps,
                     reward_mode, dt, dx, u, phi, reward};
        job[t] = j;
    }
    for (int t = 1; t < T; ++t) pthread_create(&tid[t], NULL, kso_worker, &job[t]);
    kso_worker(&job[0]);
    for (int t = 1; t < T; ++t) pthread_join(tid[t], NULL);
    return 0;
}

/* K periods with per-period actions [K,B,J]; optional per-period outputs (may be NULL):
 * obs_out [K,B,N] float32 (cast of the state), reward_out [K,B]. */
int kso_rollout(int K, int B, int N, int J, int cfg_steps, double dt, double dx, int reward_mode,
                double *u, const float *actions, const float *F, float *obs_out, double *reward_out)
{
    float *phi = (float *)malloc(sizeof(float) * (size_t)B * N);
    double *rew = (double *)malloc(sizeof(double) * (size_t)B);
    if (!phi || !rew) return -2;
    int rc = 0;
    for (int k = 0; k < K && rc == 0; ++k) {
        if (actions) kso_forcing(B, J, N, actions + (size_t)k * B * J, F, phi);
        else memset(phi, 0, sizeof(float) * (size_t)B * N);
        rc = kso_step(B, N, cfg_steps, dt, dx, reward_mode, u, phi, rew);
        if (reward_out) memcpy(reward_out + (size_t)k * B, rew, sizeof(double) * (size_t)B);
        if (obs_out)
            for (size_t i = 0; i < (size_t)B * N; ++i) obs_out[(size_t)k * B * N + i] = (float)u[i];
    }
    free(phi);
    free(rew);
    return rc;
}
