/* TEST INFRASTRUCTURE ONLY - plain-C FP64 restatement of the reference Kalman-filter hot path.
 *
 * Not product code: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs call it
 * (through oracle/c_oracle.py).  It restates, in execution order (paths under /root/reference):
 *
 *   kalman_filter/kalman_filter.py:79-105   get_odom            -> form_measurement()
 *   kalman_filter/kalman_filter.py:108-117  set_measurements    -> form_measurement()
 *   kalman_filter/kalman_filter.py:184-193  rotation_matrix_body_world -> rot_zyx()
 *   misc/force_controller.py:269-291        next_state          -> propagate_mean()
 *   kalman_filter/kalman_filter.py:119-138  predict             -> propagate_cov(), model 0
 *   kalman_filter/kalman_filter.py:153-158  predict_mpc (cov)   -> propagate_cov(), model 1
 *   kalman_filter/kalman_filter.py:164-174  update              -> joint_update()
 *
 * Dense 12x12 / 10x10 arithmetic in the reference's operand order: K = (P H^T) inv(S) with a
 * pivoted LU inverse (np.linalg.inv -> LAPACK getrf/getri), P <- (I - K H) P, no symmetrisation.
 * Multiplications by the structural zeros/ones of H are replaced by the equivalent row/column
 * selections (bit-identical apart from the sign of zero).
 *
 * Parity pin: tests/golden/ (outputs of the unmodified reference class, oracle/gen_golden.py).
 *
 * Data layout matches the device layout: per-step inputs are [T][C][S] (stream index fastest),
 * per-trajectory arrays are [C][N] (trajectory index fastest).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NX 12
#define NZ 10

static const int SEL[NZ] = {0, 1, 2, 5, 6, 7, 8, 9, 10, 11}; /* kalman_filter.py:15-24 */

typedef struct KfOracleArgs {
    int64_t n_traj, n_steps, n_streams;
    int32_t cov_model;      /* 0: F_d = I + dt F (predict)   1: F_d = exp(dt F) element-wise (predict_mpc) */
    int32_t q_kind, r_kind; /* 0: one dense matrix shared [n*n]   1: per-trajectory diagonal [n][N] */
    int32_t x0_per_traj;    /* 0: x0 is [12] shared   1: [12][N] */
    int32_t p0_kind;        /* 0: P0 = Q   1: dense shared [144]   2: dense per trajectory [144][N] */
    int32_t n_threads;      /* <=0: OpenMP default */
    double dt, mass, inertia[3], gravity;
    const double *imu, *p, *dp, *contact, *f; /* [T][6|12|12|4|12][S] */
    const double *body_ref;                   /* [T][12][S], cov_model 1 only */
    const int32_t *stream_index;              /* [N] or NULL -> stream = trajectory % S */
    const double *x0, *P0, *Q, *R;
    int64_t ckpt_every; /* 0: none */
    /* outputs, any may be NULL */
    double *x_steps, *x_model_steps, *p_world_steps; /* [T][12][N] */
    double *z_steps;                                  /* [T][10][N] */
    double *p_trace_steps, *k_gain_steps, *nis_steps; /* [T][N] */
    double *P_ckpt;                                   /* [T/ckpt_every][144][N] */
    double *x_final, *P_final, *K_final;              /* [12][N], [144][N], [120][N] */
    uint32_t *status;                                 /* [N] bit0: singular S, bit1: non-finite, bit2: all-swing step */
} KfOracleArgs;

static void rot_zyx(double a, double b, double c, double R[3][3]) {
    const double sa = sin(a), ca = cos(a), sb = sin(b), cb = cos(b), sc = sin(c), cc = cos(c);
    R[0][0] = cc * cb; R[0][1] = cc * (sb * sa) - sc * ca; R[0][2] = cc * (sb * ca) + sc * sa;
    R[1][0] = sc * cb; R[1][1] = sc * (sb * sa) + cc * ca; R[1][2] = sc * (sb * ca) - cc * sa;
    R[2][0] = -sb;     R[2][1] = cb * sa;                  R[2][2] = cb * ca;
}

/* returns 1 when no leg is in stance (the reference raises ValueError there) */
static int form_measurement(const double imu[6], const double p[12], const double dp[12], const double contact[4],
                            double z[NZ]) {
    double nc = 0, sx = 0, sy = 0, sv = 0, sz = 0;
    for (int l = 0; l < 4; ++l) {
        nc += contact[l];
        if (contact[l] == 1.0) { sx += dp[3 * l]; sy += dp[3 * l + 1]; sz += p[3 * l + 2]; }
        if (contact[l] == 0.0) { sv += dp[3 * l + 2]; }
    }
    int all_swing = (nc == 0.0);
    double vb[3] = {0, 0, 0}, zo = 0;
    if (!all_swing) { vb[0] = -1 * sx / nc; vb[1] = -1 * sy / nc; vb[2] = -1 * sv / nc; zo = -1 * sz / nc; }
    double R[3][3];
    rot_zyx(imu[0], imu[1], imu[2], R);
    z[0] = imu[0]; z[1] = imu[1]; z[2] = imu[2]; z[3] = zo;
    z[4] = imu[3]; z[5] = imu[4]; z[6] = imu[5];
    for (int i = 0; i < 3; ++i) z[7 + i] = R[i][0] * vb[0] + R[i][1] * vb[1] + R[i][2] * vb[2];
    return all_swing;
}

static void inv3(const double A[3][3], double B[3][3]) {
    const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    const double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
    const double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    const double id = 1.0 / det;
    B[0][0] = c00 * id; B[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id; B[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
    B[1][0] = c01 * id; B[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id; B[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
    B[2][0] = c02 * id; B[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id; B[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
}

/* next_state (force_controller.py:269-291): x <- (I + A dt) x + (B dt) f + dt g, p rotated in place */
static void propagate_mean(const KfOracleArgs *a, double x[NX], double p[12], const double f[12], double R[3][3]) {
    rot_zyx(x[0], x[1], x[2], R);
    double Ihat[3][3], Iinv[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += R[i][k] * (a->inertia[k] * R[j][k]);
            Ihat[i][j] = s;
        }
    inv3(Ihat, Iinv);
    for (int l = 0; l < 4; ++l) {
        double pw[3];
        for (int i = 0; i < 3; ++i) pw[i] = R[i][0] * p[3 * l] + R[i][1] * p[3 * l + 1] + R[i][2] * p[3 * l + 2];
        p[3 * l] = pw[0]; p[3 * l + 1] = pw[1]; p[3 * l + 2] = pw[2];
    }
    double xn[NX];
    const double dt = a->dt;
    for (int i = 0; i < 3; ++i) {
        /* int64 A: the attitude rows carry trunc(R^T) (force_controller.py:248-251,271) */
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (trunc(R[k][i]) * dt) * x[6 + k];
        xn[i] = x[i] + s;
        xn[3 + i] = x[3 + i] + dt * x[9 + i];
    }
    double dw[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
    for (int l = 0; l < 4; ++l) {
        const double *pw = p + 3 * l, *fl = f + 3 * l;
        const double sk[3][3] = {{0, -pw[2], pw[1]}, {pw[2], 0, -pw[0]}, {-pw[1], pw[0], 0}};
        for (int i = 0; i < 3; ++i) {
            double row = 0;
            for (int j = 0; j < 3; ++j) {
                double b = 0; /* B[6+i][3l+j] = (Iinv skew)_ij */
                for (int k = 0; k < 3; ++k) b += Iinv[i][k] * sk[k][j];
                row += (b * dt) * fl[j];
            }
            dw[i] += row;
            dv[i] += ((1.0 / a->mass) * dt) * fl[i];
        }
    }
    for (int i = 0; i < 3; ++i) { xn[6 + i] = x[6 + i] + dw[i]; xn[9 + i] = x[9 + i] + dv[i]; }
    xn[11] += dt * a->gravity;
    memcpy(x, xn, sizeof xn);
}

static void matmul12(const double *A, const double *B, double *C, int transpose_b) {
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NX; ++j) {
            double s = 0;
            for (int k = 0; k < NX; ++k) s += A[i * NX + k] * (transpose_b ? B[j * NX + k] : B[k * NX + j]);
            C[i * NX + j] = s;
        }
}

static void propagate_cov(const KfOracleArgs *a, double *P, const double R[3][3], const double *Q) {
    double F[NX * NX], Fd[NX * NX], W[NX * NX];
    memset(F, 0, sizeof F);
    for (int i = 0; i < 3; ++i) {
        F[(3 + i) * NX + 9 + i] = 1.0;
        for (int j = 0; j < 3; ++j) F[i * NX + 6 + j] = R[j][i];
    }
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NX; ++j)
            Fd[i * NX + j] = a->cov_model ? exp(a->dt * F[i * NX + j]) : ((i == j) ? 1.0 : 0.0) + a->dt * F[i * NX + j];
    matmul12(Fd, P, W, 0);
    matmul12(W, Fd, P, 1);
    for (int i = 0; i < NX * NX; ++i) P[i] += Q[i];
}

/* inverse by LU with partial pivoting (what np.linalg.inv does through getrf/getri); returns 1 if singular */
static int inv10(const double *S, double *Sinv) {
    double A[NZ][2 * NZ];
    for (int i = 0; i < NZ; ++i)
        for (int j = 0; j < NZ; ++j) { A[i][j] = S[i * NZ + j]; A[i][NZ + j] = (i == j); }
    for (int c = 0; c < NZ; ++c) {
        int piv = c;
        for (int r = c + 1; r < NZ; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (A[piv][c] == 0.0 || !isfinite(A[piv][c])) return 1;
        if (piv != c) for (int j = 0; j < 2 * NZ; ++j) { double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        for (int r = c + 1; r < NZ; ++r) {
            const double m = A[r][c] / A[c][c];
            for (int j = c; j < 2 * NZ; ++j) A[r][j] -= m * A[c][j];
        }
    }
    for (int c = NZ - 1; c >= 0; --c) {
        for (int j = 0; j < NZ; ++j) {
            double s = A[c][NZ + j];
            for (int k = c + 1; k < NZ; ++k) s -= A[c][k] * Sinv[k * NZ + j];
            Sinv[c * NZ + j] = s / A[c][c];
        }
    }
    return 0;
}

static int joint_update(double x[NX], double *P, const double z[NZ], const double *Rn, double *K, double *p_trace,
                        double *k_gain, double *nis) {
    double y[NZ], S[NZ * NZ], Sinv[NZ * NZ];
    for (int i = 0; i < NZ; ++i) {
        y[i] = z[i] - x[SEL[i]];
        for (int j = 0; j < NZ; ++j) S[i * NZ + j] = P[SEL[i] * NX + SEL[j]] + Rn[i * NZ + j];
    }
    if (inv10(S, Sinv)) return 1;
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NZ; ++j) {
            double s = 0;
            for (int k = 0; k < NZ; ++k) s += P[i * NX + SEL[k]] * Sinv[k * NZ + j];
            K[i * NZ + j] = s;
        }
    double q = 0;
    for (int i = 0; i < NZ; ++i) {
        double s = 0;
        for (int j = 0; j < NZ; ++j) s += Sinv[i * NZ + j] * y[j];
        q += y[i] * s;
    }
    *nis = q;
    for (int i = 0; i < NX; ++i) {
        double s = 0;
        for (int j = 0; j < NZ; ++j) s += K[i * NZ + j] * y[j];
        x[i] += s;
    }
    double M[NX * NX], Pn[NX * NX]; /* M = I - K H */
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NX; ++j) M[i * NX + j] = (i == j) ? 1.0 : 0.0;
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NZ; ++j) M[i * NX + SEL[j]] -= K[i * NZ + j];
    matmul12(M, P, Pn, 0);
    memcpy(P, Pn, sizeof Pn);
    double tr = 0, kg = 0;
    for (int i = 0; i < NX; ++i) tr += P[i * NX + i];
    for (int i = 0; i < NZ; ++i) kg += K[i * NZ + i]; /* np.trace of the 12x10 K */
    *p_trace = tr; *k_gain = kg;
    return 0;
}

static void run_one(const KfOracleArgs *a, int64_t i) {
    const int64_t N = a->n_traj, S = a->n_streams, T = a->n_steps;
    const int64_t s = a->stream_index ? a->stream_index[i] : (i % S);
    double x[NX], P[NX * NX], Q[NX * NX], Rn[NZ * NZ], K[NX * NZ];
    memset(K, 0, sizeof K);
    if (a->q_kind == 0) memcpy(Q, a->Q, sizeof Q);
    else { memset(Q, 0, sizeof Q); for (int c = 0; c < NX; ++c) Q[c * NX + c] = a->Q[c * N + i]; }
    if (a->r_kind == 0) memcpy(Rn, a->R, sizeof Rn);
    else { memset(Rn, 0, sizeof Rn); for (int c = 0; c < NZ; ++c) Rn[c * NZ + c] = a->R[c * N + i]; }
    for (int c = 0; c < NX; ++c) x[c] = a->x0_per_traj ? a->x0[c * N + i] : a->x0[c];
    if (a->p0_kind == 0) memcpy(P, Q, sizeof P);
    else if (a->p0_kind == 1) memcpy(P, a->P0, sizeof P);
    else for (int c = 0; c < NX * NX; ++c) P[c] = a->P0[c * N + i];
    uint32_t status = 0;
    double ptr = 0, kg = 0, nis = 0;
    for (int c = 0; c < NX; ++c) ptr += P[c * NX + c];
    for (int64_t t = 0; t < T; ++t) {
        double imu[6], p[12], dp[12], contact[4], f[12], z[NZ], R[3][3];
        for (int c = 0; c < 6; ++c) imu[c] = a->imu[(t * 6 + c) * S + s];
        for (int c = 0; c < 12; ++c) { p[c] = a->p[(t * 12 + c) * S + s]; dp[c] = a->dp[(t * 12 + c) * S + s]; f[c] = a->f[(t * 12 + c) * S + s]; }
        for (int c = 0; c < 4; ++c) contact[c] = a->contact[(t * 4 + c) * S + s];
        if (form_measurement(imu, p, dp, contact, z)) status |= 4u;
        propagate_mean(a, x, p, f, R);
        if (a->cov_model == 1) {
            const double *br = a->body_ref + (t * 12) * S + s;
            rot_zyx(br[0], br[S], br[2 * S], R);
        }
        propagate_cov(a, P, R, Q);
        if (a->x_model_steps) for (int c = 0; c < NX; ++c) a->x_model_steps[(t * NX + c) * N + i] = x[c];
        if (a->p_world_steps) for (int c = 0; c < 12; ++c) a->p_world_steps[(t * 12 + c) * N + i] = p[c];
        if (a->z_steps) for (int c = 0; c < NZ; ++c) a->z_steps[(t * NZ + c) * N + i] = z[c];
        if (joint_update(x, P, z, Rn, K, &ptr, &kg, &nis)) status |= 1u;
        for (int c = 0; c < NX; ++c) if (!isfinite(x[c])) status |= 2u;
        if (a->x_steps) for (int c = 0; c < NX; ++c) a->x_steps[(t * NX + c) * N + i] = x[c];
        if (a->p_trace_steps) a->p_trace_steps[t * N + i] = ptr;
        if (a->k_gain_steps) a->k_gain_steps[t * N + i] = kg;
        if (a->nis_steps) a->nis_steps[t * N + i] = nis;
        if (a->P_ckpt && a->ckpt_every > 0 && (t + 1) % a->ckpt_every == 0) {
            const int64_t k = (t + 1) / a->ckpt_every - 1;
            for (int c = 0; c < NX * NX; ++c) a->P_ckpt[(k * NX * NX + c) * N + i] = P[c];
        }
    }
    if (a->x_final) for (int c = 0; c < NX; ++c) a->x_final[c * N + i] = x[c];
    if (a->P_final) for (int c = 0; c < NX * NX; ++c) a->P_final[c * N + i] = P[c];
    if (a->K_final) for (int c = 0; c < NX * NZ; ++c) a->K_final[c * N + i] = K[c];
    if (a->status) a->status[i] = status;
}

/* trajectories are independent: a static block partition over POSIX threads (libgomp is not in the image) */
typedef struct { const KfOracleArgs *a; int64_t begin, end; } Span;

static void *span_main(void *arg) {
    const Span *sp = (const Span *)arg;
    for (int64_t i = sp->begin; i < sp->end; ++i) run_one(sp->a, i);
    return NULL;
}

int kf_oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int kf_oracle_run(const KfOracleArgs *a) {
    if (!a || a->n_traj < 0 || a->n_steps < 0 || a->n_streams <= 0) return -1;
    int nt = a->n_threads > 0 ? a->n_threads : kf_oracle_max_threads();
    if (nt > a->n_traj) nt = (int)(a->n_traj > 0 ? a->n_traj : 1);
    if (nt > 1024) nt = 1024;
    pthread_t tid[1024];
    Span span[1024];
    const int64_t chunk = (a->n_traj + nt - 1) / nt;
    for (int k = 0; k < nt; ++k) {
        span[k].a = a;
        span[k].begin = k * chunk;
        span[k].end = (k + 1) * chunk < a->n_traj ? (k + 1) * chunk : a->n_traj;
        if (k > 0 && pthread_create(&tid[k], NULL, span_main, &span[k]) != 0) return -2;
    }
    span_main(&span[0]);
    for (int k = 1; k < nt; ++k) pthread_join(tid[k], NULL);
    return 0;
}

size_t kf_oracle_args_size(void) { return sizeof(KfOracleArgs); }
