/*
 * dp_oracle.c -- CPU restatement of the density_planner hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product package (density_planner_b200/) never does.
 *
 * Parity is pinned on outputs of the unmodified reference (tests/golden/ npz files, minted by
 * tests/golden/make_golden.py); the reference itself ships no golden vectors.
 *
 * What is restated (reference file:line):
 *   - Controller.__call__                systems/sytem_CAR.py:25-37
 *   - U_FUNC.forward                     data/trained_controller/model_CAR.py:19-28
 *   - Car.a_func / b_func                systems/sytem_CAR.py:128-176
 *   - dudx_func/dfdx_func/divf_func      systems/ControlAffineSystem.py:69-122 (autograd Jacobian restated
 *                                        as forward-mode tangents; clip mask inclusive like torch.clamp)
 *   - get_next_x / get_next_rho / get_next_rholog   systems/ControlAffineSystem.py:124-157
 *   - compute_density                    systems/ControlAffineSystem.py:409-463
 *
 * Compiled twice: REAL=float (the fp32 restatement; `-ffp-contract=off`, sequential-k dot products) and
 * REAL=double (an fp64 "truth" used to judge which of two fp32 answers is closer).
 *
 * Weight blob layout (fp32, 43,592 values), per net in order {w1-net (NOUT=48), w2-net (NOUT=24)}:
 *   W1[128][4] b1[128] W2[128][128] b2[128] W3[NOUT][128] b3[NOUT]      (torch Linear.weight is [out][in])
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#ifndef SUFFIX
#define SUFFIX _f32
#endif
#define FN(name) CAT(name, SUFFIX)

typedef REAL real;

#define HID 128
#define NIN 4
#define NOUT_A 48
#define NOUT_B 24
#define NET_SIZE(nout) (HID * NIN + HID + HID * HID + HID + (nout) * HID + (nout))

static inline real r_tanh(real x) { return sizeof(real) == 4 ? (real)tanhf((float)x) : (real)tanh((double)x); }
static inline real r_sin(real x) { return sizeof(real) == 4 ? (real)sinf((float)x) : (real)sin((double)x); }
static inline real r_cos(real x) { return sizeof(real) == 4 ? (real)cosf((float)x) : (real)cos((double)x); }
static inline real r_log(real x) { return sizeof(real) == 4 ? (real)logf((float)x) : (real)log((double)x); }

/* One MLP (4 -> 128 tanh -> 128 tanh -> nout) applied to a value row and ND tangent rows.
 * in[4]: value input; din[ND][4]: input tangents.  W2/W3 are passed TRANSPOSED ([k][j]) so the k-outer loop
 * vectorises over j while keeping the sequential-k fp32 summation order. */
typedef struct {
    const float *W1, *b1, *W2t, *b2, *W3t, *b3;
    int nout;
} net_t;

#define MAXD 3

static void mlp_fwd_tan(const net_t *n, const real in[4], int nd, const real din[][4], real *out, real dout[][NOUT_A]) {
    real h[HID], dh[MAXD][HID], a2[HID], da2[MAXD][HID];
    for (int j = 0; j < HID; ++j) {
        real s = 0;
        for (int k = 0; k < NIN; ++k) s += (real)n->W1[j * NIN + k] * in[k];
        s += (real)n->b1[j];
        real t = r_tanh(s);
        h[j] = t;
        real g = (real)1 - t * t;
        for (int d = 0; d < nd; ++d) {
            real ds = 0;
            for (int k = 0; k < NIN; ++k) ds += (real)n->W1[j * NIN + k] * din[d][k];
            dh[d][j] = g * ds;
        }
    }
    for (int j = 0; j < HID; ++j) a2[j] = 0;
    for (int d = 0; d < nd; ++d)
        for (int j = 0; j < HID; ++j) da2[d][j] = 0;
    for (int k = 0; k < HID; ++k) {
        const float *w = n->W2t + (size_t)k * HID;
        real hk = h[k];
        for (int j = 0; j < HID; ++j) a2[j] += (real)w[j] * hk;
        for (int d = 0; d < nd; ++d) {
            real dk = dh[d][k];
            for (int j = 0; j < HID; ++j) da2[d][j] += (real)w[j] * dk;
        }
    }
    for (int j = 0; j < HID; ++j) {
        real t = r_tanh(a2[j] + (real)n->b2[j]);
        real g = (real)1 - t * t;
        h[j] = t;
        for (int d = 0; d < nd; ++d) dh[d][j] = g * da2[d][j];
    }
    int no = n->nout;
    for (int j = 0; j < no; ++j) out[j] = 0;
    for (int d = 0; d < nd; ++d)
        for (int j = 0; j < no; ++j) dout[d][j] = 0;
    for (int k = 0; k < HID; ++k) {
        const float *w = n->W3t + (size_t)k * no;
        real hk = h[k];
        for (int j = 0; j < no; ++j) out[j] += (real)w[j] * hk;
        for (int d = 0; d < nd; ++d) {
            real dk = dh[d][k];
            for (int j = 0; j < no; ++j) dout[d][j] += (real)w[j] * dk;
        }
    }
    for (int j = 0; j < no; ++j) out[j] += (real)n->b3[j];
}

typedef struct {
    net_t a, b;
    float *store; /* transposed copies */
} ctrl_t;

static void ctrl_init(ctrl_t *c, const float *blob) {
    const float *p = blob;
    size_t tsz = 2 * (size_t)HID * HID + (size_t)(NOUT_A + NOUT_B) * HID;
    c->store = (float *)malloc(tsz * sizeof(float));
    float *st = c->store;
    net_t *nets[2] = {&c->a, &c->b};
    int nouts[2] = {NOUT_A, NOUT_B};
    for (int m = 0; m < 2; ++m) {
        net_t *n = nets[m];
        n->nout = nouts[m];
        n->W1 = p; p += HID * NIN;
        n->b1 = p; p += HID;
        const float *W2 = p; p += HID * HID;
        n->b2 = p; p += HID;
        const float *W3 = p; p += n->nout * HID;
        n->b3 = p; p += n->nout;
        float *W2t = st; st += HID * HID;
        float *W3t = st; st += n->nout * HID;
        for (int j = 0; j < HID; ++j)
            for (int k = 0; k < HID; ++k) W2t[k * HID + j] = W2[j * HID + k];
        for (int j = 0; j < n->nout; ++j)
            for (int k = 0; k < HID; ++k) W3t[k * n->nout + j] = W3[j * HID + k];
        n->W2t = W2t;
        n->W3t = W3t;
    }
}
static void ctrl_free(ctrl_t *c) { free(c->store); }

/* Controller value + Jacobian columns.
 *   x[5] state, xr[4] reference (first four comps), ur[2] reference input
 *   dirs: bitmask of state directions to differentiate (bit i = d/dx_i, i in 0..4)
 *   u_raw[2]; J[2][5] = d u_raw / d x (columns not requested are left 0)
 * Follows sytem_CAR.py:33-36 and model_CAR.py:24-26 op for op in the value path. */
static void controller_eval(const ctrl_t *c, const real x[5], const real xr[4], const real ur[2], unsigned dirs,
                            real u_raw[2], real J[2][5]) {
    real xe[4];
    xe[0] = x[0] - xr[0];
    xe[1] = x[1] - xr[1];
    xe[2] = x[2] - xr[2];
    xe[3] = x[3] - xr[3];
    xe[2] = xe[2] + x[4]; /* xe[:, dist_state] += x[:, 4]  (sytem_CAR.py:34) */
    real in[4];
    in[0] = x[2];
    in[1] = x[3];
    in[2] = x[2] - xe[2]; /* (x - xe)[:, 2:4]  (model_CAR.py:24) */
    in[3] = x[3] - xe[3];

    /* MLP tangent directions: theta (in0), v (in1), d (in2 gets -1).  d(in2)/dtheta = 1-1 = 0 exactly. */
    int mlp_dir[3], nd = 0;
    real din[MAXD][4];
    const int cand[3] = {2, 3, 4};
    for (int q = 0; q < 3; ++q) {
        int s = cand[q];
        if (!(dirs & (1u << s))) continue;
        for (int k = 0; k < 4; ++k) din[nd][k] = 0;
        if (s == 2) din[nd][0] = 1;
        if (s == 3) din[nd][1] = 1;
        if (s == 4) din[nd][2] = -1;
        mlp_dir[nd++] = s;
    }
    real oa[NOUT_A], doa[MAXD][NOUT_A], ob[NOUT_A], dob[MAXD][NOUT_A];
    mlp_fwd_tan(&c->a, in, nd, (const real(*)[4])din, oa, doa);
    mlp_fwd_tan(&c->b, in, nd, (const real(*)[4])din, ob, dob);

    /* u = w2 . tanh(w1 . xe) + uref ; w1 = oa reshaped (12,4), w2 = ob reshaped (2,12) */
    real tz[12], gz[12];
    for (int i = 0; i < 12; ++i) {
        real s = 0;
        for (int j = 0; j < 4; ++j) s += oa[4 * i + j] * xe[j];
        tz[i] = r_tanh(s);
        gz[i] = (real)1 - tz[i] * tz[i];
    }
    for (int r = 0; r < 2; ++r) {
        real s = 0;
        for (int i = 0; i < 12; ++i) s += ob[12 * r + i] * tz[i];
        u_raw[r] = s + ur[r];
    }
    for (int r = 0; r < 2; ++r)
        for (int s = 0; s < 5; ++s) J[r][s] = 0;
    /* directions through xe only: px (xe0), py (xe1) */
    for (int s = 0; s < 2; ++s) {
        if (!(dirs & (1u << s))) continue;
        for (int r = 0; r < 2; ++r) {
            real acc = 0;
            for (int i = 0; i < 12; ++i) acc += ob[12 * r + i] * (gz[i] * oa[4 * i + s]);
            J[r][s] = acc;
        }
    }
    for (int d = 0; d < nd; ++d) {
        int s = mlp_dir[d];
        int xecol = (s == 2) ? 2 : (s == 3) ? 3 : 2; /* d(xe)/d(theta)=e2, /dv=e3, /dd=e2 */
        for (int r = 0; r < 2; ++r) {
            real acc = 0;
            for (int i = 0; i < 12; ++i) {
                real dz = 0;
                for (int j = 0; j < 4; ++j) dz += doa[d][4 * i + j] * xe[j];
                dz += oa[4 * i + xecol];
                acc += dob[d][12 * r + i] * tz[i] + ob[12 * r + i] * (gz[i] * dz);
            }
            J[r][s] = acc;
        }
    }
}

static inline real clipu(real u) { return u < (real)-3 ? (real)-3 : (u > (real)3 ? (real)3 : u); }
static inline real clipmask(real u) { return (u >= (real)-3 && u <= (real)3) ? (real)1 : (real)0; }

/* ---- single-step API (N samples, shared xref/uref) --------------------------------------------------- */
/* outputs (any may be NULL): u_raw[N][2], u[N][2], dudx[N][2][5], f[N][5], div[N] */
void FN(dpo_step)(const float *blob, const real *x, const real *xref5, const real *uref2, long N, real *u_raw, real *u,
                  real *dudx, real *f, real *div) {
    ctrl_t c;
    ctrl_init(&c, blob);
    unsigned dirs = dudx ? 0x1Fu : (div ? 0xCu : 0u);
#pragma omp parallel for schedule(static)
    for (long s = 0; s < N; ++s) {
        real ur_[2], J[2][5];
        controller_eval(&c, x + 5 * s, xref5, uref2, dirs, ur_, J);
        real uc0 = clipu(ur_[0]), uc1 = clipu(ur_[1]);
        real m0 = clipmask(ur_[0]), m1 = clipmask(ur_[1]);
        if (u_raw) { u_raw[2 * s] = ur_[0]; u_raw[2 * s + 1] = ur_[1]; }
        if (u) { u[2 * s] = uc0; u[2 * s + 1] = uc1; }
        if (dudx)
            for (int k = 0; k < 5; ++k) {
                dudx[10 * s + k] = m0 * J[0][k];
                dudx[10 * s + 5 + k] = m1 * J[1][k];
            }
        if (f) {
            const real *xs = x + 5 * s;
            f[5 * s + 0] = xs[3] * r_cos(xs[2]);
            f[5 * s + 1] = xs[3] * r_sin(xs[2]);
            f[5 * s + 2] = uc0;
            f[5 * s + 3] = uc1;
            f[5 * s + 4] = 0;
        }
        if (div) div[s] = m0 * J[0][2] + m1 * J[1][3];
    }
    ctrl_free(&c);
}

/* ---- rollout (compute_density, ControlAffineSystem.py:409-463) ---------------------------------------- */
#define DPO_LOG 1u      /* log_density=True  : l += log(1 - div*dt)   (:147-157) */
#define DPO_DENSITY 2u  /* compute_density=True                          */
#define DPO_CUT 4u      /* cutting=True : clamp >1e30 (and <0 for linear), NaN -> 1e30 (:451-462) */

/* xe0[N][5]; xref[5][T+1]; uref[2][T]; rho0[N] or NULL.
 * Outputs are written every `stride` steps (t = 0, stride, 2*stride, ... <= T), Ts = T/stride + 1 slots:
 *   xe_out[N][5][Ts]  (x - xref, the reference's return layout)     rho_out[N][Ts]
 *   clip_margin[N] = min over (t,i) of | |u_raw_i| - 3 |  (for the clip-flip audit) or NULL
 *   cut_flags[2] (or NULL): [0]=1 if any value was clamped, [1]=1 if any NaN was replaced */
void FN(dpo_rollout)(const float *blob, const real *xe0, const real *xref, const real *uref, const real *rho0, real dt,
                     long N, int T, unsigned flags, int stride, real *xe_out, real *rho_out, real *clip_margin,
                     int *cut_flags) {
    ctrl_t c;
    ctrl_init(&c, blob);
    const int Ts = T / stride + 1;
    const int logd = (flags & DPO_LOG) != 0, dens = (flags & DPO_DENSITY) != 0, cut = (flags & DPO_CUT) != 0;
    int any_clamp = 0, any_nan = 0;
#pragma omp parallel for schedule(static) reduction(| : any_clamp, any_nan)
    for (long s = 0; s < N; ++s) {
        real x[5];
        for (int k = 0; k < 5; ++k) x[k] = xe0[5 * s + k] + xref[(size_t)k * (T + 1)];
        real rho = rho0 ? rho0[s] : (logd ? (real)0 : (real)1);
        real margin = (real)1e30;
        for (int i = 0; i <= T; ++i) {
            if (i % stride == 0) {
                int o = i / stride;
                if (xe_out)
                    for (int k = 0; k < 5; ++k) xe_out[((size_t)s * 5 + k) * Ts + o] = x[k] - xref[(size_t)k * (T + 1) + i];
                if (rho_out && dens) {
                    real r = rho;
                    if (cut) {
                        if (r > (real)1e30) { r = (real)1e30; any_clamp = 1; }
                        if (!logd && r < 0) { r = 0; any_clamp = 1; }
                        if (r != r) { r = (real)1e30; any_nan = 1; }
                    }
                    rho_out[(size_t)s * Ts + o] = r;
                }
            }
            if (i == T) break;
            real xr[4] = {xref[i], xref[(size_t)(T + 1) + i], xref[(size_t)2 * (T + 1) + i], xref[(size_t)3 * (T + 1) + i]};
            real ur[2] = {uref[i], uref[(size_t)T + i]};
            real u_raw[2], J[2][5];
            controller_eval(&c, x, xr, ur, dens ? 0xCu : 0u, u_raw, J);
            real a0 = (real)fabs((double)u_raw[0]) - 3, a1 = (real)fabs((double)u_raw[1]) - 3;
            a0 = a0 < 0 ? -a0 : a0;
            a1 = a1 < 0 ? -a1 : a1;
            if (a0 < margin) margin = a0;
            if (a1 < margin) margin = a1;
            if (dens) {
                real div = clipmask(u_raw[0]) * J[0][2] + clipmask(u_raw[1]) * J[1][3];
                if (logd)
                    rho = rho + r_log((real)1 - div * dt);
                else
                    rho = rho + (-div * rho) * dt;
            }
            real f0 = x[3] * r_cos(x[2]), f1 = x[3] * r_sin(x[2]);
            x[0] = x[0] + f0 * dt;
            x[1] = x[1] + f1 * dt;
            x[2] = x[2] + clipu(u_raw[0]) * dt;
            x[3] = x[3] + clipu(u_raw[1]) * dt;
            x[4] = x[4] + (real)0 * dt;
        }
        if (clip_margin) clip_margin[s] = margin;
    }
    if (cut_flags) { cut_flags[0] = any_clamp; cut_flags[1] = any_nan; }
    ctrl_free(&c);
}

int FN(dpo_sizeof_real)(void) { return (int)sizeof(real); }

#ifdef _OPENMP
#include <omp.h>
/* number of OpenMP threads the rollout loops use (torchrun exports OMP_NUM_THREADS=1; the CPU arm overrides it) */
int FN(dpo_set_threads)(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#endif
