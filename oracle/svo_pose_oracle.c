/*
 * svo_pose_oracle.c — CPU restatement of the pose stage that follows the matchers in
 * Tracking::Tracklastframe (src/Tracking.cc:108-121).  TEST INFRASTRUCTURE ONLY (same rules as
 * svo_oracle.c: only tests/, smoke() and bench.py's CPU legs may load it).
 *
 * (1) svo_o_pose_optimize — Optimizer::PoseOptimization (src/Optimizer.cc:15-86): one
 *     VertexSE3Expmap, one EdgeSE3ProjectXYZOnlyPose per matched map point (identity information,
 *     Huber kernel with delta = sqrt(5.991)), g2o Levenberg-Marquardt with a dense 6x6 solver,
 *     optimize(10).  g2o IS vendored in the reference, so every step cites it (paths relative to
 *     Thirdparty/g2o/g2o):
 *       outer loop                core/sparse_optimizer.cpp:354-419
 *       LM step, lambda schedule  core/optimization_algorithm_levenberg.cpp:62-190
 *       robust chi2               core/sparse_optimizer.cpp:100-114
 *       quadratic form            core/base_unary_edge.hpp:43-72, core/base_edge.h:96-102
 *       Huber                     core/robust_kernel_impl.cpp:78-92
 *       error / Jacobian          types/types_six_dof_expmap.h:153-157, types_six_dof_expmap.cpp:266-296
 *       oplus, exp, product       types/types_six_dof_expmap.h:73-76, types/se3quat.h:100-113,217-249,274-279
 *       dense solve               solvers/linear_solver_dense.h:63-115 (Eigen::LDLT)
 *       cv::Mat <-> SE3Quat       src/convert.cc:6-17,49-63
 *     Eigen is not installed here, so Eigen's pieces are restated from its published algorithms:
 *     Quaterniond(Matrix3d), quaternion * vector, quaternion product, toRotationMatrix, LDLT.
 *     PARITY: PINNED to the reference's own code — oracle/_ref/libsvo_ref_g2o.so is src/Optimizer.cc,
 *     src/convert.cc and the vendored g2o compiled UNMODIFIED (oracle/Makefile `ref_g2o`, against the
 *     stand-in Eigen header oracle/ref_stubs_g2o/minieigen.hpp), and tests/test_ref_pin_pose.py finds
 *     the float32 pose Optimizer::PoseOptimization stores equal to this file's bit for bit (6-500
 *     points, outliers, perturbed starts); tests/golden/ref_pose.npz records those answers.  The test
 *     suite also checks closed-form properties (exact data -> exact pose, chi2 never increases, Huber
 *     limits) and the GPU kernel against this file within 1e-6.
 *
 * (2) svo_o_pnp_ransac — the role of cv::solvePnPRansac(pts3d, pts2d, K, noDist, rvec, tvec, false,
 *     100, 8.0, 0.99, inliers) at src/pnpmatch.cc:227.  OpenCV is un-vendored and its RANSAC
 *     (EPnP on 5-point samples, cv::RNG, adaptive iteration count, iterative refit) is NOT restated:
 *     this file DEFINES a data-parallel variant — P3P on 3-point samples drawn by a counter-based
 *     hash, every one of the `iterations` samples scored against all points with the same squared
 *     reprojection threshold, first maximum wins, Gauss-Newton refit on the winner's inliers.
 *     PARITY UNPINNED vs OpenCV by construction; tests compare it with cv2.solvePnPRansac
 *     statistically (pose within tolerance, inlier sets overlapping) and the GPU kernel with this
 *     file (same samples, same inlier mask, pose within 1e-6).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <string.h>
#include "svo_oracle.h"

/* ------------------------------------------------------------------ SE3Quat (types/se3quat.h) */
typedef struct { double x, y, z, w; double t[3]; } se3q;

static void q_normalize_pos(se3q *s)   /* se3quat.h:274-279 */
{
    if (s->w < 0) { s->x = -s->x; s->y = -s->y; s->z = -s->z; s->w = -s->w; }
    double n2 = s->x * s->x + s->y * s->y + s->z * s->z + s->w * s->w;
    if (n2 > 0) { double n = sqrt(n2); s->x /= n; s->y /= n; s->z /= n; s->w /= n; }
}

static void q_from_R(const double R[9], se3q *s)   /* Eigen Quaternion(Matrix3) */
{
    double q[4];
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
    s->x = q[0]; s->y = q[1]; s->z = q[2]; s->w = q[3];
}

static void q_rotate(const se3q *s, const double v[3], double o[3])   /* Eigen _transformVector */
{
    double ux = s->y * v[2] - s->z * v[1], uy = s->z * v[0] - s->x * v[2], uz = s->x * v[1] - s->y * v[0];
    ux += ux; uy += uy; uz += uz;
    o[0] = v[0] + s->w * ux + (s->y * uz - s->z * uy);
    o[1] = v[1] + s->w * uy + (s->z * ux - s->x * uz);
    o[2] = v[2] + s->w * uz + (s->x * uy - s->y * ux);
}

static void q_to_R(const se3q *s, double R[9])   /* Eigen toRotationMatrix */
{
    double tx = 2 * s->x, ty = 2 * s->y, tz = 2 * s->z;
    double twx = tx * s->w, twy = ty * s->w, twz = tz * s->w;
    double txx = tx * s->x, txy = ty * s->x, txz = tz * s->x;
    double tyy = ty * s->y, tyz = tz * s->y, tzz = tz * s->z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

static se3q se3_mul(const se3q *a, const se3q *b)   /* se3quat.h:100-106 */
{
    se3q r; double rt[3];
    q_rotate(a, b->t, rt);
    r.t[0] = a->t[0] + rt[0]; r.t[1] = a->t[1] + rt[1]; r.t[2] = a->t[2] + rt[2];
    r.w = a->w * b->w - a->x * b->x - a->y * b->y - a->z * b->z;
    r.x = a->w * b->x + a->x * b->w + a->y * b->z - a->z * b->y;
    r.y = a->w * b->y + a->y * b->w + a->z * b->x - a->x * b->z;
    r.z = a->w * b->z + a->z * b->w + a->x * b->y - a->y * b->x;
    q_normalize_pos(&r);
    return r;
}

static void mat3_mul(const double A[9], const double B[9], double C[9])
{
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

static se3q se3_exp(const double u[6])   /* se3quat.h:217-249: omega = u[0..2], upsilon = u[3..5] */
{
    const double *om = u, *up = u + 3;
    double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0}, O2[9], R[9], V[9];
    mat3_mul(O, O, O2);
    if (theta < 0.00001) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i];
        memcpy(V, R, sizeof V);
    } else {
        double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3);
        for (int i = 0; i < 9; ++i) {
            R[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
            V[i] = (i % 4 == 0 ? 1.0 : 0.0) + b * O[i] + c * O2[i];
        }
    }
    se3q s;
    q_from_R(R, &s);
    for (int i = 0; i < 3; ++i) s.t[i] = V[i * 3] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
    q_normalize_pos(&s);
    return s;
}

static se3q se3_from_T32(const float T[16])   /* src/convert.cc:6-17 */
{
    double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    se3q s; q_from_R(R, &s);
    s.t[0] = T[3]; s.t[1] = T[7]; s.t[2] = T[11];
    q_normalize_pos(&s);
    return s;
}

static void se3_to_T32(const se3q *s, float T[16])   /* src/convert.cc:49-63, se3quat.h:263-271 */
{
    double R[9]; q_to_R(s, R);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = (float)R[i * 3 + j]; T[i * 4 + 3] = (float)s->t[i]; }
    T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}

/* Eigen::LDLT (pivoted, lower) on a 6x6 system; returns 0 when a negative pivot shows up
 * (isPositive() false, linear_solver_dense.h:107).  Zero pivots solve to zero like Eigen. */
static int ldlt6_solve(const double Hin[36], const double b[6], double x[6])
{
    double A[36]; int perm[6]; memcpy(A, Hin, sizeof A);
    const int n = 6; int positive = 1;
    for (int k = 0; k < n; ++k) {
        int p = k; double big = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) if (fabs(A[i * n + i]) > big) { big = fabs(A[i * n + i]); p = i; }
        perm[k] = p;
        if (p != k) {   /* symmetric swap of rows/cols k and p (full storage) */
            for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
            for (int i = 0; i < n; ++i) { double t = A[i * n + k]; A[i * n + k] = A[i * n + p]; A[i * n + p] = t; }
        }
        /* A[k][k] -= sum_j L[k][j]^2 D[j]; column below */
        for (int j = 0; j < k; ++j) A[k * n + k] -= A[k * n + j] * A[k * n + j] * A[j * n + j];
        double d = A[k * n + k];
        if (d < 0) positive = 0;
        for (int i = k + 1; i < n; ++i) {
            double s = A[i * n + k];
            for (int j = 0; j < k; ++j) s -= A[i * n + j] * A[k * n + j] * A[j * n + j];
            A[i * n + k] = d != 0 ? s / d : 0.0;
        }
    }
    if (!positive) return 0;
    double y[6];
    for (int i = 0; i < n; ++i) y[i] = b[i];
    for (int k = 0; k < n; ++k) if (perm[k] != k) { double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i * n + j] * y[j];
    for (int i = 0; i < n; ++i) y[i] = A[i * n + i] != 0 ? y[i] / A[i * n + i] : 0.0;
    for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) y[i] -= A[j * n + i] * y[j];
    for (int k = n - 1; k >= 0; --k) if (perm[k] != k) { double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
    for (int i = 0; i < n; ++i) x[i] = y[i];
    return 1;
}

typedef struct { const float *Xw, *obs; int n; double fx, fy, cx, cy; double delta; } pose_problem;

static void huber(double e, double delta, double rho[3])   /* robust_kernel_impl.cpp:78-92; delta<=0: no kernel */
{
    double dsqr = delta * delta;
    if (delta <= 0 || e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
    else { double sq = sqrt(e); rho[0] = 2 * sq * delta - dsqr; rho[1] = delta / sq; rho[2] = -0.5 * rho[1] / e; }
}

static void edge_error(const pose_problem *P, const se3q *s, int i, double e[2], double xyz[3])
{
    double X[3] = {P->Xw[3 * i], P->Xw[3 * i + 1], P->Xw[3 * i + 2]}, r[3];
    q_rotate(s, X, r);                                   /* se3quat.h:211-214 map() */
    xyz[0] = r[0] + s->t[0]; xyz[1] = r[1] + s->t[1]; xyz[2] = r[2] + s->t[2];
    double px = xyz[0] / xyz[2], py = xyz[1] / xyz[2];   /* project2d */
    e[0] = (double)P->obs[2 * i] - (px * P->fx + P->cx); /* types_six_dof_expmap.h:153-157 */
    e[1] = (double)P->obs[2 * i + 1] - (py * P->fy + P->cy);
}

static double robust_chi2(const pose_problem *P, const se3q *s)   /* sparse_optimizer.cpp:100-114 */
{
    double chi = 0, rho[3], e[2], xyz[3];
    for (int i = 0; i < P->n; ++i) { edge_error(P, s, i, e, xyz); huber(e[0] * e[0] + e[1] * e[1], P->delta, rho); chi += rho[0]; }
    return chi;
}

static void build_system(const pose_problem *P, const se3q *s, double H[36], double b[6])
{
    memset(H, 0, 36 * sizeof(double)); memset(b, 0, 6 * sizeof(double));
    for (int i = 0; i < P->n; ++i) {
        double e[2], xyz[3], rho[3], J[12];
        edge_error(P, s, i, e, xyz);
        double x = xyz[0], y = xyz[1], invz = 1.0 / xyz[2], invz_2 = invz * invz;   /* types_six_dof_expmap.cpp:266-288 */
        J[0] = x * y * invz_2 * P->fx; J[1] = -(1 + (x * x * invz_2)) * P->fx; J[2] = y * invz * P->fx;
        J[3] = -invz * P->fx; J[4] = 0; J[5] = x * invz_2 * P->fx;
        J[6] = (1 + y * y * invz_2) * P->fy; J[7] = -x * y * invz_2 * P->fy; J[8] = -x * invz * P->fy;
        J[9] = 0; J[10] = -invz * P->fy; J[11] = y * invz_2 * P->fy;
        huber(e[0] * e[0] + e[1] * e[1], P->delta, rho);
        for (int r = 0; r < 6; ++r) {                    /* base_unary_edge.hpp:56-63 */
            b[r] -= rho[1] * (J[r] * e[0] + J[6 + r] * e[1]);
            for (int c = 0; c < 6; ++c) H[r * 6 + c] += rho[1] * (J[r] * J[c] + J[6 + r] * J[6 + c]);
        }
    }
}

/* One LM step (optimization_algorithm_levenberg.cpp:62-165).  Returns 1 = OK, 0 = Terminate. */
typedef struct { double lambda, ni; int nbad; } lm_state;
static int lm_solve(const pose_problem *P, se3q *est, int iteration, lm_state *L, int max_trials)
{
    double H[36], b[6], x[6];
    double currentChi = robust_chi2(P, est), tempChi = currentChi, iniChi = currentChi;
    build_system(P, est, H, b);
    if (iteration == 0) {
        double md = 0; for (int j = 0; j < 6; ++j) md = fmax(fabs(H[j * 7]), md);
        L->lambda = 1e-5 * md; L->ni = 2; L->nbad = 0;
    }
    double rho = 0; int qmax = 0;
    do {
        se3q backup = *est;                              /* push */
        double Hl[36]; memcpy(Hl, H, sizeof Hl);
        for (int j = 0; j < 6; ++j) Hl[j * 7] += L->lambda;
        int ok2 = ldlt6_solve(Hl, b, x);
        if (!ok2) memset(x, 0, sizeof x);
        se3q d = se3_exp(x);                             /* oplus: exp(update) * estimate */
        *est = se3_mul(&d, est);
        tempChi = robust_chi2(P, est);
        if (!ok2) tempChi = DBL_MAX;
        rho = currentChi - tempChi;
        double scale = 0; for (int j = 0; j < 6; ++j) scale += x[j] * (L->lambda * x[j] + b[j]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && isfinite(tempChi)) {
            double alpha = 1. - pow(2 * rho - 1, 3);
            alpha = fmin(alpha, 2. / 3.);
            double sf = fmax(1. / 3., alpha);
            L->lambda *= sf; L->ni = 2; currentChi = tempChi;
        } else {
            L->lambda *= L->ni; L->ni *= 2; *est = backup;   /* pop */
        }
        ++qmax;
    } while (rho < 0 && qmax < max_trials);
    if (qmax == max_trials || rho == 0) return 0;
    if ((iniChi - currentChi) * 1e3 < iniChi) L->nbad++; else L->nbad = 0;
    if (L->nbad >= 3) return 0;
    return 1;
}

/* Optimizer::PoseOptimization (src/Optimizer.cc:15-86).  Xw: n x 3 world points of the matched map
 * points, obs: n x 2 keypoints_l[i].pt, Tcw_*: 4x4 row-major float (cv::Mat CV_32F).  Returns n
 * (nInitialCorrespondences); stats (may be NULL): [0] outer iterations run, [1] final robust chi2. */
int svo_o_pose_optimize(const float *Xw, const float *obs, int n, float fx, float fy, float cx, float cy,
                        const float *Tcw_in, float *Tcw_out, int iterations, double *stats)
{
    pose_problem P = {Xw, obs, n, fx, fy, cx, cy, (double)(float)sqrt(5.991)};   /* deltaMono is a float (src/Optimizer.cc:36) widened by setDelta */
    se3q est = se3_from_T32(Tcw_in);
    lm_state L = {-1., 2., 0};
    int it = 0, ok = 1;
    for (; it < iterations && ok; ++it) ok = lm_solve(&P, &est, it, &L, 10);   /* sparse_optimizer.cpp:376-414 */
    se3_to_T32(&est, Tcw_out);
    if (stats) { stats[0] = it; stats[1] = robust_chi2(&P, &est); }
    return n;
}

/* ------------------------------------------------------------------ P3P RANSAC (defined here) */
static uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

/* sample 3 distinct indices for iteration `it`; returns 0 when it cannot (n < 3) */
static int draw3(uint32_t seed, int it, int n, int s[3])
{
    if (n < 3) return 0;
    int c = 0;
    for (int k = 0; k < 3; ++k) {
        for (;; ++c) {
            if (c >= 64) return 0;
            int v = (int)(mix32(seed + 0x9E3779B9u * (uint32_t)(it * 64 + c + 1)) % (uint32_t)n);
            int dup = 0; for (int j = 0; j < k; ++j) dup |= s[j] == v;
            if (!dup) { s[k] = v; ++c; break; }
        }
    }
    return 1;
}

static double poly_eval(const double *c, int deg, double x) { double r = c[deg]; for (int i = deg - 1; i >= 0; --i) r = r * x + c[i]; return r; }

/* real roots of c[0] + c[1] x + ... + c[4] x^4 (c[4] != 0), Ferrari + two Newton steps each */
static int quartic_roots(const double c[5], double r[4])
{
    double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
    double a2 = a * a;
    double p = b - 3 * a2 / 8, q = cc - a * b / 2 + a2 * a / 8, rr = d - a * cc / 4 + a2 * b / 16 - 3 * a2 * a2 / 256;
    double y[4]; int n = 0;
    if (fabs(q) < 1e-14 * (1 + fabs(p) + fabs(rr))) {          /* biquadratic */
        double disc = p * p - 4 * rr;
        if (disc >= 0) {
            double sd = sqrt(disc), z1 = (-p + sd) / 2, z2 = (-p - sd) / 2;
            if (z1 >= 0) { y[n++] = sqrt(z1); y[n++] = -sqrt(z1); }
            if (z2 >= 0) { y[n++] = sqrt(z2); y[n++] = -sqrt(z2); }
        }
    } else {
        /* largest real root z of z^3 + 2p z^2 + (p^2 - 4r) z - q^2 = 0 (positive because f(0) = -q^2 < 0) */
        double A = 2 * p, B = p * p - 4 * rr, C = -q * q;
        double Q = (A * A - 3 * B) / 9, R = (2 * A * A * A - 9 * A * B + 27 * C) / 54, z;
        if (R * R < Q * Q * Q) {
            double th = acos(R / sqrt(Q * Q * Q)), sq = -2 * sqrt(Q);
            double z0 = sq * cos(th / 3) - A / 3, z1 = sq * cos((th + 2 * M_PI) / 3) - A / 3, z2 = sq * cos((th - 2 * M_PI) / 3) - A / 3;
            z = fmax(z0, fmax(z1, z2));
        } else {
            double Aa = -copysign(cbrt(fabs(R) + sqrt(R * R - Q * Q * Q)), R);
            double Bb = Aa != 0 ? Q / Aa : 0;
            z = Aa + Bb - A / 3;
        }
        for (int k = 0; k < 3; ++k) {                           /* polish the cubic root */
            double f = ((z + A) * z + B) * z + C, df = (3 * z + 2 * A) * z + B;
            if (df != 0) z -= f / df;
        }
        if (z <= 0) return 0;
        double s = sqrt(z), t1 = (p + z - q / s) / 2, t2 = (p + z + q / s) / 2;
        double d1 = z - 4 * t1, d2 = z - 4 * t2;                /* y^2 + s y + t1, y^2 - s y + t2 */
        if (d1 >= 0) { double sd = sqrt(d1); y[n++] = (-s + sd) / 2; y[n++] = (-s - sd) / 2; }
        if (d2 >= 0) { double sd = sqrt(d2); y[n++] = (s + sd) / 2; y[n++] = (s - sd) / 2; }
    }
    double dc[4] = {c[1], 2 * c[2], 3 * c[3], 4 * c[4]};
    for (int i = 0; i < n; ++i) {
        double x = y[i] - a / 4;
        for (int k = 0; k < 2; ++k) { double f = poly_eval(c, 4, x), df = poly_eval(dc, 3, x); if (df != 0) x -= f / df; }
        r[i] = x;
    }
    return n;
}

static void cross3(const double a[3], const double b[3], double o[3]) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static int unit3(double a[3]) { double n = sqrt(dot3(a, a)); if (!(n > 1e-12)) return 0; a[0] /= n; a[1] /= n; a[2] /= n; return 1; }

static int frame3(const double P1[3], const double P2[3], const double P3[3], double E[9])   /* rows: e1, e2, e3 */
{
    double d1[3] = {P2[0] - P1[0], P2[1] - P1[1], P2[2] - P1[2]}, d2[3] = {P3[0] - P1[0], P3[1] - P1[1], P3[2] - P1[2]};
    if (!unit3(d1)) return 0;
    double e3[3]; cross3(d1, d2, e3);
    if (!unit3(e3)) return 0;
    double e2[3]; cross3(e3, d1, e2);
    memcpy(E, d1, 24); memcpy(E + 3, e2, 24); memcpy(E + 6, e3, 24);
    return 1;
}

/* P3P by elimination to a quartic in v = s3/s1 (Fischler-Bolles / Grunert form).  X: three world
 * points, yb: three unit bearings.  Writes up to 4 poses (R row-major 9 + t 3); returns the count. */
static int p3p(const double X[3][3], const double yb[3][3], double out[4][12])
{
    double d23[3] = {X[1][0] - X[2][0], X[1][1] - X[2][1], X[1][2] - X[2][2]};
    double d13[3] = {X[0][0] - X[2][0], X[0][1] - X[2][1], X[0][2] - X[2][2]};
    double d12[3] = {X[0][0] - X[1][0], X[0][1] - X[1][1], X[0][2] - X[1][2]};
    double a2 = dot3(d23, d23), b2 = dot3(d13, d13), c2 = dot3(d12, d12);
    if (!(a2 > 1e-12 && b2 > 1e-12 && c2 > 1e-12)) return 0;
    double Ew[9];
    if (!frame3(X[0], X[1], X[2], Ew)) return 0;
    double ca = dot3(yb[1], yb[2]), cb = dot3(yb[0], yb[2]), cg = dot3(yb[0], yb[1]);
    double k = (a2 - c2) / b2, m = c2 / b2;
    double P[3] = {1 + k, -2 * k * cb, k - 1};      /* u = P(v) / Q(v) */
    double Q[2] = {2 * cg, -2 * ca};
    double q[3] = {1, -2 * cb, 1};                  /* s1^2 q(v) = b^2 */
    /* P^2 - 2 cg P Q + (1 - m q) Q^2 = 0 */
    double c[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c[i + j] += P[i] * P[j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) c[i + j] -= 2 * cg * P[i] * Q[j];
    double Q2[3] = {Q[0] * Q[0], 2 * Q[0] * Q[1], Q[1] * Q[1]};
    double w[3] = {1 - m * q[0], -m * q[1], -m * q[2]};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c[i + j] += w[i] * Q2[j];
    double cmax = 0; for (int i = 0; i < 5; ++i) cmax = fmax(cmax, fabs(c[i]));
    if (!(fabs(c[4]) > 1e-12 * cmax)) return 0;
    double roots[4]; int nr = quartic_roots(c, roots), ns = 0;
    for (int i = 0; i < nr; ++i) {
        double v = roots[i];
        if (!(v > 0) || !isfinite(v)) continue;
        double Qv = Q[0] + Q[1] * v;
        if (fabs(Qv) < 1e-9) continue;
        double u = (P[0] + (P[1] + P[2] * v) * v) / Qv;
        if (!(u > 0)) continue;
        double qv = q[0] + (q[1] + q[2] * v) * v;
        if (!(qv > 0)) continue;
        double s1 = sqrt(b2 / qv), s2 = u * s1, s3 = v * s1;
        double C1[3] = {s1 * yb[0][0], s1 * yb[0][1], s1 * yb[0][2]};
        double C2[3] = {s2 * yb[1][0], s2 * yb[1][1], s2 * yb[1][2]};
        double C3[3] = {s3 * yb[2][0], s3 * yb[2][1], s3 * yb[2][2]};
        double Ec[9];
        if (!frame3(C1, C2, C3, Ec)) continue;
        double *R = out[ns], *t = out[ns] + 9;
        for (int r = 0; r < 3; ++r) for (int cidx = 0; cidx < 3; ++cidx)
            R[r * 3 + cidx] = Ec[r] * Ew[cidx] + Ec[3 + r] * Ew[3 + cidx] + Ec[6 + r] * Ew[6 + cidx];
        for (int r = 0; r < 3; ++r) t[r] = C1[r] - (R[r * 3] * X[0][0] + R[r * 3 + 1] * X[0][1] + R[r * 3 + 2] * X[0][2]);
        ++ns;
    }
    return ns;
}

/* float32 scoring, one op per line order shared with the CUDA kernel (no FMA on either side) */
static int is_inlier_f32(const float R[9], const float t[3], const float *X, const float *o,
                         float fx, float fy, float cx, float cy, float thr2)
{
    float xc = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    float yc = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    float zc = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    if (!(zc > 0.f)) return 0;
    float iz = 1.f / zc;
    float du = fx * (xc * iz) + cx - o[0], dv = fy * (yc * iz) + cy - o[1];
    return du * du + dv * dv <= thr2;
}

/* Data-parallel stand-in for cv::solvePnPRansac at src/pnpmatch.cc:227 (see the file header).
 * pts3d n x 3, pts2d n x 2 (float like cv::Point3f/Point2f).  Writes R (3x3 row-major double = cv::Rodrigues
 * of the reference's rvec, src/pnpmatch.cc:237-238), t (3), inlier mask (n); returns the inlier count of the
 * winning hypothesis (0: no model).  info (may be NULL): [0] winning iteration, [1] its solution index,
 * [2] hypotheses scored. */
int svo_o_pnp_ransac(const float *pts3d, const float *pts2d, int n, float fx, float fy, float cx, float cy,
                     int iterations, float reproj_err, uint32_t seed, int refine_iters,
                     double *R_out, double *t_out, uint8_t *inlier, int32_t *info)
{
    int best_cnt = 0, best_it = -1, best_sol = -1, scored = 0; double best[12];
    const float thr2 = reproj_err * reproj_err;
    for (int it = 0; it < iterations; ++it) {
        int s[3];
        if (!draw3(seed, it, n, s)) continue;
        double X[3][3], yb[3][3], hyp[4][12];
        for (int k = 0; k < 3; ++k) {
            for (int j = 0; j < 3; ++j) X[k][j] = pts3d[3 * s[k] + j];
            yb[k][0] = ((double)pts2d[2 * s[k]] - cx) / fx; yb[k][1] = ((double)pts2d[2 * s[k] + 1] - cy) / fy; yb[k][2] = 1;
            unit3(yb[k]);
        }
        int ns = p3p(X, yb, hyp);
        for (int h = 0; h < ns; ++h) {
            float Rf[9], tf[3]; int ok = 1;
            for (int j = 0; j < 9; ++j) { Rf[j] = (float)hyp[h][j]; ok &= isfinite(Rf[j]); }
            for (int j = 0; j < 3; ++j) { tf[j] = (float)hyp[h][9 + j]; ok &= isfinite(tf[j]); }
            if (!ok) continue;
            int cnt = 0;
            for (int i = 0; i < n; ++i) cnt += is_inlier_f32(Rf, tf, pts3d + 3 * i, pts2d + 2 * i, fx, fy, cx, cy, thr2);
            ++scored;
            if (cnt > best_cnt) { best_cnt = cnt; best_it = it; best_sol = h; memcpy(best, hyp[h], sizeof best); }
        }
    }
    if (info) { info[0] = best_it; info[1] = best_sol; info[2] = scored; }
    if (best_cnt < 3) { if (inlier) memset(inlier, 0, (size_t)n); return 0; }
    float Rf[9], tf[3];
    for (int j = 0; j < 9; ++j) Rf[j] = (float)best[j];
    for (int j = 0; j < 3; ++j) tf[j] = (float)best[9 + j];
    /* refit on the inliers: Gauss-Newton on the squared reprojection error with the same left-multiplied
     * exp update as the LM above (lambda = 0, no kernel), until the step is tiny */
    float *Xi = 0, *oi = 0; int m = 0;
    Xi = (float *)__builtin_malloc(sizeof(float) * 3 * (size_t)n); oi = (float *)__builtin_malloc(sizeof(float) * 2 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        int in = is_inlier_f32(Rf, tf, pts3d + 3 * i, pts2d + 2 * i, fx, fy, cx, cy, thr2);
        if (inlier) inlier[i] = (uint8_t)in;
        if (in) { memcpy(Xi + 3 * m, pts3d + 3 * i, 12); memcpy(oi + 2 * m, pts2d + 2 * i, 8); ++m; }
    }
    se3q est; q_from_R(best, &est); est.t[0] = best[9]; est.t[1] = best[10]; est.t[2] = best[11]; q_normalize_pos(&est);
    pose_problem P = {Xi, oi, m, fx, fy, cx, cy, 0.0};
    for (int k = 0; k < refine_iters; ++k) {
        double H[36], b[6], x[6];
        build_system(&P, &est, H, b);
        if (!ldlt6_solve(H, b, x)) break;
        double nx = 0; for (int j = 0; j < 6; ++j) nx += x[j] * x[j];
        if (!isfinite(nx)) break;
        se3q d = se3_exp(x); est = se3_mul(&d, &est);
        if (nx < 1e-20) break;
    }
    __builtin_free(Xi); __builtin_free(oi);
    q_to_R(&est, R_out); t_out[0] = est.t[0]; t_out[1] = est.t[1]; t_out[2] = est.t[2];
    return best_cnt;
}
