// pose.cu — the pose stage that follows the matchers in Tracking::Tracklastframe (src/Tracking.cc:108-121):
//   k_pnp_ransac : the role of cv::solvePnPRansac at src/pnpmatch.cc:227 (100 iterations, 8 px, all points) as a
//                  data-parallel P3P RANSAC — every sample solved and scored at once, first maximum wins,
//                  Gauss-Newton refit on the winner's inliers.  OpenCV's sampler/solver is not reproduced
//                  (opt-in, non-parity); the definition is oracle/svo_pose_oracle.c.
//   k_pose_lm    : Optimizer::PoseOptimization (src/Optimizer.cc:15-86) — g2o Levenberg-Marquardt on one
//                  VertexSE3Expmap with EdgeSE3ProjectXYZOnlyPose edges and a Huber kernel, optimize(10).
//                  Follows Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:62-165 step by step;
//                  the per-edge sums are block reductions instead of a sequential loop (float64, so results
//                  agree with the sequential form to ~1e-12; tests allow 1e-6).
// One CTA per problem (frame): the edges of a frame are spread over the CTA's threads, the 6x6 solve and the
// SE(3) update run on thread 0.  Compute-light and latency-bound: a batch of frames fills the SMs.
#include "svo_internal.cuh"
#include <float.h>

namespace {

constexpr int PT = 256;            // threads per CTA
constexpr int PW = PT / 32;
constexpr double kPi = 3.14159265358979323846;

struct Se3 { double x, y, z, w, t[3]; };

__device__ void q_normalize_pos(Se3 &s)   // se3quat.h:274-279
{
    if (s.w < 0) { s.x = -s.x; s.y = -s.y; s.z = -s.z; s.w = -s.w; }
    const double n2 = s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;
    if (n2 > 0) { const double n = sqrt(n2); s.x /= n; s.y /= n; s.z /= n; s.w /= n; }
}

__device__ void q_from_R(const double *R, Se3 &s)   // Eigen Quaternion(Matrix3)
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
        if (R[8] > R[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
    s.x = q[0]; s.y = q[1]; s.z = q[2]; s.w = q[3];
}

__device__ void q_to_R(const Se3 &s, double *R)   // Eigen toRotationMatrix
{
    const double tx = 2 * s.x, ty = 2 * s.y, tz = 2 * s.z;
    const double twx = tx * s.w, twy = ty * s.w, twz = tz * s.w;
    const double txx = tx * s.x, txy = ty * s.x, txz = tz * s.x;
    const double tyy = ty * s.y, tyz = tz * s.y, tzz = tz * s.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

__device__ Se3 se3_mul(const Se3 &a, const Se3 &b)   // se3quat.h:100-106
{
    Se3 r;
    double ux = a.y * b.t[2] - a.z * b.t[1], uy = a.z * b.t[0] - a.x * b.t[2], uz = a.x * b.t[1] - a.y * b.t[0];
    ux += ux; uy += uy; uz += uz;
    r.t[0] = a.t[0] + (b.t[0] + a.w * ux + (a.y * uz - a.z * uy));
    r.t[1] = a.t[1] + (b.t[1] + a.w * uy + (a.z * ux - a.x * uz));
    r.t[2] = a.t[2] + (b.t[2] + a.w * uz + (a.x * uy - a.y * ux));
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    q_normalize_pos(r);
    return r;
}

__device__ Se3 se3_exp(const double *u)   // se3quat.h:217-249 (omega = u[0..2], upsilon = u[3..5])
{
    const double *om = u, *up = u + 3;
    const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double O2[9], R[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    if (theta < 0.00001) {
        for (int i = 0; i < 9; ++i) { R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / (theta * theta * theta);
        for (int i = 0; i < 9; ++i) {
            const double id = i % 4 == 0 ? 1.0 : 0.0;
            R[i] = id + a * O[i] + b * O2[i];
            V[i] = id + b * O[i] + c * O2[i];
        }
    }
    Se3 s;
    q_from_R(R, s);
    for (int i = 0; i < 3; ++i) s.t[i] = V[i * 3] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
    q_normalize_pos(s);
    return s;
}

// Eigen::LDLT (pivot = largest remaining diagonal entry, lower storage); false on a negative pivot.
__device__ bool ldlt6_solve(const double *Hin, const double *b, double *x)
{
    double A[36]; int perm[6];
    for (int i = 0; i < 36; ++i) A[i] = Hin[i];
    bool positive = true;
    for (int k = 0; k < 6; ++k) {
        int p = k; double big = fabs(A[k * 7]);
        for (int i = k + 1; i < 6; ++i) if (fabs(A[i * 7]) > big) { big = fabs(A[i * 7]); p = i; }
        perm[k] = p;
        if (p != k) {
            for (int j = 0; j < 6; ++j) { const double t = A[k * 6 + j]; A[k * 6 + j] = A[p * 6 + j]; A[p * 6 + j] = t; }
            for (int i = 0; i < 6; ++i) { const double t = A[i * 6 + k]; A[i * 6 + k] = A[i * 6 + p]; A[i * 6 + p] = t; }
        }
        for (int j = 0; j < k; ++j) A[k * 7] -= A[k * 6 + j] * A[k * 6 + j] * A[j * 7];
        const double d = A[k * 7];
        if (d < 0) positive = false;
        for (int i = k + 1; i < 6; ++i) {
            double s = A[i * 6 + k];
            for (int j = 0; j < k; ++j) s -= A[i * 6 + j] * A[k * 6 + j] * A[j * 7];
            A[i * 6 + k] = d != 0 ? s / d : 0.0;
        }
    }
    if (!positive) return false;
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = b[i];
    for (int k = 0; k < 6; ++k) if (perm[k] != k) { const double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) y[i] -= A[i * 6 + j] * y[j];
    for (int i = 0; i < 6; ++i) y[i] = A[i * 7] != 0 ? y[i] / A[i * 7] : 0.0;
    for (int i = 5; i >= 0; --i) for (int j = i + 1; j < 6; ++j) y[i] -= A[j * 6 + i] * y[j];
    for (int k = 5; k >= 0; --k) if (perm[k] != k) { const double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
    for (int i = 0; i < 6; ++i) x[i] = y[i];
    return true;
}

// Sum NV per-thread doubles over the CTA; the totals land in tot[0..NV) (shared) for every thread to read.
template <int NV>
__device__ void block_sum(double (&v)[NV], double (*red)[28], double *tot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = v[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < PW; ++w) s += red[w][threadIdx.x];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
}

struct Cam { double fx, fy, cx, cy; };

// error of one edge (types_six_dof_expmap.h:153-157) at pose (R, t) = P[0..12); xyz out
__device__ __forceinline__ void edge_error(const double *P, const Cam &K, const float *X, const float *o, double &e0, double &e1,
                                           double &x, double &y, double &z)
{
    const double X0 = X[0], X1 = X[1], X2 = X[2];
    x = P[0] * X0 + P[1] * X1 + P[2] * X2 + P[9];
    y = P[3] * X0 + P[4] * X1 + P[5] * X2 + P[10];
    z = P[6] * X0 + P[7] * X1 + P[8] * X2 + P[11];
    e0 = (double)o[0] - ((x / z) * K.fx + K.cx);
    e1 = (double)o[1] - ((y / z) * K.fy + K.cy);
}

// Huber (robust_kernel_impl.cpp:78-92); delta <= 0: plain squared error
__device__ __forceinline__ void huber(double e, double delta, double &rho0, double &rho1)
{
    if (delta <= 0 || e <= delta * delta) { rho0 = e; rho1 = 1.; }
    else { const double sq = sqrt(e); rho0 = 2 * sq * delta - delta * delta; rho1 = delta / sq; }
}

// v[0..21) upper triangle of H (row-major), v[21..27) b, v[27] robust chi2 — base_unary_edge.hpp:56-63 over this
// thread's edges.  `w` (may be NULL) restricts the sum to edges with w[i] != 0 (the RANSAC refit).
__device__ void accumulate_system(const double *P, const Cam &K, double delta, const float *p3, const float *p2, int n,
                                  const uint8_t *w, double (&v)[28])
{
#pragma unroll
    for (int k = 0; k < 28; ++k) v[k] = 0;
    for (int i = threadIdx.x; i < n; i += PT) {
        if (w && !w[i]) continue;
        double e0, e1, x, y, z;
        edge_error(P, K, p3 + 3 * i, p2 + 2 * i, e0, e1, x, y, z);
        const double invz = 1.0 / z, invz_2 = invz * invz;    // types_six_dof_expmap.cpp:266-288
        double J0[6], J1[6];
        J0[0] = x * y * invz_2 * K.fx; J0[1] = -(1 + (x * x * invz_2)) * K.fx; J0[2] = y * invz * K.fx;
        J0[3] = -invz * K.fx; J0[4] = 0; J0[5] = x * invz_2 * K.fx;
        J1[0] = (1 + y * y * invz_2) * K.fy; J1[1] = -x * y * invz_2 * K.fy; J1[2] = -x * invz * K.fy;
        J1[3] = 0; J1[4] = -invz * K.fy; J1[5] = y * invz_2 * K.fy;
        double rho0, rho1;
        huber(e0 * e0 + e1 * e1, delta, rho0, rho1);
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int c = r; c < 6; ++c) v[k++] += rho1 * (J0[r] * J0[c] + J1[r] * J1[c]);
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) v[21 + r] -= rho1 * (J0[r] * e0 + J1[r] * e1);
        v[27] += rho0;
    }
}

__device__ double partial_chi2(const double *P, const Cam &K, double delta, const float *p3, const float *p2, int n)
{
    double chi = 0;
    for (int i = threadIdx.x; i < n; i += PT) {
        double e0, e1, x, y, z, rho0, rho1;
        edge_error(P, K, p3 + 3 * i, p2 + 2 * i, e0, e1, x, y, z);
        huber(e0 * e0 + e1 * e1, delta, rho0, rho1);
        chi += rho0;
    }
    return chi;
}

__device__ void unpack_system(const double *tot, double *H, double *b)
{
    int k = 0;
    for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { H[r * 6 + c] = tot[k]; H[c * 6 + r] = tot[k]; ++k; }
    for (int r = 0; r < 6; ++r) b[r] = tot[21 + r];
}

__device__ void publish_pose(const Se3 &s, double *P)
{
    q_to_R(s, P);
    P[9] = s.t[0]; P[10] = s.t[1]; P[11] = s.t[2];
}

// ---------------------------------------------------------------------------------------------------------------
// Optimizer::PoseOptimization.  One CTA per problem.
__global__ void __launch_bounds__(PT) k_pose_lm(PoseArgs a)
{
    __shared__ double red[PW][28];
    __shared__ double tot[28];
    __shared__ double P[12];
    __shared__ int flag;          // 0: next trial, 1: iteration done (OK), 2: terminate
    const PoseHdr &h = a.hdr[blockIdx.x];
    const int n = h.n;
    const float *p3 = a.p3 + 3 * (size_t)h.off, *p2 = a.p2 + 2 * (size_t)h.off;
    const Cam K = {(double)h.fx, (double)h.fy, (double)h.cx, (double)h.cy};
    const double delta = (double)(float)sqrt(5.991);          // deltaMono is a float (src/Optimizer.cc:36)
    Se3 est;                                                   // thread 0 only
    double lambda = -1., ni = 2.; int nbad = 0;
    if (threadIdx.x == 0) {                                    // src/convert.cc:6-17
        const double R[9] = {h.Tcw[0], h.Tcw[1], h.Tcw[2], h.Tcw[4], h.Tcw[5], h.Tcw[6], h.Tcw[8], h.Tcw[9], h.Tcw[10]};
        q_from_R(R, est);
        est.t[0] = h.Tcw[3]; est.t[1] = h.Tcw[7]; est.t[2] = h.Tcw[11];
        q_normalize_pos(est);
        publish_pose(est, P);
    }
    __syncthreads();
    double v[28];
    double currentChi = 0;
    int it = 0;
    for (; it < a.lm_iterations; ++it) {                       // sparse_optimizer.cpp:376-414
        accumulate_system(P, K, delta, p3, p2, n, nullptr, v);
        block_sum<28>(v, red, tot);
        double H[36], b[6], x[6], iniChi = 0, rho = 0; int qmax = 0;
        if (threadIdx.x == 0) {
            unpack_system(tot, H, b);
            currentChi = tot[27]; iniChi = currentChi;
            if (it == 0) {                                     // computeLambdaInit, tau = 1e-5
                double md = 0;
                for (int j = 0; j < 6; ++j) md = fmax(fabs(H[j * 7]), md);
                lambda = 1e-5 * md; ni = 2; nbad = 0;
            }
        }
        for (;;) {                                             // optimization_algorithm_levenberg.cpp:103-151
            Se3 backup; bool ok2 = true;
            if (threadIdx.x == 0) {
                backup = est;
                double Hl[36];
                for (int j = 0; j < 36; ++j) Hl[j] = H[j];
                for (int j = 0; j < 6; ++j) Hl[j * 7] += lambda;
                ok2 = ldlt6_solve(Hl, b, x);
                if (!ok2) for (int j = 0; j < 6; ++j) x[j] = 0;
                const Se3 d = se3_exp(x);
                est = se3_mul(d, est);                         // oplus: exp(update) * estimate
                publish_pose(est, P);
            }
            __syncthreads();
            double c1[1] = {partial_chi2(P, K, delta, p3, p2, n)};
            block_sum<1>(c1, red, tot);
            if (threadIdx.x == 0) {
                double tempChi = tot[0];
                if (!ok2) tempChi = DBL_MAX;
                rho = currentChi - tempChi;
                double scale = 0;
                for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + b[j]);
                scale += 1e-3;
                rho /= scale;
                if (rho > 0 && isfinite(tempChi)) {
                    double alpha = 1. - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
                    alpha = fmin(alpha, 2. / 3.);
                    const double sf = fmax(1. / 3., alpha);
                    lambda *= sf; ni = 2; currentChi = tempChi;
                } else {
                    lambda *= ni; ni *= 2; est = backup;       // pop
                    publish_pose(est, P);
                }
                ++qmax;
                if (rho < 0 && qmax < 10) flag = 0;
                else {
                    int f = 1;
                    if (qmax == 10 || rho == 0) f = 2;
                    else {
                        if ((iniChi - currentChi) * 1e3 < iniChi) ++nbad; else nbad = 0;   // "Stop criterium (Raul)"
                        if (nbad >= 3) f = 2;
                    }
                    flag = f;
                }
            }
            __syncthreads();
            if (flag != 0) break;
        }
        const int f = flag;
        __syncthreads();
        if (f == 2) { ++it; break; }
    }
    if (a.lm_iterations <= 0) {                                // optimize(0): report the cost of the input pose
        double c1[1] = {partial_chi2(P, K, delta, p3, p2, n)};
        block_sum<1>(c1, red, tot);
        currentChi = tot[0];
    }
    if (threadIdx.x == 0) {                                    // src/convert.cc:49-63
        float *T = a.Tcw_out + 16 * (size_t)blockIdx.x;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = (float)P[i * 3 + j]; T[i * 4 + 3] = (float)P[9 + i]; }
        T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
        if (a.lm_stats) { a.lm_stats[2 * blockIdx.x] = it; a.lm_stats[2 * blockIdx.x + 1] = currentChi; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// P3P RANSAC
__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__device__ bool draw3(uint32_t seed, int it, int n, int *s)
{
    if (n < 3) return false;
    int c = 0;
    for (int k = 0; k < 3; ++k) {
        for (;; ++c) {
            if (c >= 64) return false;
            const int v = (int)(mix32(seed + 0x9E3779B9u * (uint32_t)(it * 64 + c + 1)) % (uint32_t)n);
            bool dup = false;
            for (int j = 0; j < k; ++j) dup |= s[j] == v;
            if (!dup) { s[k] = v; ++c; break; }
        }
    }
    return true;
}

__device__ double poly4(const double *c, double x) { return (((c[4] * x + c[3]) * x + c[2]) * x + c[1]) * x + c[0]; }
__device__ double dpoly4(const double *c, double x) { return ((4 * c[4] * x + 3 * c[3]) * x + 2 * c[2]) * x + c[1]; }

// real roots of c0 + c1 x + .. + c4 x^4: Ferrari through the largest root of the resolvent cubic, two Newton steps each
__device__ int quartic_roots(const double *c, double *r)
{
    const double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
    const double a2 = a * a;
    const double p = b - 3 * a2 / 8, q = cc - a * b / 2 + a2 * a / 8, rr = d - a * cc / 4 + a2 * b / 16 - 3 * a2 * a2 / 256;
    double y[4]; int n = 0;
    if (fabs(q) < 1e-14 * (1 + fabs(p) + fabs(rr))) {
        const double disc = p * p - 4 * rr;
        if (disc >= 0) {
            const double sd = sqrt(disc), z1 = (-p + sd) / 2, z2 = (-p - sd) / 2;
            if (z1 >= 0) { y[n++] = sqrt(z1); y[n++] = -sqrt(z1); }
            if (z2 >= 0) { y[n++] = sqrt(z2); y[n++] = -sqrt(z2); }
        }
    } else {
        const double A = 2 * p, B = p * p - 4 * rr, C = -q * q;
        const double Q = (A * A - 3 * B) / 9, R = (2 * A * A * A - 9 * A * B + 27 * C) / 54;
        double z;
        if (R * R < Q * Q * Q) {
            const double th = acos(R / sqrt(Q * Q * Q)), sq = -2 * sqrt(Q);
            const double z0 = sq * cos(th / 3) - A / 3, z1 = sq * cos((th + 2 * kPi) / 3) - A / 3, z2 = sq * cos((th - 2 * kPi) / 3) - A / 3;
            z = fmax(z0, fmax(z1, z2));
        } else {
            const double Aa = -copysign(cbrt(fabs(R) + sqrt(R * R - Q * Q * Q)), R);
            const double Bb = Aa != 0 ? Q / Aa : 0;
            z = Aa + Bb - A / 3;
        }
        for (int k = 0; k < 3; ++k) {
            const double f = ((z + A) * z + B) * z + C, df = (3 * z + 2 * A) * z + B;
            if (df != 0) z -= f / df;
        }
        if (z <= 0) return 0;
        const double s = sqrt(z), t1 = (p + z - q / s) / 2, t2 = (p + z + q / s) / 2;
        const double d1 = z - 4 * t1, d2 = z - 4 * t2;
        if (d1 >= 0) { const double sd = sqrt(d1); y[n++] = (-s + sd) / 2; y[n++] = (-s - sd) / 2; }
        if (d2 >= 0) { const double sd = sqrt(d2); y[n++] = (s + sd) / 2; y[n++] = (s - sd) / 2; }
    }
    for (int i = 0; i < n; ++i) {
        double x = y[i] - a / 4;
        for (int k = 0; k < 2; ++k) { const double f = poly4(c, x), df = dpoly4(c, x); if (df != 0) x -= f / df; }
        r[i] = x;
    }
    return n;
}

__device__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ bool unit3(double *a) { const double n = sqrt(dot3(a, a)); if (!(n > 1e-12)) return false; a[0] /= n; a[1] /= n; a[2] /= n; return true; }
__device__ void cross3(const double *a, const double *b, double *o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }

// orthonormal frame of a triangle: rows e1 = P2 - P1, e2 = e3 x e1, e3 = e1 x (P3 - P1)
__device__ bool frame3(const double *P1, const double *P2, const double *P3, double *E)
{
    double d1[3] = {P2[0] - P1[0], P2[1] - P1[1], P2[2] - P1[2]}, d2[3] = {P3[0] - P1[0], P3[1] - P1[1], P3[2] - P1[2]};
    if (!unit3(d1)) return false;
    double e3[3]; cross3(d1, d2, e3);
    if (!unit3(e3)) return false;
    double e2[3]; cross3(e3, d1, e2);
    for (int k = 0; k < 3; ++k) { E[k] = d1[k]; E[3 + k] = e2[k]; E[6 + k] = e3[k]; }
    return true;
}

// P3P: eliminate to a quartic in v = s3/s1 (u = s2/s1 = P(v)/Q(v)); X three world points, yb three unit bearings.
__device__ int p3p(const double (*X)[3], const double (*yb)[3], double (*out)[12])
{
    const double d23[3] = {X[1][0] - X[2][0], X[1][1] - X[2][1], X[1][2] - X[2][2]};
    const double d13[3] = {X[0][0] - X[2][0], X[0][1] - X[2][1], X[0][2] - X[2][2]};
    const double d12[3] = {X[0][0] - X[1][0], X[0][1] - X[1][1], X[0][2] - X[1][2]};
    const double a2 = dot3(d23, d23), b2 = dot3(d13, d13), c2 = dot3(d12, d12);
    if (!(a2 > 1e-12 && b2 > 1e-12 && c2 > 1e-12)) return 0;
    double Ew[9];
    if (!frame3(X[0], X[1], X[2], Ew)) return 0;
    const double ca = dot3(yb[1], yb[2]), cb = dot3(yb[0], yb[2]), cg = dot3(yb[0], yb[1]);
    const double k = (a2 - c2) / b2, m = c2 / b2;
    const double P[3] = {1 + k, -2 * k * cb, k - 1};
    const double Q[2] = {2 * cg, -2 * ca};
    const double q[3] = {1, -2 * cb, 1};
    double c[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c[i + j] += P[i] * P[j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) c[i + j] -= 2 * cg * P[i] * Q[j];
    const double Q2[3] = {Q[0] * Q[0], 2 * Q[0] * Q[1], Q[1] * Q[1]};
    const double w[3] = {1 - m * q[0], -m * q[1], -m * q[2]};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c[i + j] += w[i] * Q2[j];
    double cmax = 0;
    for (int i = 0; i < 5; ++i) cmax = fmax(cmax, fabs(c[i]));
    if (!(fabs(c[4]) > 1e-12 * cmax)) return 0;
    double roots[4];
    const int nr = quartic_roots(c, roots);
    int ns = 0;
    for (int i = 0; i < nr; ++i) {
        const double v = roots[i];
        if (!(v > 0) || !isfinite(v)) continue;
        const double Qv = Q[0] + Q[1] * v;
        if (fabs(Qv) < 1e-9) continue;
        const double u = (P[0] + (P[1] + P[2] * v) * v) / Qv;
        if (!(u > 0)) continue;
        const double qv = q[0] + (q[1] + q[2] * v) * v;
        if (!(qv > 0)) continue;
        const double s1 = sqrt(b2 / qv), s2 = u * s1, s3 = v * s1;
        const double C1[3] = {s1 * yb[0][0], s1 * yb[0][1], s1 * yb[0][2]};
        const double C2[3] = {s2 * yb[1][0], s2 * yb[1][1], s2 * yb[1][2]};
        const double C3[3] = {s3 * yb[2][0], s3 * yb[2][1], s3 * yb[2][2]};
        double Ec[9];
        if (!frame3(C1, C2, C3, Ec)) continue;
        double *R = out[ns], *t = out[ns] + 9;
        for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc)
            R[r * 3 + cc] = Ec[r] * Ew[cc] + Ec[3 + r] * Ew[3 + cc] + Ec[6 + r] * Ew[6 + cc];
        for (int r = 0; r < 3; ++r) t[r] = C1[r] - (R[r * 3] * X[0][0] + R[r * 3 + 1] * X[0][1] + R[r * 3 + 2] * X[0][2]);
        ++ns;
    }
    return ns;
}

// float32 score with separately rounded operations in the oracle's order
__device__ __forceinline__ bool is_inlier(const float *H, float X0, float X1, float X2, float o0, float o1,
                                          float fx, float fy, float cx, float cy, float thr2)
{
    const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(H[0], X0), __fmul_rn(H[1], X1)), __fmul_rn(H[2], X2)), H[9]);
    const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(H[3], X0), __fmul_rn(H[4], X1)), __fmul_rn(H[5], X2)), H[10]);
    const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(H[6], X0), __fmul_rn(H[7], X1)), __fmul_rn(H[8], X2)), H[11]);
    if (!(zc > 0.f)) return false;
    const float iz = __fdiv_rn(1.f, zc);
    const float du = __fsub_rn(__fadd_rn(__fmul_rn(fx, __fmul_rn(xc, iz)), cx), o0);
    const float dv = __fsub_rn(__fadd_rn(__fmul_rn(fy, __fmul_rn(yc, iz)), cy), o1);
    return __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)) <= thr2;
}

constexpr int MAX_IT = 512;        // RANSAC samples per problem

// One CTA per problem.  Dynamic shared memory: the points as five planes (X, Y, Z, u, v) when they fit.
__global__ void __launch_bounds__(PT) k_pnp_ransac(PoseArgs a, int smem_points)
{
    extern __shared__ float pts[];
    __shared__ double red[PW][28];
    __shared__ double tot[28];
    __shared__ double P[12];
    __shared__ int cnt[MAX_IT * 4];
    __shared__ unsigned short list[MAX_IT * 4];
    __shared__ int nh;
    __shared__ unsigned long long wbest[PW];
    __shared__ float Hb[12];
    __shared__ int stop;
    const PoseHdr &h = a.hdr[blockIdx.x];
    const int n = h.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *p3 = a.p3 + 3 * (size_t)h.off, *p2 = a.p2 + 2 * (size_t)h.off;
    const float fx = h.fx, fy = h.fy, cx = h.cx, cy = h.cy, thr2 = a.thr2;
    double *hyp = a.hyp + (size_t)blockIdx.x * a.ransac_iterations * 4 * 12;
    float *hypf = a.hypf + (size_t)blockIdx.x * a.ransac_iterations * 4 * 12;
    uint8_t *mask = a.mask + h.off;
    const bool staged = n <= smem_points;
    const int np = staged ? ((n + 31) & ~31) : 0;     // plane stride
    if (threadIdx.x == 0) nh = 0;
    for (int i = threadIdx.x; i < a.ransac_iterations * 4; i += PT) cnt[i] = -1;
    if (staged)
        for (int i = threadIdx.x; i < n; i += PT) {
            pts[i] = p3[3 * i]; pts[np + i] = p3[3 * i + 1]; pts[2 * np + i] = p3[3 * i + 2];
            pts[3 * np + i] = p2[2 * i]; pts[4 * np + i] = p2[2 * i + 1];
        }
    __syncthreads();
    // ---- hypotheses: one sample per thread
    for (int it = threadIdx.x; it < a.ransac_iterations; it += PT) {
        int s[3];
        if (!draw3(a.seed, it, n, s)) continue;
        double X[3][3], yb[3][3], sol[4][12];
        for (int k = 0; k < 3; ++k) {
            for (int j = 0; j < 3; ++j) X[k][j] = p3[3 * s[k] + j];
            yb[k][0] = ((double)p2[2 * s[k]] - (double)cx) / (double)fx;
            yb[k][1] = ((double)p2[2 * s[k] + 1] - (double)cy) / (double)fy;
            yb[k][2] = 1;
            unit3(yb[k]);
        }
        const int ns = p3p(X, yb, sol);
        for (int j = 0; j < ns; ++j) {
            bool ok = true;
            float f[12];
            for (int k = 0; k < 12; ++k) { f[k] = (float)sol[j][k]; ok &= isfinite(f[k]); }
            if (!ok) continue;
            const int slot = it * 4 + j;
            for (int k = 0; k < 12; ++k) { hyp[slot * 12 + k] = sol[j][k]; hypf[slot * 12 + k] = f[k]; }
            list[atomicAdd(&nh, 1)] = (unsigned short)slot;
        }
    }
    __syncthreads();
    // ---- score: a warp takes four hypotheses per pass over the points
    const int NH = nh;
    for (int g = warp * 4; g < NH; g += PW * 4) {
        float H[4][12]; int c[4] = {0, 0, 0, 0};
        const int m = min(4, NH - g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int slot = list[g + (j < m ? j : 0)];
#pragma unroll
            for (int k = 0; k < 12; ++k) H[j][k] = hypf[slot * 12 + k];
        }
        for (int i = lane; i < n; i += 32) {
            float X0, X1, X2, o0, o1;
            if (staged) { X0 = pts[i]; X1 = pts[np + i]; X2 = pts[2 * np + i]; o0 = pts[3 * np + i]; o1 = pts[4 * np + i]; }
            else { X0 = p3[3 * i]; X1 = p3[3 * i + 1]; X2 = p3[3 * i + 2]; o0 = p2[2 * i]; o1 = p2[2 * i + 1]; }
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] += is_inlier(H[j], X0, X1, X2, o0, o1, fx, fy, cx, cy, thr2) ? 1 : 0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int o = 16; o; o >>= 1) c[j] += __shfl_xor_sync(0xffffffffu, c[j], o);
            if (lane == 0 && j < m) cnt[list[g + j]] = c[j];
        }
    }
    __syncthreads();
    // ---- first maximum in (iteration, solution) order
    unsigned long long best = 0;
    for (int s = threadIdx.x; s < a.ransac_iterations * 4; s += PT)
        if (cnt[s] > 0) best = max(best, ((unsigned long long)cnt[s] << 16) | (unsigned)(0xffff - s));
#pragma unroll
    for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) wbest[warp] = best;
    __syncthreads();
    best = wbest[0];
#pragma unroll
    for (int w = 1; w < PW; ++w) best = max(best, wbest[w]);
    const int best_cnt = (int)(best >> 16), best_slot = best ? 0xffff - (int)(best & 0xffff) : -1;
    int *info = a.info + 4 * blockIdx.x;
    if (threadIdx.x == 0) { info[0] = best_cnt >= 3 ? best_cnt : 0; info[1] = best_slot >> 2; info[2] = best_slot < 0 ? -1 : (best_slot & 3); info[3] = NH; }
    if (best_cnt < 3) {
        for (int i = threadIdx.x; i < n; i += PT) mask[i] = 0;
        if (threadIdx.x < 12) a.pose_out[12 * (size_t)blockIdx.x + threadIdx.x] = 0;
        return;
    }
    if (threadIdx.x < 12) Hb[threadIdx.x] = hypf[best_slot * 12 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += PT)
        mask[i] = is_inlier(Hb, p3[3 * i], p3[3 * i + 1], p3[3 * i + 2], p2[2 * i], p2[2 * i + 1], fx, fy, cx, cy, thr2) ? 1 : 0;
    // ---- Gauss-Newton refit on the inliers (same left-multiplied exp update as the LM, lambda = 0, no kernel)
    Se3 est;
    if (threadIdx.x == 0) {
        const double *b = hyp + best_slot * 12;
        q_from_R(b, est);
        est.t[0] = b[9]; est.t[1] = b[10]; est.t[2] = b[11];
        q_normalize_pos(est);
        publish_pose(est, P);
        stop = 0;
    }
    __syncthreads();     // also orders the mask writes before the reads below
    const Cam K = {(double)fx, (double)fy, (double)cx, (double)cy};
    double v[28];
    for (int k = 0; k < a.refine_iterations; ++k) {
        accumulate_system(P, K, 0.0, p3, p2, n, mask, v);
        block_sum<28>(v, red, tot);
        if (threadIdx.x == 0) {
            double H[36], b[6], x[6];
            unpack_system(tot, H, b);
            if (!ldlt6_solve(H, b, x)) stop = 1;
            else {
                double nx = 0;
                for (int j = 0; j < 6; ++j) nx += x[j] * x[j];
                if (!isfinite(nx)) stop = 1;
                else {
                    const Se3 d = se3_exp(x);
                    est = se3_mul(d, est);
                    publish_pose(est, P);
                    if (nx < 1e-20) stop = 1;
                }
            }
        }
        __syncthreads();
        if (stop) break;
    }
    if (threadIdx.x < 12) a.pose_out[12 * (size_t)blockIdx.x + threadIdx.x] = P[threadIdx.x];
}

}  // namespace

int pose_max_iterations() { return MAX_IT; }

int setup_pose()
{
    return cudaFuncSetAttribute(k_pnp_ransac, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) == cudaSuccess ? 0 : -1;
}

void launch_pnp_ransac(const PoseArgs &a, int nproblems, int max_n, cudaStream_t st, long long *launches)
{
    if (nproblems <= 0) return;
    const int cap_points = 160 * 1024 / 20 - 32;
    const int smem_points = max_n <= cap_points ? ((max_n + 31) & ~31) : 0;
    k_pnp_ransac<<<nproblems, PT, (size_t)smem_points * 20, st>>>(a, smem_points);
    ++*launches;
}

void launch_pose_lm(const PoseArgs &a, int nproblems, cudaStream_t st, long long *launches)
{
    if (nproblems <= 0) return;
    k_pose_lm<<<nproblems, PT, 0, st>>>(a);
    ++*launches;
}
