// b2_ndt_math.cuh -- scalar numerics + the on-device Newton / More-Thuente controller of the NDT
// aligner.  Everything here is __host__ __device__ so the same source is (a) run by thread 0 of every
// match CTA on the GPU and (b) compiled for the host by tests (csrc/b2_hostcheck.cu) to check the
// control flow without a GPU.
//
// Reference text followed (paths relative to /root/reference/lidar_localization/):
//   NDTM = src/models/registration/ndt_registration_manual/NormalDistributionsTransform.cpp
//   computeTransformation NDTM:310-389, computeStepLengthMT NDTM:648-756, trialValueSelectionMT
//   NDTM:760-837, updateIntervalMT NDTM:840-873, computeAngleDerivatives NDTM:523-645, and the Eigen
//   3.2.92 expressions they call (JacobiSVD::solve, Transform::rotation().eulerAngles(0,1,2),
//   Translation*AngleAxis^3) -- see SURVEY.md Appendix A.
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#define B2_HD_NOINLINE __host__ __device__
#define B2_SVD_ATTR __host__ __device__ __noinline__
#else
#define B2_HD inline
#define B2_HD_NOINLINE
#define B2_SVD_ATTR inline
#endif

namespace b2 {

// ---- IEEE single operations that must not be contracted into FMAs (float pose matrix, point
// transform and voxel index arithmetic of the reference are plain x86 mul/add).
#if defined(__CUDA_ARCH__)
B2_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
B2_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
B2_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
B2_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
B2_HD float fsqrt(float a) { return __fsqrt_rn(a); }
B2_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
B2_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
B2_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
B2_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
B2_HD double dsqrt(double a) { return __dsqrt_rn(a); }
#else
// host build: compiled with -ffp-contract=off
B2_HD float fmul(float a, float b) { return a * b; }
B2_HD float fadd(float a, float b) { return a + b; }
B2_HD float fsub(float a, float b) { return a - b; }
B2_HD float fdiv(float a, float b) { return a / b; }
B2_HD float fsqrt(float a) { return sqrtf(a); }
B2_HD double dmul(double a, double b) { return a * b; }
B2_HD double dadd(double a, double b) { return a + b; }
B2_HD double dsub(double a, double b) { return a - b; }
B2_HD double ddiv(double a, double b) { return a / b; }
B2_HD double dsqrt(double a) { return sqrt(a); }
#endif

// ---- double sin / cos shared by host and device.
// Same source and same operations (explicit fma) on both sides, so the host-compiled test shim computes
// exactly what the GPU computes.  Also a toolchain necessity: CUDA 12.9's ptxas crashes on the libdevice
// sin/cos slow-path call inside a setmaxnreg region of the warp-specialised match kernel.
// Cody-Waite reduction by pi/2 with a three-part constant, then the fdlibm minimax kernels on
// [-pi/4, pi/4] (error < 1 ulp for |x| < ~1e5, far beyond any pose angle).
B2_HD void b2_sincos(double x, double &sn, double &cs) {
    const double n = rint(x * 6.36619772367581382433e-01);
    double r = fma(n, -1.57079632679489655800e+00, x);
    r = fma(n, -6.12323399573676603587e-17, r);
    r = fma(n, -8.47842766036889956997e-32, r);
    const double z = r * r;
    // __kernel_sin
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s = fma(z * r, ps, r);
    // __kernel_cos
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double c = 1.0 - fma(0.5, z, -(z * z * pc));
    const long long q = (long long)n & 3LL;
    sn = (q == 0) ? s : (q == 1) ? c : (q == 2) ? -s : -c;
    cs = (q == 0) ? c : (q == 1) ? -s : (q == 2) ? -c : s;
}
B2_HD double b2_sin(double x) { double s, c; b2_sincos(x, s, c); return s; }
B2_HD double b2_cos(double x) { double s, c; b2_sincos(x, s, c); return c; }

// float trig evaluated through double and rounded once: reproducible on host and device
B2_HD float sin_f32(float x) { return (float)b2_sin((double)x); }
B2_HD float cos_f32(float x) { return (float)b2_cos((double)x); }
B2_HD float atan2_f32(float y, float x) { return (float)atan2((double)y, (double)x); }

// ------------------------------------------------------------------ pose <-> matrix ----------
// (Translation(p0,p1,p2) * AngleAxis(p3,X) * AngleAxis(p4,Y) * AngleAxis(p5,Z)).matrix() in float
// (NDTM:370-373, 691-694); T column-major.
B2_HD void rot_axis_sc_f32(int axis, float s, float c, float R[9]) {
    float one_c = fsub(1.0f, c);
    float e[3] = {0.f, 0.f, 0.f};
    e[axis] = 1.0f;
    float sa[3] = {fmul(s, e[0]), fmul(s, e[1]), fmul(s, e[2])};
    float ca[3] = {fmul(one_c, e[0]), fmul(one_c, e[1]), fmul(one_c, e[2])};
    float tmp;
    tmp = fmul(ca[0], e[1]); R[1] = fsub(tmp, sa[2]); R[3] = fadd(tmp, sa[2]);
    tmp = fmul(ca[0], e[2]); R[2] = fadd(tmp, sa[1]); R[6] = fsub(tmp, sa[1]);
    tmp = fmul(ca[1], e[2]); R[5] = fsub(tmp, sa[0]); R[7] = fadd(tmp, sa[0]);
    for (int k = 0; k < 3; ++k) R[k * 3 + k] = fadd(fmul(ca[k], e[k]), c);
}
B2_HD void mat3_mul_f32(const float A[9], const float Bm[9], float C[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = fadd(fadd(fmul(A[i * 3 + 0], Bm[0 * 3 + j]), fmul(A[i * 3 + 1], Bm[1 * 3 + j])),
                                fmul(A[i * 3 + 2], Bm[2 * 3 + j]));
}
// fs[k] / fc[k] = sin_f32 / cos_f32 of (float)p[3+k]
B2_HD_NOINLINE inline void pose_to_matrix_sc_f32(const double p[6], const float fs[3], const float fc[3], float T[16]) {
    float Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
    rot_axis_sc_f32(0, fs[0], fc[0], Rx);
    rot_axis_sc_f32(1, fs[1], fc[1], Ry);
    rot_axis_sc_f32(2, fs[2], fc[2], Rz);
    mat3_mul_f32(Rx, Ry, Rxy);
    mat3_mul_f32(Rxy, Rz, R);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[c * 4 + r] = R[r * 3 + c];
    T[3] = T[7] = T[11] = 0.0f;
    T[12] = (float)p[0]; T[13] = (float)p[1]; T[14] = (float)p[2]; T[15] = 1.0f;
}
B2_HD_NOINLINE inline void pose_to_matrix_f32(const double p[6], float T[16]) {
    float fs[3], fc[3];
    for (int k = 0; k < 3; ++k) { fs[k] = sin_f32((float)p[3 + k]); fc[k] = cos_f32((float)p[3 + k]); }
    pose_to_matrix_sc_f32(p, fs, fc, T);
}

// Transform<float,3,Affine>::rotation() = U V^T of JacobiSVD<Matrix3f>(linear) (Eigen Transform.h
// :1057-1073), then eulerAngles(0,1,2) (EulerAngles.h:36-99).  Tin column-major 4x4.
B2_HD void jac_make_f(float x, float y, float z, float &c, float &s) {
    if (y == 0.0f) { c = 1.0f; s = 0.0f; return; }
    float tau = fdiv(fsub(x, z), fmul(2.0f, fabsf(y)));
    float w = fsqrt(fadd(fmul(tau, tau), 1.0f));
    float t = (tau > 0.0f) ? fdiv(1.0f, fadd(tau, w)) : fdiv(1.0f, fsub(tau, w));
    float sign_t = t > 0.0f ? 1.0f : -1.0f;
    float n = fdiv(1.0f, fsqrt(fadd(fmul(t, t), 1.0f)));
    s = fmul(fmul(fmul(-sign_t, fdiv(y, fabsf(y))), fabsf(t)), n);
    c = n;
}
B2_HD void rot_plane_f(float *x, int incx, float *y, int incy, int n, float c, float s) {
    if (c == 1.0f && s == 0.0f) return;
    for (int i = 0; i < n; ++i) {
        float xi = x[i * incx], yi = y[i * incy];
        x[i * incx] = fadd(fmul(c, xi), fmul(s, yi));
        y[i * incy] = fadd(fmul(-s, xi), fmul(c, yi));
    }
}
B2_HD_NOINLINE inline void euler_from_matrix_f32(const float Tin[16], float out[3]) {
    enum { N = 3 };
    float W[9], U[9], V[9];
#define F_(A, r, c) (A)[(c) * N + (r)]
    const float precision = 2.0f * FLT_EPSILON;
    const float considerAsZero = 2.0f * 1.40129846e-45f;
    float scale = 0.0f;
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) { float a = fabsf(Tin[c * 4 + r]); if (a > scale) scale = a; }
    if (scale == 0.0f) scale = 1.0f;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            F_(W, r, c) = fdiv(Tin[c * 4 + r], scale);
            F_(U, r, c) = F_(V, r, c) = (r == c) ? 1.0f : 0.0f;
        }
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 1000) {
        finished = true;
        for (int p = 1; p < N; ++p)
            for (int q = 0; q < p; ++q) {
                float mpp = fabsf(F_(W, p, p)), mqq = fabsf(F_(W, q, q));
                float thr = fmul(precision, (mpp > mqq ? mpp : mqq));
                if (!(thr > considerAsZero)) thr = considerAsZero;
                if (fabsf(F_(W, p, q)) > thr || fabsf(F_(W, q, p)) > thr) {
                    finished = false;
                    float m00 = F_(W, p, p), m01 = F_(W, p, q), m10 = F_(W, q, p), m11 = F_(W, q, q);
                    float r1c, r1s;
                    float t = fadd(m00, m11), d = fsub(m10, m01);
                    if (d == 0.0f) { r1s = 0.0f; r1c = 1.0f; }
                    else {
                        float u = fdiv(t, d);
                        float tmp = fsqrt(fadd(1.0f, fmul(u, u)));
                        r1s = fdiv(1.0f, tmp);
                        r1c = fdiv(u, tmp);
                    }
                    float a00 = fadd(fmul(r1c, m00), fmul(r1s, m10)), a01 = fadd(fmul(r1c, m01), fmul(r1s, m11));
                    float a11 = fadd(fmul(-r1s, m01), fmul(r1c, m11));
                    float jrc, jrs;
                    jac_make_f(a00, a01, a11, jrc, jrs);
                    float jlc = fsub(fmul(r1c, jrc), fmul(r1s, -jrs));
                    float jls = fadd(fmul(r1c, -jrs), fmul(r1s, jrc));
                    rot_plane_f(&F_(W, p, 0), N, &F_(W, q, 0), N, N, jlc, jls);
                    rot_plane_f(&F_(U, 0, p), 1, &F_(U, 0, q), 1, N, jlc, jls);
                    rot_plane_f(&F_(W, 0, p), 1, &F_(W, 0, q), 1, N, jrc, -jrs);
                    rot_plane_f(&F_(V, 0, p), 1, &F_(V, 0, q), 1, N, jrc, -jrs);
                }
            }
    }
    float sv[3];
    for (int i = 0; i < N; ++i) {
        float wii = F_(W, i, i);
        float a = fabsf(wii);
        sv[i] = a;
        if (a != 0.0f) { float f = fdiv(wii, a); for (int r = 0; r < N; ++r) F_(U, r, i) = fmul(F_(U, r, i), f); }
    }
    for (int i = 0; i < N; ++i) {
        int pos = i;
        float mx = sv[i];
        for (int j = i + 1; j < N; ++j) if (sv[j] > mx) { mx = sv[j]; pos = j; }
        if (mx == 0.0f) break;
        if (pos != i) {
            float t = sv[i]; sv[i] = sv[pos]; sv[pos] = t;
            for (int r = 0; r < N; ++r) {
                t = F_(U, r, i); F_(U, r, i) = F_(U, r, pos); F_(U, r, pos) = t;
                t = F_(V, r, i); F_(V, r, i) = F_(V, r, pos); F_(V, r, pos) = t;
            }
        }
    }
    float UVt[9], Rm[9];
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c)
            F_(UVt, r, c) = fadd(fadd(fmul(F_(U, r, 0), F_(V, c, 0)), fmul(F_(U, r, 1), F_(V, c, 1))), fmul(F_(U, r, 2), F_(V, c, 2)));
#define DET3H(m, a, b, c) fmul(F_(m, 0, a), fsub(fmul(F_(m, 1, b), F_(m, 2, c)), fmul(F_(m, 1, c), F_(m, 2, b))))
    float x = fadd(fsub(DET3H(UVt, 0, 1, 2), DET3H(UVt, 1, 0, 2)), DET3H(UVt, 2, 0, 1));
#undef DET3H
    for (int r = 0; r < N; ++r) F_(U, r, 0) = fdiv(F_(U, r, 0), x);
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c)
            F_(Rm, r, c) = fadd(fadd(fmul(F_(U, r, 0), F_(V, c, 0)), fmul(F_(U, r, 1), F_(V, c, 1))), fmul(F_(U, r, 2), F_(V, c, 2)));
    // eulerAngles(0,1,2): odd = 0, i=0, j=1, k=2
    const float pi = (float)3.14159265358979323846;
    float r0 = atan2_f32(F_(Rm, 1, 2), F_(Rm, 2, 2));
    float c2 = fsqrt(fadd(fmul(F_(Rm, 0, 0), F_(Rm, 0, 0)), fmul(F_(Rm, 0, 1), F_(Rm, 0, 1))));
    float r1;
    if (r0 > 0.0f) { r0 = fsub(r0, pi); r1 = atan2_f32(-F_(Rm, 0, 2), -c2); }
    else           { r1 = atan2_f32(-F_(Rm, 0, 2), c2); }
    float s1 = sin_f32(r0), c1 = cos_f32(r0);
    float r2 = atan2_f32(fsub(fmul(s1, F_(Rm, 2, 0)), fmul(c1, F_(Rm, 1, 0))), fsub(fmul(c1, F_(Rm, 1, 1)), fmul(s1, F_(Rm, 2, 1))));
    out[0] = -r0; out[1] = -r1; out[2] = -r2;
#undef F_
}

// ------------------------------------------------------------------ angle derivative tables --
// NDTM:523-645.  j[8][3] = j_ang_a..h ; h[15][3] = h_ang_a2,a3,b2,b3,c2,c3,d1,d2,d3,e1,e2,e3,f1,f2,f3
struct AngTab {
    double j[8][3];
    double h[15][3];
};
// snapped trig of NDTM:527-549: |angle| < 10e-5 -> (cos, sin) = (1, 0)
B2_HD double ang_sin(double a) { return (fabs(a) < 10e-5) ? 0.0 : b2_sin(a); }
B2_HD double ang_cos(double a) { return (fabs(a) < 10e-5) ? 1.0 : b2_cos(a); }
// parts: bit 0 = j_ang_a..h, bit 1 = h_ang_a2..d2, bit 2 = h_ang_d3..f3 (every entry is an expression of its own, so
// three lanes of the controller warp can each write a part)
B2_HD_NOINLINE inline void angle_derivatives_sc(double sx, double cx, double sy, double cy, double sz, double cz, AngTab &A, int parts = 7);
B2_HD_NOINLINE inline void angle_derivatives(const double p[6], AngTab &A) {
    angle_derivatives_sc(ang_sin(p[3]), ang_cos(p[3]), ang_sin(p[4]), ang_cos(p[4]), ang_sin(p[5]), ang_cos(p[5]), A);
}
B2_HD_NOINLINE inline void angle_derivatives_sc(double sx, double cx, double sy, double cy, double sz, double cz, AngTab &A, int parts) {
    double (*j)[3] = A.j;
    double (*h)[3] = A.h;
    if (parts & 1) {
    j[0][0] = -sx * sz + cx * sy * cz; j[0][1] = -sx * cz - cx * sy * sz; j[0][2] = -cx * cy;
    j[1][0] = cx * sz + sx * sy * cz;  j[1][1] = cx * cz - sx * sy * sz;  j[1][2] = -sx * cy;
    j[2][0] = -sy * cz;                j[2][1] = sy * sz;                 j[2][2] = cy;
    j[3][0] = sx * cy * cz;            j[3][1] = -sx * cy * sz;           j[3][2] = sx * sy;
    j[4][0] = -cx * cy * cz;           j[4][1] = cx * cy * sz;            j[4][2] = -cx * sy;
    j[5][0] = -cy * sz;                j[5][1] = -cy * cz;                j[5][2] = 0;
    j[6][0] = cx * cz - sx * sy * sz;  j[6][1] = -cx * sz - sx * sy * cz; j[6][2] = 0;
    j[7][0] = sx * cz + cx * sy * sz;  j[7][1] = cx * sy * cz - sx * sz;  j[7][2] = 0;
    }
    if (parts & 2) {
    h[0][0] = -cx * sz - sx * sy * cz; h[0][1] = -cx * cz + sx * sy * sz; h[0][2] = sx * cy;
    h[1][0] = -sx * sz + cx * sy * cz; h[1][1] = -cx * sy * sz - sx * cz; h[1][2] = -cx * cy;
    h[2][0] = cx * cy * cz;            h[2][1] = -cx * cy * sz;           h[2][2] = cx * sy;
    h[3][0] = sx * cy * cz;            h[3][1] = -sx * cy * sz;           h[3][2] = sx * sy;
    h[4][0] = -sx * cz - cx * sy * sz; h[4][1] = sx * sz - cx * sy * cz;  h[4][2] = 0;
    h[5][0] = cx * cz - sx * sy * sz;  h[5][1] = -sx * sy * cz - cx * sz; h[5][2] = 0;
    h[6][0] = -cy * cz;                h[6][1] = cy * sz;                 h[6][2] = sy;
    h[7][0] = -sx * sy * cz;           h[7][1] = sx * sy * sz;            h[7][2] = sx * cy;
    }
    if (parts & 4) {
    h[8][0] = cx * sy * cz;            h[8][1] = -cx * sy * sz;           h[8][2] = -cx * cy;
    h[9][0] = sy * sz;                 h[9][1] = sy * cz;                 h[9][2] = 0;
    h[10][0] = -sx * cy * sz;          h[10][1] = -sx * cy * cz;          h[10][2] = 0;
    h[11][0] = cx * cy * sz;           h[11][1] = cx * cy * cz;           h[11][2] = 0;
    h[12][0] = -cy * cz;               h[12][1] = cy * sz;                h[12][2] = 0;
    h[13][0] = -cx * sz - sx * sy * cz;h[13][1] = -cx * cz + sx * sy * sz;h[13][2] = 0;
    h[14][0] = -sx * sz + cx * sy * cz;h[14][1] = -cx * sy * sz - sx * cz;h[14][2] = 0;
    }
}

// ------------------------------------------------------------------ 3x3 double numerics -------
// Used by the target-grid build; written with explicit non-contracted operations and in the same
// operation order as oracle/ndt_oracle.c so that per-voxel covariances / inverses agree bit for bit.
B2_HD double cof3(const double *m, int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return dsub(dmul(m[i1 * 3 + j1], m[i2 * 3 + j2]), dmul(m[i1 * 3 + j2], m[i2 * 3 + j1]));
}
B2_HD_NOINLINE inline void inverse3(const double A[9], double out[9]) {
    double c0[3] = {cof3(A, 0, 0), cof3(A, 1, 0), cof3(A, 2, 0)};
    double det = dadd(dadd(dmul(c0[0], A[0]), dmul(c0[1], A[3])), dmul(c0[2], A[6]));
    double invdet = ddiv(1.0, det);
    out[0] = dmul(c0[0], invdet); out[1] = dmul(c0[1], invdet); out[2] = dmul(c0[2], invdet);
    out[3] = dmul(cof3(A, 0, 1), invdet); out[4] = dmul(cof3(A, 1, 1), invdet); out[5] = dmul(cof3(A, 2, 1), invdet);
    out[6] = dmul(cof3(A, 0, 2), invdet); out[7] = dmul(cof3(A, 1, 2), invdet); out[8] = dmul(cof3(A, 2, 2), invdet);
}
// symmetric 3x3 eigen decomposition (cyclic Jacobi, ascending), A row-major, V columns = vectors
B2_HD_NOINLINE inline void eig3_sym(const double Ain[9], double evals[3], double V[9]) {
    double A[9];
    for (int i = 0; i < 9; ++i) A[i] = Ain[i];
    A[1] = A[3]; A[2] = A[6]; A[5] = A[7];
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = dadd(dadd(fabs(A[1]), fabs(A[2])), fabs(A[5]));
        double dia = dadd(dadd(fabs(A[0]), fabs(A[4])), fabs(A[8]));
        if (off == 0.0 || off <= 1e-300 || off < dmul(1e-22, dia)) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double apq = A[p * 3 + q];
                if (apq == 0.0) continue;
                double app = A[p * 3 + p], aqq = A[q * 3 + q];
                double theta = ddiv(dsub(aqq, app), dmul(2.0, apq));
                double t = ddiv((theta >= 0 ? 1.0 : -1.0), dadd(fabs(theta), dsqrt(dadd(dmul(theta, theta), 1.0))));
                double c = ddiv(1.0, dsqrt(dadd(dmul(t, t), 1.0))), s = dmul(t, c);
                for (int k = 0; k < 3; ++k) {
                    double akp = A[k * 3 + p], akq = A[k * 3 + q];
                    A[k * 3 + p] = dsub(dmul(c, akp), dmul(s, akq));
                    A[k * 3 + q] = dadd(dmul(s, akp), dmul(c, akq));
                }
                for (int k = 0; k < 3; ++k) {
                    double apk = A[p * 3 + k], aqk = A[q * 3 + k];
                    A[p * 3 + k] = dsub(dmul(c, apk), dmul(s, aqk));
                    A[q * 3 + k] = dadd(dmul(s, apk), dmul(c, aqk));
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                    V[k * 3 + p] = dsub(dmul(c, vkp), dmul(s, vkq));
                    V[k * 3 + q] = dadd(dmul(s, vkp), dmul(c, vkq));
                }
            }
    }
    evals[0] = A[0]; evals[1] = A[4]; evals[2] = A[8];
    for (int i = 0; i < 2; ++i) {
        int k = i;
        for (int j = i + 1; j < 3; ++j) if (evals[j] < evals[k]) k = j;
        if (k != i) {
            double t = evals[i]; evals[i] = evals[k]; evals[k] = t;
            for (int r = 0; r < 3; ++r) { t = V[r * 3 + i]; V[r * 3 + i] = V[r * 3 + k]; V[r * 3 + k] = t; }
        }
    }
}

// second pass of VoxelGridCovariance::applyFilter for one leaf (Appendix A.2; NDTM/VoxelGrid.cpp
// :273-323).  sum[3], acc[9] (row-major, accumulated from Identity).  Returns nr_points (n or -1).
B2_HD_NOINLINE inline int leaf_finish(const double sum[3], const double acc[9], int n, int min_pts, double eig_mult,
                                      double mean[3], double cov[9], double icov[9], double evals[3]) {
    const double nd = (double)n;
    for (int a = 0; a < 3; ++a) mean[a] = ddiv(sum[a], nd);
    for (int a = 0; a < 9; ++a) { cov[a] = 0.0; icov[a] = 0.0; }
    evals[0] = evals[1] = evals[2] = 0.0;
    if (n < min_pts) return n;
    const double f = ddiv(dsub(nd, 1.0), nd);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double t = dmul(sum[a], mean[b]);
            double mm = dmul(mean[a], mean[b]);
            double v = dadd(ddiv(dsub(acc[a * 3 + b], dmul(2.0, t)), nd), mm);
            cov[a * 3 + b] = dmul(v, f);
        }
    double V[9];
    eig3_sym(cov, evals, V);
    if (evals[0] < 0 || evals[1] < 0 || evals[2] <= 0) return -1;
    double mn = dmul(eig_mult, evals[2]);
    if (evals[0] < mn) {
        evals[0] = mn;
        if (evals[1] < mn) evals[1] = mn;
        double Vi[9], VD[9];
        inverse3(V, Vi);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) VD[a * 3 + b] = dmul(V[a * 3 + b], evals[b]);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                cov[a * 3 + b] = dadd(dadd(dmul(VD[a * 3 + 0], Vi[0 * 3 + b]), dmul(VD[a * 3 + 1], Vi[1 * 3 + b])), dmul(VD[a * 3 + 2], Vi[2 * 3 + b]));
    }
    inverse3(cov, icov);
    double mxc = -DBL_MAX, mnc = DBL_MAX;
    for (int a = 0; a < 9; ++a) { if (icov[a] > mxc) mxc = icov[a]; if (icov[a] < mnc) mnc = icov[a]; }
    if (mxc == (double)INFINITY || mnc == -(double)INFINITY) return -1;
    return n;
}

// ------------------------------------------------------------------ 6x6 Newton solve ----------
// delta = JacobiSVD(H).solve(b)   (NDTM:353-355).  H column-major.
// Full two-sided Jacobi SVD as Eigen runs it (JacobiSVD.h:675-785, SVDBase.h:130-139,260-272).
B2_HD void jac_make_d(double x, double y, double z, double &c, double &s) {
    if (y == 0.0) { c = 1.0; s = 0.0; return; }
    double tau = (x - z) / (2.0 * fabs(y));
    double w = sqrt(tau * tau + 1.0);
    double t = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
    double sign_t = t > 0.0 ? 1.0 : -1.0;
    double n = 1.0 / sqrt(t * t + 1.0);
    s = -sign_t * (y / fabs(y)) * fabs(t) * n;
    c = n;
}
B2_HD void rot_plane_d(double *x, int incx, double *y, int incy, int n, double c, double s) {
    if (c == 1.0 && s == 0.0) return;
    for (int i = 0; i < n; ++i) {
        double xi = x[i * incx], yi = y[i * incy];
        x[i * incx] = c * xi + s * yi;
        y[i * incy] = -s * xi + c * yi;
    }
}
B2_SVD_ATTR int svd_solve6(const double Hin[36], const double b[6], double x[6]) {
    enum { N = 6 };
    double W[36], U[36], V[36], sv[6];
#define M_(A, r, c) (A)[(c) * N + (r)]
    const double precision = 2.0 * DBL_EPSILON;
    const double considerAsZero = 2.0 * 4.9406564584124654e-324;
    double scale = 0.0;
    for (int i = 0; i < 36; ++i) { double a = fabs(Hin[i]); if (a > scale) scale = a; }
    if (scale == 0.0) scale = 1.0;
    for (int i = 0; i < 36; ++i) { W[i] = Hin[i] / scale; U[i] = V[i] = (i % 7 == 0) ? 1.0 : 0.0; }
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 1000) {
        finished = true;
        for (int p = 1; p < N; ++p)
            for (int q = 0; q < p; ++q) {
                double mpp = fabs(M_(W, p, p)), mqq = fabs(M_(W, q, q));
                double thr = precision * (mpp > mqq ? mpp : mqq);
                if (!(thr > considerAsZero)) thr = considerAsZero;
                if (fabs(M_(W, p, q)) > thr || fabs(M_(W, q, p)) > thr) {
                    finished = false;
                    double m00 = M_(W, p, p), m01 = M_(W, p, q), m10 = M_(W, q, p), m11 = M_(W, q, q);
                    double r1c, r1s;
                    double t = m00 + m11, d = m10 - m01;
                    if (d == 0.0) { r1s = 0.0; r1c = 1.0; }
                    else {
                        double u = t / d;
                        double tmp = sqrt(1.0 + u * u);
                        r1s = 1.0 / tmp;
                        r1c = u / tmp;
                    }
                    double a00 = r1c * m00 + r1s * m10, a01 = r1c * m01 + r1s * m11;
                    double a11 = -r1s * m01 + r1c * m11;
                    double jrc, jrs;
                    jac_make_d(a00, a01, a11, jrc, jrs);
                    double jlc = r1c * jrc - r1s * (-jrs);
                    double jls = r1c * (-jrs) + r1s * jrc;
                    rot_plane_d(&M_(W, p, 0), N, &M_(W, q, 0), N, N, jlc, jls);
                    rot_plane_d(&M_(U, 0, p), 1, &M_(U, 0, q), 1, N, jlc, jls);
                    rot_plane_d(&M_(W, 0, p), 1, &M_(W, 0, q), 1, N, jrc, -jrs);
                    rot_plane_d(&M_(V, 0, p), 1, &M_(V, 0, q), 1, N, jrc, -jrs);
                }
            }
    }
    for (int i = 0; i < N; ++i) {
        double wii = M_(W, i, i);
        double a = fabs(wii);
        sv[i] = a;
        if (a != 0.0) { double f = wii / a; for (int r = 0; r < N; ++r) M_(U, r, i) *= f; }
    }
    for (int i = 0; i < N; ++i) sv[i] *= scale;
    int nonzero = N;
    for (int i = 0; i < N; ++i) {
        int pos = i;
        double mx = sv[i];
        for (int j = i + 1; j < N; ++j) if (sv[j] > mx) { mx = sv[j]; pos = j; }
        if (mx == 0.0) { nonzero = i; break; }
        if (pos != i) {
            double t = sv[i]; sv[i] = sv[pos]; sv[pos] = t;
            for (int r = 0; r < N; ++r) {
                t = M_(U, r, i); M_(U, r, i) = M_(U, r, pos); M_(U, r, pos) = t;
                t = M_(V, r, i); M_(V, r, i) = M_(V, r, pos); M_(V, r, pos) = t;
            }
        }
    }
    double pthr = sv[0] * (6.0 * DBL_EPSILON);
    if (!(pthr > DBL_MIN)) pthr = (pthr != pthr) ? pthr : DBL_MIN;
    int i = nonzero - 1;
    while (i >= 0 && sv[i] < pthr) --i;
    int rank = i + 1;
    double tmp[6];
    for (int k = 0; k < rank; ++k) {
        double acc = 0.0;
        for (int r = 0; r < N; ++r) acc += M_(U, r, k) * b[r];
        tmp[k] = (1.0 / sv[k]) * acc;
    }
    for (int r = 0; r < N; ++r) {
        double acc = 0.0;
        for (int k = 0; k < rank; ++k) acc += M_(V, r, k) * tmp[k];
        x[r] = acc;
    }
#undef M_
    return rank;
}

// Fast path: for a numerically full-rank H the truncated-SVD solve IS H^-1 b, which Gaussian
// elimination with partial pivoting delivers in ~1/50 of the serial work.  Returns an estimate of
// min|pivot| / max|pivot|; the caller falls back to svd_solve6 when it is tiny or not finite, so the
// rank-truncation semantics of Eigen's solve() are kept for (near-)singular Hessians.
B2_HD_NOINLINE inline double lu_solve6(const double Hin[36], const double b[6], double x[6]) {
    enum { N = 6 };
    // fully unrolled with static indices so the augmented matrix lives in registers (no local memory):
    // partial pivoting is done by compare-and-swap sweeps that leave the largest |entry| in row k.
    double A[N][N + 1];
#pragma unroll
    for (int r = 0; r < N; ++r) {
#pragma unroll
        for (int c = 0; c < N; ++c) A[r][c] = Hin[c * N + r];
        A[r][N] = b[r];
    }
    double pmin = DBL_MAX, pmax = 0.0;
    bool singular = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int r = k + 1; r < N; ++r) {
            const bool sw = fabs(A[r][k]) > fabs(A[k][k]);
#pragma unroll
            for (int c = k; c <= N; ++c) {
                const double u = A[k][c], v = A[r][c];
                A[k][c] = sw ? v : u;
                A[r][c] = sw ? u : v;
            }
        }
        const double best = fabs(A[k][k]);
        if (!(best > 0.0)) singular = true;   // zero or NaN pivot
        pmin = fmin(pmin, best);
        pmax = fmax(pmax, best);
        const double inv = 1.0 / A[k][k];
#pragma unroll
        for (int r = k + 1; r < N; ++r) {
            const double f = A[r][k] * inv;
#pragma unroll
            for (int c = k + 1; c <= N; ++c) A[r][c] -= f * A[k][c];
        }
    }
    // back substitution with the reciprocals of the pivots (the warp version computes the six of them side by side:
    // one division latency instead of six in the dependent chain)
    double rinv[N];
#pragma unroll
    for (int r = 0; r < N; ++r) rinv[r] = 1.0 / A[r][r];
    double xs[N];
#pragma unroll
    for (int r = N - 1; r >= 0; --r) {
        double s = A[r][N];
#pragma unroll
        for (int c = r + 1; c < N; ++c) s -= A[r][c] * xs[c];
        xs[r] = s * rinv[r];
    }
#pragma unroll
    for (int r = 0; r < N; ++r) x[r] = xs[r];
    if (singular) return 0.0;
    return pmin / pmax;
}

B2_HD_NOINLINE inline void newton_solve6(const double H[36], const double b[6], double x[6], int force_svd) {
    if (!force_svd) {
        double rc = lu_solve6(H, b, x);
        bool fin = true;
        for (int i = 0; i < 6; ++i) fin = fin && (x[i] == x[i]) && (fabs(x[i]) <= DBL_MAX);
        if (rc > 1e-9 && fin) return;
    }
    svd_solve6(H, b, x);
}

// ------------------------------------------------------------------ More-Thuente helpers ------
B2_HD double psi_mt(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
B2_HD double dpsi_mt(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

B2_HD_NOINLINE inline double trial_value_selection_mt(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u,
                                                      double a_t, double f_t, double g_t) {
    if (f_t > f_l) {
        double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
        double w = sqrt(z * z - g_t * g_l);
        double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
        if (fabs(a_c - a_l) < fabs(a_q - a_l)) return a_c;
        return 0.5 * (a_q + a_c);
    } else if (g_t * g_l < 0) {
        double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
        double w = sqrt(z * z - g_t * g_l);
        double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        if (fabs(a_c - a_t) >= fabs(a_s - a_t)) return a_c;
        return a_s;
    } else if (fabs(g_t) <= fabs(g_l)) {
        double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
        double w = sqrt(z * z - g_t * g_l);
        double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        double a_t_next = (fabs(a_c - a_t) < fabs(a_s - a_t)) ? a_c : a_s;
        if (a_t > a_l) return fmin(a_t + 0.66 * (a_u - a_t), a_t_next);
        return fmax(a_t + 0.66 * (a_u - a_t), a_t_next);
    } else {
        double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
        double w = sqrt(z * z - g_t * g_u);
        return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
    }
}
B2_HD_NOINLINE inline int update_interval_mt(double &a_l, double &f_l, double &g_l, double &a_u, double &f_u, double &g_u,
                                             double a_t, double f_t, double g_t) {
    if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return 0; }
    else if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return 0; }
    else if (g_t * (a_l - a_t) < 0) {
        a_u = a_l; f_u = f_l; g_u = g_l;
        a_l = a_t; f_l = f_t; g_l = g_t;
        return 0;
    }
    return 1;
}

// ------------------------------------------------------------------ controller ----------------
// The aligner is a loop "evaluate one derivative pass -> controller decides the next pass or stops".
// A pass is described by PassReq (float transform for the points, angle tables, Hessian on/off) and
// returns Acc (score, gradient, upper Hessian, pair count).
constexpr int ACC_N = 29;   // [0] score, [1..6] gradient, [7..27] Hessian upper (row-major i<=j), [28] pairs

struct NdtConst {
    double d1, d2;          // gauss_d1_, gauss_d2_
    double step_size, trans_eps;
    int    max_iter, pcl17_compat, force_svd;
    float  res;
};

enum CtlState { ST_INIT = 0, ST_MT_FIRST = 1, ST_MT_TRIAL = 2, ST_MT_HESS = 3, ST_DONE = 4 };

struct Ctl {
    // request for the next pass
    float  T[16];
    AngTab ang;
    int    hess;
    int    state;
    int    need_trig;   // 0: request complete; 1: T and angle tables must be built from x_req; 2: tables only
    double x_req[6];    // pose the pending request refers to
    // optimiser state
    double p[6], x_t[6], dir[6];
    double score, g[6], H[36];
    double phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t, psi_t, d_psi_t, step_min, step_max;
    int    open_interval, interval_converged, step_iterations;
    int    nr_iter, converged, passes, mt_trials;
    long long pairs;
    float  finalT[16];
    double trans_probability;
    double npoints;
};

B2_HD void ctl_unpack_acc(Ctl &c, const double *acc, bool take_score_grad, bool take_hess) {
    if (take_score_grad) {
        c.score = acc[0];
        for (int i = 0; i < 6; ++i) c.g[i] = acc[1 + i];
    }
    if (take_hess) {
        int k = 7;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j) { c.H[j * 6 + i] = acc[k]; c.H[i * 6 + j] = acc[k]; ++k; }
    }
    c.pairs += (long long)acc[28];
    c.passes++;
}

// A request is completed in two steps so that the twelve sin/cos evaluations can be spread over the
// lanes of a warp on the GPU: ctl_request records the pose, ctl_finish_request builds the float pose
// matrix (NDTM:691-694) and the angle tables (NDTM:523-645) from the trig values.
B2_HD void ctl_request(Ctl &c, const double x[6], int hess, int state) {
    for (int i = 0; i < 6; ++i) c.x_req[i] = x[i];
    c.need_trig = 1;
    c.hess = hess;
    c.state = state;
}
// fs/fc: float sin/cos of (float)x_req[3+k] (pose matrix);  ds/dc: snapped double sin/cos of x_req[3+k]
B2_HD void ctl_trig_serial(const Ctl &c, float fs[3], float fc[3], double ds[3], double dc[3]) {
    for (int k = 0; k < 3; ++k) {
        fs[k] = sin_f32((float)c.x_req[3 + k]); fc[k] = cos_f32((float)c.x_req[3 + k]);
        ds[k] = ang_sin(c.x_req[3 + k]); dc[k] = ang_cos(c.x_req[3 + k]);
    }
}
// who = -1: everything (serial); who = 0 / 1 / 2: the share of lane `who` of the controller warp (lane 0: pose matrix +
// first part of the angle tables, lanes 1 and 2: the other two parts); the caller clears need_trig afterwards
B2_HD_NOINLINE inline void ctl_finish_request(Ctl &c, const float fs[3], const float fc[3], const double ds[3], const double dc[3], int who = -1) {
    if (c.need_trig == 1 && who <= 0) {
        pose_to_matrix_sc_f32(c.x_req, fs, fc, c.T);
        for (int i = 0; i < 16; ++i) c.finalT[i] = c.T[i];
    }
    angle_derivatives_sc(ds[0], dc[0], ds[1], dc[1], ds[2], dc[2], c.ang, who < 0 ? 7 : (1 << who));
    if (who < 0) c.need_trig = 0;
}
B2_HD void ctl_finish_request_serial(Ctl &c) {
    if (!c.need_trig) return;
    float fs[3], fc[3];
    double ds[3], dc[3];
    ctl_trig_serial(c, fs, fc, ds, dc);
    ctl_finish_request(c, fs, fc, ds, dc);
}

// start: guess (column-major float 4x4).  Sets up the initial pass (NDTM:323-346).
B2_HD_NOINLINE inline void ctl_start(Ctl &c, const NdtConst &k, const float guess[16], double npoints) {
    (void)k;
    bool ident = true;
    for (int i = 0; i < 16; ++i) ident = ident && (guess[i] == ((i % 5 == 0) ? 1.0f : 0.0f));
    for (int i = 0; i < 16; ++i) c.finalT[i] = ident ? ((i % 5 == 0) ? 1.0f : 0.0f) : guess[i];
    for (int i = 0; i < 16; ++i) c.T[i] = c.finalT[i];
    float ang[3];
    euler_from_matrix_f32(c.finalT, ang);
    c.p[0] = (double)c.finalT[12]; c.p[1] = (double)c.finalT[13]; c.p[2] = (double)c.finalT[14];
    c.p[3] = (double)ang[0]; c.p[4] = (double)ang[1]; c.p[5] = (double)ang[2];
    for (int i = 0; i < 6; ++i) c.x_req[i] = c.p[i];
    c.need_trig = 2;          // T is the guess itself; only the angle tables are needed
    c.hess = 1;
    c.state = ST_INIT;
    c.nr_iter = 0; c.converged = 0; c.passes = 0; c.mt_trials = 0; c.pairs = 0;
    c.score = 0.0; c.trans_probability = 0.0; c.npoints = npoints;
    for (int i = 0; i < 6; ++i) { c.g[i] = 0.0; c.dir[i] = 0.0; c.x_t[i] = c.p[i]; }
    for (int i = 0; i < 36; ++i) c.H[i] = 0.0;
}

// The controller step is split so that the GPU can run the 6x6 Newton solve warp-cooperatively between the
// two serial halves:  ctl_pre consumes the finished pass (state machine, More-Thuente bookkeeping, pose
// update + convergence test) and answers CTL_DONE, CTL_PASS (another pass is requested: c.T / c.ang / c.hess
// describe it) or CTL_NEWTON (solve H delta = -g, then call ctl_post_newton(delta), which answers the same
// three codes).  ctl_step is the serial composition (host tests, oracle pins).
enum CtlCode { CTL_DONE = 0, CTL_PASS = 1, CTL_NEWTON = 2 };

// NDTM:368-383: apply the accepted step, test convergence
B2_HD_NOINLINE inline int ctl_finish_iter(Ctl &c, const NdtConst &k) {
    double delta_p_norm = c.a_t;
    for (int i = 0; i < 6; ++i) c.p[i] = c.p[i] + c.dir[i] * delta_p_norm;
    if (c.nr_iter > k.max_iter || (c.nr_iter && (fabs(delta_p_norm) < k.trans_eps))) c.converged = 1;
    c.nr_iter++;
    if (c.converged) {
        c.trans_probability = (c.npoints > 0) ? c.score / c.npoints : 0.0;
        c.state = ST_DONE;
        return CTL_DONE;
    }
    return CTL_NEWTON;
}

B2_HD_NOINLINE inline int ctl_pre(Ctl &c, const NdtConst &k, const double *acc) {
    const double mu = 1.e-4, nu = 0.9;
    const int max_step_iterations = 10;
    bool mt_check = false;      // evaluate the More-Thuente loop condition

    if (c.state == ST_INIT) {
        ctl_unpack_acc(c, acc, true, true);
        return CTL_NEWTON;
    } else if (c.state == ST_MT_FIRST) {
        ctl_unpack_acc(c, acc, true, true);
        c.phi_t = -c.score;
        double d = 0.0;
        for (int i = 0; i < 6; ++i) d += c.g[i] * c.dir[i];
        c.d_phi_t = -d;
        c.psi_t = psi_mt(c.a_t, c.phi_t, c.phi_0, c.d_phi_0, mu);
        c.d_psi_t = dpsi_mt(c.d_phi_t, c.d_phi_0, mu);
        mt_check = true;
    } else if (c.state == ST_MT_TRIAL) {
        // computeDerivatives(..., compute_hessian=false) zeroes the Hessian (NDTM:402-403)
        ctl_unpack_acc(c, acc, true, false);
        for (int i = 0; i < 36; ++i) c.H[i] = 0.0;
        c.mt_trials++;
        c.phi_t = -c.score;                       // PCL: phi_t = -score (in-tree copy accumulates; A.4)
        double d = 0.0;
        for (int i = 0; i < 6; ++i) d += c.g[i] * c.dir[i];
        c.d_phi_t = -d;
        c.psi_t = psi_mt(c.a_t, c.phi_t, c.phi_0, c.d_phi_0, mu);
        c.d_psi_t = dpsi_mt(c.d_phi_t, c.d_phi_0, mu);
        if (c.open_interval && (c.psi_t <= 0 && c.d_psi_t >= 0)) {
            c.open_interval = 0;
            c.f_l += c.phi_0 - mu * c.d_phi_0 * c.a_l;
            c.g_l += mu * c.d_phi_0;
            c.f_u += c.phi_0 - mu * c.d_phi_0 * c.a_u;
            c.g_u += mu * c.d_phi_0;
        }
        if (c.open_interval) c.interval_converged = update_interval_mt(c.a_l, c.f_l, c.g_l, c.a_u, c.f_u, c.g_u, c.a_t, c.psi_t, c.d_psi_t);
        else                 c.interval_converged = update_interval_mt(c.a_l, c.f_l, c.g_l, c.a_u, c.f_u, c.g_u, c.a_t, c.phi_t, c.d_phi_t);
        c.step_iterations++;
        mt_check = true;
    } else if (c.state == ST_MT_HESS) {
        ctl_unpack_acc(c, acc, false, true);      // computeHessian (NDTM:901-936)
    } else {
        return CTL_DONE;
    }

    if (mt_check) {
        if (!c.interval_converged && c.step_iterations < max_step_iterations &&
            !(c.psi_t <= 0 && c.d_phi_t <= -nu * c.d_phi_0)) {
            if (c.open_interval) c.a_t = trial_value_selection_mt(c.a_l, c.f_l, c.g_l, c.a_u, c.f_u, c.g_u, c.a_t, c.psi_t, c.d_psi_t);
            else                 c.a_t = trial_value_selection_mt(c.a_l, c.f_l, c.g_l, c.a_u, c.f_u, c.g_u, c.a_t, c.phi_t, c.d_phi_t);
            c.a_t = (c.a_t < c.step_max) ? c.a_t : c.step_max;
            c.a_t = (c.a_t > c.step_min) ? c.a_t : c.step_min;
            for (int i = 0; i < 6; ++i) c.x_t[i] = c.p[i] + c.dir[i] * c.a_t;
            ctl_request(c, c.x_t, 0, ST_MT_TRIAL);
            return CTL_PASS;
        }
        if (c.step_iterations) {
            // same pose, Hessian only
            c.hess = 1;
            c.state = ST_MT_HESS;
            c.need_trig = 0;
            return CTL_PASS;
        }
    }
    return ctl_finish_iter(c, k);
}

// NDTM:353-365 after the solve, then the computeStepLengthMT prologue NDTM:656-698
// |delta| as ctl_post_newton forms it (the controller warp evaluates it on every lane and divides on six of them)
B2_HD double ctl_delta_norm(const double delta[6]) {
    double n2 = 0.0;
    for (int i = 0; i < 6; ++i) n2 += delta[i] * delta[i];
    return sqrt(n2);
}
// have_dir: c.dir already holds delta / |delta| (written by six lanes of the controller warp)
B2_HD_NOINLINE inline int ctl_post_newton(Ctl &c, const NdtConst &k, const double delta[6], bool have_dir = false) {
    const double mu = 1.e-4;
    double nrm = ctl_delta_norm(delta);
    if (nrm == 0 || nrm != nrm) {
        c.trans_probability = c.score / c.npoints;
        c.converged = (nrm == nrm) ? 1 : 0;
        c.state = ST_DONE;
        return CTL_DONE;
    }
    if (!have_dir)
        for (int i = 0; i < 6; ++i) c.dir[i] = delta[i] / nrm;
    c.phi_0 = -c.score;
    double d = 0.0;
    for (int i = 0; i < 6; ++i) d += c.g[i] * c.dir[i];
    c.d_phi_0 = -d;
    if (c.d_phi_0 >= 0) {
        if (c.d_phi_0 == 0) { c.a_t = 0.0; return ctl_finish_iter(c, k); }
        c.d_phi_0 *= -1;
        for (int i = 0; i < 6; ++i) c.dir[i] *= -1;
    }
    c.step_max = k.step_size;
    c.step_min = k.trans_eps / 2;
    c.step_iterations = 0;
    c.a_l = 0; c.a_u = 0;
    c.f_l = psi_mt(c.a_l, c.phi_0, c.phi_0, c.d_phi_0, mu);
    c.g_l = dpsi_mt(c.d_phi_0, c.d_phi_0, mu);
    c.f_u = psi_mt(c.a_u, c.phi_0, c.phi_0, c.d_phi_0, mu);
    c.g_u = dpsi_mt(c.d_phi_0, c.d_phi_0, mu);
    c.interval_converged = k.pcl17_compat ? ((c.step_max - c.step_min) > 0) : ((c.step_max - c.step_min) < 0);
    c.open_interval = 1;
    c.a_t = nrm;
    c.a_t = fmin(c.a_t, c.step_max);
    c.a_t = fmax(c.a_t, c.step_min);
    for (int i = 0; i < 6; ++i) c.x_t[i] = c.p[i] + c.dir[i] * c.a_t;
    ctl_request(c, c.x_t, 1, ST_MT_FIRST);
    return CTL_PASS;
}

// Consume the finished pass and decide what happens next.  Returns 1 when another pass is requested
// (c.T / c.ang / c.hess describe it), 0 when the alignment is finished.
B2_HD_NOINLINE inline int ctl_step(Ctl &c, const NdtConst &k, const double *acc) {
    int code = ctl_pre(c, k, acc);
    for (int guard = 0; guard < 100000 && code == CTL_NEWTON; ++guard) {
        double neg_g[6], delta[6];
        for (int i = 0; i < 6; ++i) neg_g[i] = -c.g[i];
        newton_solve6(c.H, neg_g, delta, k.force_svd);
        code = ctl_post_newton(c, k, delta);
    }
    if (code == CTL_NEWTON) { c.state = ST_DONE; code = CTL_DONE; }
    return code == CTL_PASS;
}

}  // namespace b2
