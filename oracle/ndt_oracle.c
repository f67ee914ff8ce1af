/*
 * ndt_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see ndt_oracle.h for scope + parity status).
 *
 * Reference citations use paths relative to /root/reference/lidar_localization/:
 *   NDTM/ = src/models/registration/ndt_registration_manual/
 *   EIG/  = third_party/eigen3/Eigen/src/
 * Arithmetic notes: compile WITHOUT -ffast-math and WITHOUT FMA contraction (-ffp-contract=off),
 * like the reference's x86-64 -O3 build (CMakeLists.txt:4-9), so float voxel indices and float
 * point transforms are reproduced operation by operation.
 */
#include "ndt_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                              */
/* ------------------------------------------------------------------------------------------ */
static inline const float *pt_xyz(orc_cloud c, size_t i) {
    return (const float *)((const char *)c.data + i * c.stride);
}
static inline float pt_i(orc_cloud c, size_t i) {
    return *(const float *)((const char *)c.data + i * c.stride + c.ioff);
}
static inline int finite3(const float *p) { return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); }

/* float32 sin/cos used when the reference builds a float pose matrix
 * (Eigen::AngleAxis<float>::toRotationMatrix -> std::sin(float)).  libm sinf is within 1 ulp but
 * not specified bit-for-bit across libm versions, so both this oracle and the CUDA path use the
 * correctly-rounded-in-practice definition float(sin(double(x))).  */
static int g_f32_trig_libm = 0; /* 1: use libm sinf/cosf exactly as the reference build would */
void orc_set_f32_trig_libm(int on) { g_f32_trig_libm = on; }
static inline float sin_f32(float x) { return g_f32_trig_libm ? sinf(x) : (float)sin((double)x); }
static inline float cos_f32(float x) { return g_f32_trig_libm ? cosf(x) : (float)cos((double)x); }
static inline float atan2_f32(float y, float x) { return g_f32_trig_libm ? atan2f(y, x) : (float)atan2((double)y, (double)x); }

/* stable LSD radix sort of (key,val) pairs by 32-bit key */
static void radix_sort_pairs(uint32_t *key, uint32_t *val, size_t n) {
    if (n < 2) return;
    uint32_t *k2 = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *v2 = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *ka = key, *va = val, *kb = k2, *vb = v2;
    for (int pass = 0; pass < 4; ++pass) {
        size_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        int sh = pass * 8;
        for (size_t i = 0; i < n; ++i) cnt[((ka[i] >> sh) & 255u) + 1]++;
        if (cnt[1] == n && pass > 0) { /* all zero digit: skip */
            int all0 = 1;
            for (size_t i = 0; i < n && all0; ++i) all0 = ((ka[i] >> sh) == 0);
            if (all0) break;
        }
        for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
        for (size_t i = 0; i < n; ++i) {
            size_t d = cnt[(ka[i] >> sh) & 255u]++;
            kb[d] = ka[i];
            vb[d] = va[i];
        }
        uint32_t *t;
        t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    if (ka != key) {
        memcpy(key, ka, n * sizeof(uint32_t));
        memcpy(val, va, n * sizeof(uint32_t));
    }
    free(k2);
    free(v2);
}

/* ------------------------------------------------------------------------------------------ */
/* voxel layout: pcl::VoxelGrid::applyFilter / VoxelGridCovariance::applyFilter prologue       */
/* (in-tree counterpart NDTM/VoxelGrid.cpp:360-430 uses x/leaf)       */
/* ------------------------------------------------------------------------------------------ */
int orc_vox_layout_compute(orc_cloud c, float lx, float ly, float lz, orc_vox_layout *L) {
    memset(L, 0, sizeof(*L));
    const float leaf[3] = {lx, ly, lz};
    for (int a = 0; a < 3; ++a) L->inv[a] = 1.0f / leaf[a];
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    size_t nf = 0;
    for (size_t i = 0; i < c.n; ++i) {
        const float *p = pt_xyz(c, i);
        if (!finite3(p)) continue; /* getMinMax3D skips non-finite points of non-dense clouds */
        for (int a = 0; a < 3; ++a) {
            if (p[a] < mn[a]) mn[a] = p[a];
            if (p[a] > mx[a]) mx[a] = p[a];
        }
        ++nf;
    }
    L->n_finite = nf;
    if (nf == 0) return 0;
    for (int a = 0; a < 3; ++a) { L->min_p[a] = mn[a]; L->max_p[a] = mx[a]; }
    /* int64 guard: static_cast<int64_t>((max-min)*inv)+1, product vs INT32_MAX */
    int64_t d[3];
    for (int a = 0; a < 3; ++a) d[a] = (int64_t)((mx[a] - mn[a]) * L->inv[a]) + 1;
    if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) return 0;
    for (int a = 0; a < 3; ++a) {
        L->min_b[a] = (int32_t)floorf(mn[a] * L->inv[a]);
        L->max_b[a] = (int32_t)floorf(mx[a] * L->inv[a]);
        L->div_b[a] = L->max_b[a] - L->min_b[a] + 1;
    }
    L->divb_mul[0] = 1;
    L->divb_mul[1] = L->div_b[0];
    L->divb_mul[2] = L->div_b[0] * L->div_b[1];
    L->ok = 1;
    return 1;
}

int32_t orc_vox_index(const orc_vox_layout *L, float x, float y, float z) {
    /* ijk = (int)(floor(x * inv) - (float)min_b)   -- float multiply, no FMA */
    int32_t i0 = (int32_t)(floorf(x * L->inv[0]) - (float)L->min_b[0]);
    int32_t i1 = (int32_t)(floorf(y * L->inv[1]) - (float)L->min_b[1]);
    int32_t i2 = (int32_t)(floorf(z * L->inv[2]) - (float)L->min_b[2]);
    return i0 * L->divb_mul[0] + i1 * L->divb_mul[1] + i2 * L->divb_mul[2];
}

/* sorted (idx, point) list of the finite points; returns count */
static size_t build_sorted_index(orc_cloud c, const orc_vox_layout *L, uint32_t **keys_out,
                                 uint32_t **ids_out) {
    uint32_t *keys = (uint32_t *)malloc((c.n ? c.n : 1) * sizeof(uint32_t));
    uint32_t *ids = (uint32_t *)malloc((c.n ? c.n : 1) * sizeof(uint32_t));
    size_t m = 0;
    for (size_t i = 0; i < c.n; ++i) {
        const float *p = pt_xyz(c, i);
        if (!finite3(p)) continue;
        keys[m] = (uint32_t)orc_vox_index(L, p[0], p[1], p[2]);
        ids[m] = (uint32_t)i;
        ++m;
    }
    radix_sort_pairs(keys, ids, m);
    *keys_out = keys;
    *ids_out = ids;
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* pcl::VoxelGrid<PointXYZI>::applyFilter  (voxel_filter.cpp:36-41 -> PCL; Appendix A.3)       */
/* ------------------------------------------------------------------------------------------ */
size_t orc_voxel_filter(orc_cloud in, float lx, float ly, float lz, float *out, int32_t *out_idx,
                        int32_t *out_cnt, int *overflow) {
    orc_vox_layout L;
    if (overflow) *overflow = 0;
    if (in.n == 0) return 0;
    if (!orc_vox_layout_compute(in, lx, ly, lz, &L)) {
        if (L.n_finite == 0) return 0;
        /* "Leaf size is too small ... output = *input" */
        if (overflow) *overflow = 1;
        for (size_t i = 0; i < in.n; ++i) {
            const float *p = pt_xyz(in, i);
            out[4 * i + 0] = p[0]; out[4 * i + 1] = p[1]; out[4 * i + 2] = p[2];
            out[4 * i + 3] = pt_i(in, i);
            if (out_idx) out_idx[i] = -1;
            if (out_cnt) out_cnt[i] = 1;
        }
        return in.n;
    }
    uint32_t *keys, *ids;
    size_t m = build_sorted_index(in, &L, &keys, &ids);
    size_t M = 0, s = 0;
    while (s < m) {
        size_t e = s + 1;
        while (e < m && keys[e] == keys[s]) ++e;
        /* centroid: float accumulation of (x,y,z,intensity), then /= float(count) */
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (size_t k = s; k < e; ++k) {
            const float *p = pt_xyz(in, ids[k]);
            acc[0] += p[0]; acc[1] += p[1]; acc[2] += p[2];
            acc[3] += pt_i(in, ids[k]);
        }
        float cnt = (float)(e - s);
        for (int a = 0; a < 4; ++a) out[4 * M + a] = acc[a] / cnt;
        if (out_idx) out_idx[M] = (int32_t)keys[s];
        if (out_cnt) out_cnt[M] = (int32_t)(e - s);
        ++M;
        s = e;
    }
    free(keys);
    free(ids);
    return M;
}

/* ------------------------------------------------------------------------------------------ */
/* Eigen numerics restated                                                                    */
/* ------------------------------------------------------------------------------------------ */

/* EIG/LU/InverseImpl.h:156-170: cofactor inverse of a 3x3 (row-major in/out here) */
static inline double cof3(const double *m, int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
void orc_inverse3(const double A[9], double out[9]) {
    double c0[3] = {cof3(A, 0, 0), cof3(A, 1, 0), cof3(A, 2, 0)};
    double det = (c0[0] * A[0] + c0[1] * A[3]) + c0[2] * A[6];
    double invdet = 1.0 / det;
    out[0] = c0[0] * invdet; out[1] = c0[1] * invdet; out[2] = c0[2] * invdet;
    out[3] = cof3(A, 0, 1) * invdet; out[4] = cof3(A, 1, 1) * invdet; out[5] = cof3(A, 2, 1) * invdet;
    out[6] = cof3(A, 0, 2) * invdet; out[7] = cof3(A, 1, 2) * invdet; out[8] = cof3(A, 2, 2) * invdet;
}

/* Symmetric 3x3 eigen decomposition, ascending eigenvalues (what PCL obtains from
 * Eigen::SelfAdjointEigenSolver<Matrix3d>, EIG/Eigenvalues/SelfAdjointEigenSolver.h; the in-tree
 * copy uses a closed-form solver NDTMi/SymmetricEigenSolver.h:55-136).  The decomposition is unique
 * up to eigenvector sign and the only consumer (V diag V^-1, inverse) is sign invariant, so a
 * cyclic Jacobi iteration to full double precision is used; it is checked against the real Eigen
 * solver in tests (oracle/_ref). */
void orc_eig3_sym(const double Ain[9], double evals[3], double V[9]) {
    double A[9];
    memcpy(A, Ain, sizeof(A));
    /* symmetrise from the lower triangle like Eigen (reads triangularView<Lower>) */
    A[1] = A[3]; A[2] = A[6]; A[5] = A[7];
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
        double dia = fabs(A[0]) + fabs(A[4]) + fabs(A[8]);
        if (off == 0.0 || off <= 1e-300 || off < 1e-22 * dia) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double apq = A[p * 3 + q];
                if (apq == 0.0) continue;
                double app = A[p * 3 + p], aqq = A[q * 3 + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                /* A <- J^T A J */
                for (int k = 0; k < 3; ++k) {
                    double akp = A[k * 3 + p], akq = A[k * 3 + q];
                    A[k * 3 + p] = c * akp - s * akq;
                    A[k * 3 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    double apk = A[p * 3 + k], aqk = A[q * 3 + k];
                    A[p * 3 + k] = c * apk - s * aqk;
                    A[q * 3 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                    V[k * 3 + p] = c * vkp - s * vkq;
                    V[k * 3 + q] = s * vkp + c * vkq;
                }
            }
    }
    evals[0] = A[0]; evals[1] = A[4]; evals[2] = A[8];
    /* ascending selection sort with column swaps (SelfAdjointEigenSolver sorts ascending) */
    for (int i = 0; i < 2; ++i) {
        int k = i;
        for (int j = i + 1; j < 3; ++j) if (evals[j] < evals[k]) k = j;
        if (k != i) {
            double t = evals[i]; evals[i] = evals[k]; evals[k] = t;
            for (int r = 0; r < 3; ++r) { t = V[r * 3 + i]; V[r * 3 + i] = V[r * 3 + k]; V[r * 3 + k] = t; }
        }
    }
}

/* EIG/Jacobi/Jacobi.h:83-113 JacobiRotation::makeJacobi(x, y, z) */
static void make_jacobi(double x, double y, double z, double *c, double *s) {
    if (y == 0.0) { *c = 1.0; *s = 0.0; return; }
    double tau = (x - z) / (2.0 * fabs(y));
    double w = sqrt(tau * tau + 1.0);
    double t = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
    double sign_t = t > 0.0 ? 1.0 : -1.0;
    double n = 1.0 / sqrt(t * t + 1.0);
    *s = -sign_t * (y / fabs(y)) * fabs(t) * n;
    *c = n;
}
/* EIG/Jacobi/Jacobi.h:301-: x <- c x + s y ; y <- -s x + c y */
static void rot_plane(double *x, int incx, double *y, int incy, int n, double c, double s) {
    if (c == 1.0 && s == 0.0) return;
    for (int i = 0; i < n; ++i) {
        double xi = x[i * incx], yi = y[i * incy];
        x[i * incx] = c * xi + s * yi;
        y[i * incy] = -s * xi + c * yi;
    }
}

/* Eigen::JacobiSVD<Matrix<double,6,6>>(H, ComputeFullU|ComputeFullV).solve(b)
 * EIG/SVD/JacobiSVD.h:675-785 (two-sided Jacobi), EIG/SVD/SVDBase.h:130-139 (rank),
 * :260-272 (_solve_impl).  Column-major 6x6.  Call site: NDTM/NormalDistributionsTransform.cpp:353-355. */
int orc_jacobi_svd_solve6(const double Hin[36], const double b[6], double x[6], double sv_out[6]) {
    enum { N = 6 };
    double W[36], U[36], V[36], sv[6];
#define M_(A, r, c) (A)[(c) * N + (r)]
    const double precision = 2.0 * DBL_EPSILON;
    const double considerAsZero = 2.0 * 4.9406564584124654e-324; /* 2*denorm_min */
    double scale = 0.0;
    for (int i = 0; i < 36; ++i) { double a = fabs(Hin[i]); if (a > scale) scale = a; }
    /* maxCoeff with NaN: comparisons false -> keep going; mirror by letting NaN fall through */
    if (scale == 0.0) scale = 1.0;
    for (int i = 0; i < 36; ++i) { W[i] = Hin[i] / scale; U[i] = V[i] = (i % 7 == 0) ? 1.0 : 0.0; }
    int finished = 0, guard = 0;
    while (!finished && guard++ < 1000) {
        finished = 1;
        for (int p = 1; p < N; ++p)
            for (int q = 0; q < p; ++q) {
                double mpp = fabs(M_(W, p, p)), mqq = fabs(M_(W, q, q));
                double thr = precision * (mpp > mqq ? mpp : mqq);
                if (!(thr > considerAsZero)) thr = considerAsZero;
                if (fabs(M_(W, p, q)) > thr || fabs(M_(W, q, p)) > thr) {
                    finished = 0;
                    /* real_2x2_jacobi_svd (JacobiSVD.h:405-435) */
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
                    /* m.applyOnTheLeft(0,1,rot1): rows 0,1 of the 2x2 */
                    double a00 = r1c * m00 + r1s * m10, a01 = r1c * m01 + r1s * m11;
                    double a10 = -r1s * m00 + r1c * m10, a11 = -r1s * m01 + r1c * m11;
                    (void)a10;
                    double jrc, jrs;
                    make_jacobi(a00, a01, a11, &jrc, &jrs);
                    /* j_left = rot1 * j_right^T ; transpose() = (c, -s);
                     * operator*: c = c1 c2 - s1 s2 ; s = c1 s2 + s1 c2 */
                    double jlc = r1c * jrc - r1s * (-jrs);
                    double jls = r1c * (-jrs) + r1s * jrc;
                    /* workMatrix.applyOnTheLeft(p,q,j_left): rows p,q */
                    rot_plane(&M_(W, p, 0), N, &M_(W, q, 0), N, N, jlc, jls);
                    /* U.applyOnTheRight(p,q,j_left.transpose()) -> apply_rotation(cols, (j^T)^T = j) */
                    rot_plane(&M_(U, 0, p), 1, &M_(U, 0, q), 1, N, jlc, jls);
                    /* workMatrix.applyOnTheRight(p,q,j_right) -> apply_rotation(cols, j_right^T) */
                    rot_plane(&M_(W, 0, p), 1, &M_(W, 0, q), 1, N, jrc, -jrs);
                    rot_plane(&M_(V, 0, p), 1, &M_(V, 0, q), 1, N, jrc, -jrs);
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
    /* rank(): threshold = diagSize*eps */
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
    if (sv_out) memcpy(sv_out, sv, sizeof(sv));
#undef M_
    return rank;
}

/* float helpers for the 3x3 JacobiSVD below (same algorithm as the 6x6 double one) */
static void make_jacobi_f(float x, float y, float z, float *c, float *s) {
    if (y == 0.0f) { *c = 1.0f; *s = 0.0f; return; }
    float tau = (x - z) / (2.0f * fabsf(y));
    float w = sqrtf(tau * tau + 1.0f);
    float t = (tau > 0.0f) ? 1.0f / (tau + w) : 1.0f / (tau - w);
    float sign_t = t > 0.0f ? 1.0f : -1.0f;
    float n = 1.0f / sqrtf(t * t + 1.0f);
    *s = -sign_t * (y / fabsf(y)) * fabsf(t) * n;
    *c = n;
}
static void rot_plane_f(float *x, int incx, float *y, int incy, int n, float c, float s) {
    if (c == 1.0f && s == 0.0f) return;
    for (int i = 0; i < n; ++i) {
        float xi = x[i * incx], yi = y[i * incy];
        x[i * incx] = c * xi + s * yi;
        y[i * incy] = -s * xi + c * yi;
    }
}

/* Transform<float,3,Affine>::rotation() (EIG/Geometry/Transform.h:1057-1073): for an Affine-mode
 * transform Eigen does NOT return the linear block; it runs JacobiSVD<Matrix3f> on it and returns
 * U * V^T (with the determinant sign folded into U's first column).  NDT's initial 6-vector is
 * therefore taken from this float polar factor (NDTM/NormalDistributionsTransform.cpp:331-337).
 * Lin/Rot are column-major 3x3. */
static void rotation_of_affine_f32(const float Lin[9], float Rot[9]) {
    enum { N = 3 };
    float W[9], U[9], V[9];
#define F_(A, r, c) (A)[(c) * N + (r)]
    const float precision = 2.0f * FLT_EPSILON;
    const float considerAsZero = 2.0f * 1.40129846e-45f;
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) { float a = fabsf(Lin[i]); if (a > scale) scale = a; }
    if (scale == 0.0f) scale = 1.0f;
    for (int i = 0; i < 9; ++i) { W[i] = Lin[i] / scale; U[i] = V[i] = (i % 4 == 0) ? 1.0f : 0.0f; }
    int finished = 0, guard = 0;
    while (!finished && guard++ < 1000) {
        finished = 1;
        for (int p = 1; p < N; ++p)
            for (int q = 0; q < p; ++q) {
                float mpp = fabsf(F_(W, p, p)), mqq = fabsf(F_(W, q, q));
                float thr = precision * (mpp > mqq ? mpp : mqq);
                if (!(thr > considerAsZero)) thr = considerAsZero;
                if (fabsf(F_(W, p, q)) > thr || fabsf(F_(W, q, p)) > thr) {
                    finished = 0;
                    float m00 = F_(W, p, p), m01 = F_(W, p, q), m10 = F_(W, q, p), m11 = F_(W, q, q);
                    float r1c, r1s;
                    float t = m00 + m11, d = m10 - m01;
                    if (d == 0.0f) { r1s = 0.0f; r1c = 1.0f; }
                    else {
                        float u = t / d;
                        float tmp = sqrtf(1.0f + u * u);
                        r1s = 1.0f / tmp;
                        r1c = u / tmp;
                    }
                    float a00 = r1c * m00 + r1s * m10, a01 = r1c * m01 + r1s * m11;
                    float a11 = -r1s * m01 + r1c * m11;
                    float jrc, jrs;
                    make_jacobi_f(a00, a01, a11, &jrc, &jrs);
                    float jlc = r1c * jrc - r1s * (-jrs);
                    float jls = r1c * (-jrs) + r1s * jrc;
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
        if (a != 0.0f) { float f = wii / a; for (int r = 0; r < N; ++r) F_(U, r, i) *= f; }
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
    /* x = (U * V^T).determinant() */
    float UVt[9];
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c)
            F_(UVt, r, c) = (F_(U, r, 0) * F_(V, c, 0) + F_(U, r, 1) * F_(V, c, 1)) + F_(U, r, 2) * F_(V, c, 2);
#define DET3H(m, a, b, c) (F_(m, 0, a) * (F_(m, 1, b) * F_(m, 2, c) - F_(m, 1, c) * F_(m, 2, b)))
    float x = DET3H(UVt, 0, 1, 2) - DET3H(UVt, 1, 0, 2) + DET3H(UVt, 2, 0, 1);
#undef DET3H
    for (int r = 0; r < N; ++r) F_(U, r, 0) /= x;
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c)
            F_(Rot, r, c) = (F_(U, r, 0) * F_(V, c, 0) + F_(U, r, 1) * F_(V, c, 1)) + F_(U, r, 2) * F_(V, c, 2);
#undef F_
}

/* EIG/Geometry/EulerAngles.h:36-99, specialised to eulerAngles(0,1,2) applied to
 * Transform<float,3,Affine>::rotation() of the 4x4 (call site
 * NDTM/NormalDistributionsTransform.cpp:331-337).  Tin is column-major 4x4. */
void orc_euler_angles_012_f32(const float Tin[16], float out[3]) {
    float Lin[9], T[16];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Lin[c * 3 + r] = Tin[c * 4 + r];
    float Rot[9];
    rotation_of_affine_f32(Lin, Rot);
    memset(T, 0, sizeof(T));
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) T[c * 4 + r] = Rot[c * 3 + r];
#define R_(r, c) T[(c) * 4 + (r)]
    const float pi = (float)3.141592653589793238462643383279502884197169399375105820974944592307816406;
    float r0 = atan2_f32(R_(1, 2), R_(2, 2));
    float c2 = sqrtf(R_(0, 0) * R_(0, 0) + R_(0, 1) * R_(0, 1));
    float r1;
    if (r0 > 0.0f) { /* odd==0 branch */
        r0 = r0 - pi;
        r1 = atan2_f32(-R_(0, 2), -c2);
    } else {
        r1 = atan2_f32(-R_(0, 2), c2);
    }
    float s1 = sin_f32(r0), c1 = cos_f32(r0);
    float r2 = atan2_f32(s1 * R_(2, 0) - c1 * R_(1, 0), c1 * R_(1, 1) - s1 * R_(2, 1));
    out[0] = -r0; out[1] = -r1; out[2] = -r2;
#undef R_
}

/* (Translation<float,3>(p0,p1,p2) * AngleAxis<float>(p3,UnitX) * AngleAxis<float>(p4,UnitY)
 *  * AngleAxis<float>(p5,UnitZ)).matrix()   NDTM/NormalDistributionsTransform.cpp:370-373,691-694.
 * AngleAxis::toRotationMatrix (EIG/Geometry/AngleAxis.h) gives for a unit axis e_k:
 * diagonal (1-c)*e_k*e_k + c, off-diagonal 0 -/+ s.  Float 3x3 products are accumulated left to
 * right as Eigen's coefficient-based product does. */
static void rot_axis_f32(int axis, float ang, float R[9] /* row-major */) {
    float s = sin_f32(ang), c = cos_f32(ang);
    float one_c = 1.0f - c;
    float e[3] = {0.f, 0.f, 0.f};
    e[axis] = 1.0f;
    float sin_axis[3] = {s * e[0], s * e[1], s * e[2]};
    float cos1_axis[3] = {one_c * e[0], one_c * e[1], one_c * e[2]};
    float tmp;
    tmp = cos1_axis[0] * e[1]; R[0 * 3 + 1] = tmp - sin_axis[2]; R[1 * 3 + 0] = tmp + sin_axis[2];
    tmp = cos1_axis[0] * e[2]; R[0 * 3 + 2] = tmp + sin_axis[1]; R[2 * 3 + 0] = tmp - sin_axis[1];
    tmp = cos1_axis[1] * e[2]; R[1 * 3 + 2] = tmp - sin_axis[0]; R[2 * 3 + 1] = tmp + sin_axis[0];
    for (int k = 0; k < 3; ++k) R[k * 3 + k] = cos1_axis[k] * e[k] + c;
}
static void mat3_mul_f32(const float A[9], const float B[9], float C[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = (A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j]) + A[i * 3 + 2] * B[2 * 3 + j];
}
void orc_pose_to_matrix_f32(const double p[6], float T[16]) {
    float Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
    rot_axis_f32(0, (float)p[3], Rx);
    rot_axis_f32(1, (float)p[4], Ry);
    rot_axis_f32(2, (float)p[5], Rz);
    mat3_mul_f32(Rx, Ry, Rxy);
    mat3_mul_f32(Rxy, Rz, R);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[c * 4 + r] = R[r * 3 + c];
    T[3] = T[7] = T[11] = 0.0f;
    T[12] = (float)p[0]; T[13] = (float)p[1]; T[14] = (float)p[2]; T[15] = 1.0f;
}

/* pcl::transformPointCloud (dense branch): x' = m00 x + m01 y + m02 z + m03, float, left to right */
void orc_transform_point_f32(const float T[16], float x, float y, float z, float out[3]) {
    out[0] = ((T[0] * x + T[4] * y) + T[8] * z) + T[12];
    out[1] = ((T[1] * x + T[5] * y) + T[9] * z) + T[13];
    out[2] = ((T[2] * x + T[6] * y) + T[10] * z) + T[14];
}

/* NDTM/NormalDistributionsTransform.cpp:315-321 */
void orc_gauss_constants(double outlier_ratio, float resolution, double *d1, double *d2) {
    double c1 = 10.0 * (1.0 - outlier_ratio);
    double c2 = outlier_ratio / pow((double)resolution, 3);
    double d3 = -log(c2);
    *d1 = -log(c1 + c2) - d3;
    *d2 = -2.0 * log((-log(c1 * exp(-0.5) + c2) - d3) / *d1);
}

void orc_params_default(orc_params *p) {
    p->res = 1.0f; p->step_size = 0.1; p->trans_eps = 0.01; p->outlier_ratio = 0.55;
    p->max_iter = 30; p->min_pts = 6; p->eig_mult = 0.01; p->pcl17_compat = 1; p->intree_compat = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* pcl::VoxelGridCovariance::applyFilter (Appendix A.2; NDTM/VoxelGrid.cpp:244-323)            */
/* ------------------------------------------------------------------------------------------ */
struct orc_grid {
    orc_vox_layout L;
    float    res;
    int      min_pts;
    size_t   n_leaves;
    orc_leaf *leaves;
    /* open addressing hash: voxel idx -> leaf slot */
    size_t   hcap;
    int32_t *hkey;
    int32_t *hval;
};

static inline size_t hash_u32(uint32_t k) {
    k ^= k >> 16; k *= 0x7feb352dU; k ^= k >> 15; k *= 0x846ca68bU; k ^= k >> 16;
    return k;
}
static int32_t grid_lookup(const orc_grid *g, int32_t idx) {
    size_t h = hash_u32((uint32_t)idx) & (g->hcap - 1);
    while (g->hkey[h] != -1) {
        if (g->hkey[h] == idx) return g->hval[h];
        h = (h + 1) & (g->hcap - 1);
    }
    return -1;
}

orc_grid *orc_grid_build(orc_cloud tgt, float res, int min_pts, double eig_mult) {
    orc_grid *g = (orc_grid *)calloc(1, sizeof(orc_grid));
    g->res = res;
    g->min_pts = min_pts;
    g->hcap = 16;
    if (tgt.n == 0 || !orc_vox_layout_compute(tgt, res, res, res, &g->L)) {
        /* empty input or overflow guard: PCL warns and leaves no cells */
        g->hkey = (int32_t *)malloc(g->hcap * sizeof(int32_t));
        g->hval = (int32_t *)malloc(g->hcap * sizeof(int32_t));
        memset(g->hkey, 0xff, g->hcap * sizeof(int32_t));
        return g;
    }
    uint32_t *keys, *ids;
    size_t m = build_sorted_index(tgt, &g->L, &keys, &ids);
    size_t nl = 0;
    for (size_t s = 0; s < m;) { size_t e = s + 1; while (e < m && keys[e] == keys[s]) ++e; ++nl; s = e; }
    g->n_leaves = nl;
    g->leaves = (orc_leaf *)calloc(nl ? nl : 1, sizeof(orc_leaf));
    while (g->hcap < 2 * nl + 16) g->hcap <<= 1;
    g->hkey = (int32_t *)malloc(g->hcap * sizeof(int32_t));
    g->hval = (int32_t *)malloc(g->hcap * sizeof(int32_t));
    memset(g->hkey, 0xff, g->hcap * sizeof(int32_t));

    size_t li = 0;
    for (size_t s = 0; s < m;) {
        size_t e = s + 1;
        while (e < m && keys[e] == keys[s]) ++e;
        orc_leaf *lf = &g->leaves[li];
        lf->idx = (int32_t)keys[s];
        int n = (int)(e - s);
        lf->n_raw = n;
        lf->nr_points = n;
        lf->weight = 1.0;
        /* first pass of applyFilter: per point, in input order (std::map leaf accumulates as the
         * cloud is walked): mean_ += p (double), cov_ += p p^T (double, cov_ starts at Identity),
         * centroid += (x,y,z,intensity) (float) */
        double sum[3] = {0, 0, 0};
        double cov[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        float cen[4] = {0.f, 0.f, 0.f, 0.f};
        for (size_t k = s; k < e; ++k) {
            const float *p = pt_xyz(tgt, ids[k]);
            double d[3] = {(double)p[0], (double)p[1], (double)p[2]};
            for (int a = 0; a < 3; ++a) sum[a] += d[a];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) cov[a * 3 + b] += d[a] * d[b];
            cen[0] += p[0]; cen[1] += p[1]; cen[2] += p[2];
            cen[3] += pt_i(tgt, ids[k]);
        }
        /* second pass */
        for (int a = 0; a < 4; ++a) lf->centroid[a] = cen[a] / (float)n;
        double mean[3];
        for (int a = 0; a < 3; ++a) mean[a] = sum[a] / (double)n;
        memcpy(lf->mean, mean, sizeof(mean));
        if (n >= min_pts) {
            lf->in_tree = 1;
            /* cov = (cov - 2*(pt_sum*mean^T))/n + mean*mean^T ; cov *= (n-1.0)/n */
            double nd = (double)n;
            double f = (nd - 1.0) / nd;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                    double t = sum[a] * mean[b];
                    double mm = mean[a] * mean[b];
                    double v = (cov[a * 3 + b] - 2.0 * t) / nd + mm;
                    cov[a * 3 + b] = v * f;
                }
            double ev[3], V[9];
            orc_eig3_sym(cov, ev, V);
            if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
                /* nr_points = -1; leaf stays in the kd-tree with icov_ = 0 (ctor default) */
                lf->nr_points = -1;
                memcpy(lf->cov, cov, sizeof(cov));
                memcpy(lf->evals, ev, sizeof(ev));
            } else {
                double mn = eig_mult * ev[2];
                if (ev[0] < mn) {
                    ev[0] = mn;
                    if (ev[1] < mn) ev[1] = mn;
                    /* cov = evecs * diag * evecs.inverse() */
                    double Vi[9], VD[9];
                    orc_inverse3(V, Vi);
                    for (int a = 0; a < 3; ++a)
                        for (int b = 0; b < 3; ++b) VD[a * 3 + b] = V[a * 3 + b] * ev[b];
                    for (int a = 0; a < 3; ++a)
                        for (int b = 0; b < 3; ++b)
                            cov[a * 3 + b] = (VD[a * 3 + 0] * Vi[0 * 3 + b] + VD[a * 3 + 1] * Vi[1 * 3 + b]) + VD[a * 3 + 2] * Vi[2 * 3 + b];
                }
                memcpy(lf->cov, cov, sizeof(cov));
                memcpy(lf->evals, ev, sizeof(ev));
                orc_inverse3(cov, lf->icov);
                double mxc = -DBL_MAX, mnc = DBL_MAX;
                for (int a = 0; a < 9; ++a) { if (lf->icov[a] > mxc) mxc = lf->icov[a]; if (lf->icov[a] < mnc) mnc = lf->icov[a]; }
                if (mxc == (double)INFINITY || mnc == -(double)INFINITY) lf->nr_points = -1;
            }
        }
        size_t h = hash_u32((uint32_t)lf->idx) & (g->hcap - 1);
        while (g->hkey[h] != -1) h = (h + 1) & (g->hcap - 1);
        g->hkey[h] = lf->idx;
        g->hval[h] = (int32_t)li;
        ++li;
        s = e;
    }
    free(keys);
    free(ids);
    return g;
}

static int cmp_leaf_idx(const void *a, const void *b) {
    const orc_leaf *x = (const orc_leaf *)a, *y = (const orc_leaf *)b;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

orc_grid *orc_grid_from_leaves(size_t n, const int32_t *ijk, const double *mean3, const double *icov9,
                               const double *weight, float res) {
    orc_grid *g = (orc_grid *)calloc(1, sizeof(orc_grid));
    g->res = res;
    g->min_pts = 1;
    g->hcap = 16;
    while (g->hcap < 2 * n + 16) g->hcap <<= 1;
    g->hkey = (int32_t *)malloc(g->hcap * sizeof(int32_t));
    g->hval = (int32_t *)malloc(g->hcap * sizeof(int32_t));
    memset(g->hkey, 0xff, g->hcap * sizeof(int32_t));
    g->n_leaves = n;
    g->leaves = (orc_leaf *)calloc(n ? n : 1, sizeof(orc_leaf));
    if (n == 0) return g;
    int32_t mn[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, mx[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) { if (ijk[3 * i + a] < mn[a]) mn[a] = ijk[3 * i + a]; if (ijk[3 * i + a] > mx[a]) mx[a] = ijk[3 * i + a]; }
    g->L.ok = 1;
    g->L.n_finite = n;
    for (int a = 0; a < 3; ++a) { g->L.inv[a] = 1.0f / res; g->L.min_b[a] = mn[a]; g->L.max_b[a] = mx[a]; g->L.div_b[a] = mx[a] - mn[a] + 1; }
    g->L.divb_mul[0] = 1; g->L.divb_mul[1] = g->L.div_b[0]; g->L.divb_mul[2] = g->L.div_b[0] * g->L.div_b[1];
    for (size_t i = 0; i < n; ++i) {
        orc_leaf *lf = &g->leaves[i];
        lf->idx = (ijk[3 * i] - mn[0]) + (ijk[3 * i + 1] - mn[1]) * g->L.divb_mul[1] + (ijk[3 * i + 2] - mn[2]) * g->L.divb_mul[2];
        lf->n_raw = lf->nr_points = 6; lf->in_tree = 1;
        for (int a = 0; a < 3; ++a) { lf->mean[a] = mean3[3 * i + a]; lf->centroid[a] = (float)mean3[3 * i + a]; }
        memcpy(lf->icov, &icov9[9 * i], 9 * sizeof(double));
        lf->weight = weight ? weight[i] : 1.0;
    }
    qsort(g->leaves, n, sizeof(orc_leaf), cmp_leaf_idx);
    for (size_t i = 0; i < n; ++i) {
        size_t h = hash_u32((uint32_t)g->leaves[i].idx) & (g->hcap - 1);
        while (g->hkey[h] != -1) h = (h + 1) & (g->hcap - 1);
        g->hkey[h] = g->leaves[i].idx;
        g->hval[h] = (int32_t)i;
    }
    return g;
}

void orc_grid_free(orc_grid *g) {
    if (!g) return;
    free(g->leaves); free(g->hkey); free(g->hval); free(g);
}
size_t orc_grid_num_leaves(const orc_grid *g) { return g->n_leaves; }
const orc_leaf *orc_grid_leaves(const orc_grid *g) { return g->leaves; }
const orc_vox_layout *orc_grid_layout(const orc_grid *g) { return &g->L; }

/* VoxelGridCovariance::radiusSearch -> KdTreeFLANN::radiusSearch on the centroid cloud:
 * exact search, flann::L2_Simple<float> distance ((dx*dx + dy*dy) + dz*dz in float), accepted when
 * dist < (float)(radius*radius), sorted by distance.  In-tree counterpart scans the index cube
 * with double distances: NDTM/VoxelGrid.cpp:432-480.  Candidate cells are found by scanning the
 * index cube around the query with a safety margin (a float centroid may round onto a cell
 * boundary), which yields exactly the kd-tree's result set. */
int orc_grid_radius_search(const orc_grid *g, float qx, float qy, float qz, double radius,
                           int32_t *slots, float *d2out, int cap) {
    if (!g->L.ok || g->n_leaves == 0) return 0;
    const float q[3] = {qx, qy, qz};
    const float r2 = (float)(radius * radius);
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        double leaf = (double)g->res;
        double m = 0.25 * leaf; /* a float centroid can lie slightly outside its own cell */
        double l = floor(((double)q[a] - radius - m) / leaf) - (double)g->L.min_b[a];
        double h = floor(((double)q[a] + radius + m) / leaf) - (double)g->L.min_b[a];
        if (l < 0) l = 0;
        if (h > g->L.div_b[a] - 1) h = g->L.div_b[a] - 1;
        if (h < l) return 0;
        lo[a] = (int)l; hi[a] = (int)h;
    }
    int cnt = 0;
    for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) {
                int32_t idx = i * g->L.divb_mul[0] + j * g->L.divb_mul[1] + k * g->L.divb_mul[2];
                int32_t s = grid_lookup(g, idx);
                if (s < 0) continue;
                const orc_leaf *lf = &g->leaves[s];
                if (!lf->in_tree) continue;
                float dx = q[0] - lf->centroid[0], dy = q[1] - lf->centroid[1], dz = q[2] - lf->centroid[2];
                float d2 = (dx * dx + dy * dy) + dz * dz;
                if (d2 < r2) {
                    /* insertion sort by (d2, slot) */
                    int pos = cnt < cap ? cnt : cap - 1;
                    if (cnt >= cap && !(d2 < d2out[cap - 1])) { ++cnt; continue; }
                    while (pos > 0 && (d2out[pos - 1] > d2 || (d2out[pos - 1] == d2 && slots[pos - 1] > s))) {
                        d2out[pos] = d2out[pos - 1]; slots[pos] = slots[pos - 1]; --pos;
                    }
                    d2out[pos] = d2; slots[pos] = s;
                    ++cnt;
                }
            }
    return cnt < cap ? cnt : cap;
}

/* The in-tree copy's VoxelGrid::radiusSearch (NDTM/VoxelGrid.cpp:432-480): scan the index cube
 * floor((t +- radius)/leaf) (float arithmetic) clamped to the occupied range, accept when the DOUBLE
 * distance to the double centroid is < radius; visiting order x, y, z.  Only used with intree_compat. */
static int grid_radius_search_intree(const orc_grid *g, float tx, float ty, float tz, float radius, int32_t *slots, int cap) {
    if (!g->L.ok || g->n_leaves == 0) return 0;
    const float t[3] = {tx, ty, tz};
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        int mx = (int)floor((t[a] + radius) / g->res);
        int mn = (int)floor((t[a] - radius) / g->res);
        if (mx > g->L.max_b[a]) mx = g->L.max_b[a];
        if (mn < g->L.min_b[a]) mn = g->L.min_b[a];
        lo[a] = mn; hi[a] = mx;
    }
    int cnt = 0;
    for (int i = lo[0]; i <= hi[0]; ++i)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int k = lo[2]; k <= hi[2]; ++k) {
                int32_t idx = (i - g->L.min_b[0]) + (j - g->L.min_b[1]) * g->L.divb_mul[1] + (k - g->L.min_b[2]) * g->L.divb_mul[2];
                int32_t s = grid_lookup(g, idx);
                if (s < 0 || !g->leaves[s].in_tree) continue;
                const orc_leaf *lf = &g->leaves[s];
                double cx = lf->mean[0] - (double)tx, cy = lf->mean[1] - (double)ty, cz = lf->mean[2] - (double)tz;
                double distance = sqrt(cx * cx + cy * cy + cz * cz);
                if (distance < radius) { if (cnt < cap) slots[cnt] = s; ++cnt; }
            }
    return cnt < cap ? cnt : cap;
}

/* ------------------------------------------------------------------------------------------ */
/* NDT derivatives  (NDTM/NormalDistributionsTransform.cpp:391-645 minus the static weight)    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double j[8][3];   /* j_ang_a .. j_ang_h */
    double h[15][3];  /* h_ang_a2,a3,b2,b3,c2,c3,d1,d2,d3,e1,e2,e3,f1,f2,f3 */
} ang_t;

/* NDTM/NormalDistributionsTransform.cpp:523-645 */
static void angle_derivatives(const double p[6], ang_t *A) {
    double cx, cy, cz, sx, sy, sz;
    if (fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = cos(p[3]); sx = sin(p[3]); }
    if (fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = cos(p[4]); sy = sin(p[4]); }
    if (fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = cos(p[5]); sz = sin(p[5]); }
    double (*j)[3] = A->j;
    double (*h)[3] = A->h;
    j[0][0] = -sx * sz + cx * sy * cz; j[0][1] = -sx * cz - cx * sy * sz; j[0][2] = -cx * cy;
    j[1][0] = cx * sz + sx * sy * cz;  j[1][1] = cx * cz - sx * sy * sz;  j[1][2] = -sx * cy;
    j[2][0] = -sy * cz;                j[2][1] = sy * sz;                 j[2][2] = cy;
    j[3][0] = sx * cy * cz;            j[3][1] = -sx * cy * sz;           j[3][2] = sx * sy;
    j[4][0] = -cx * cy * cz;           j[4][1] = cx * cy * sz;            j[4][2] = -cx * sy;
    j[5][0] = -cy * sz;                j[5][1] = -cy * cz;                j[5][2] = 0;
    j[6][0] = cx * cz - sx * sy * sz;  j[6][1] = -cx * sz - sx * sy * cz; j[6][2] = 0;
    j[7][0] = sx * cz + cx * sy * sz;  j[7][1] = cx * sy * cz - sx * sz;  j[7][2] = 0;

    h[0][0] = -cx * sz - sx * sy * cz; h[0][1] = -cx * cz + sx * sy * sz; h[0][2] = sx * cy;   /* a2 */
    h[1][0] = -sx * sz + cx * sy * cz; h[1][1] = -cx * sy * sz - sx * cz; h[1][2] = -cx * cy;  /* a3 */
    h[2][0] = cx * cy * cz;            h[2][1] = -cx * cy * sz;           h[2][2] = cx * sy;   /* b2 */
    h[3][0] = sx * cy * cz;            h[3][1] = -sx * cy * sz;           h[3][2] = sx * sy;   /* b3 */
    h[4][0] = -sx * cz - cx * sy * sz; h[4][1] = sx * sz - cx * sy * cz;  h[4][2] = 0;         /* c2 */
    h[5][0] = cx * cz - sx * sy * sz;  h[5][1] = -sx * sy * cz - cx * sz; h[5][2] = 0;         /* c3 */
    h[6][0] = -cy * cz;                h[6][1] = cy * sz;                 h[6][2] = sy;        /* d1 */
    h[7][0] = -sx * sy * cz;           h[7][1] = sx * sy * sz;            h[7][2] = sx * cy;   /* d2 */
    h[8][0] = cx * sy * cz;            h[8][1] = -cx * sy * sz;           h[8][2] = -cx * cy;  /* d3 */
    h[9][0] = sy * sz;                 h[9][1] = sy * cz;                 h[9][2] = 0;         /* e1 */
    h[10][0] = -sx * cy * sz;          h[10][1] = -sx * cy * cz;          h[10][2] = 0;        /* e2 */
    h[11][0] = cx * cy * sz;           h[11][1] = cx * cy * cz;           h[11][2] = 0;        /* e3 */
    h[12][0] = -cy * cz;               h[12][1] = cy * sz;                h[12][2] = 0;        /* f1 */
    h[13][0] = -cx * sz - sx * sy * cz;h[13][1] = -cx * cz + sx * sy * sz;h[13][2] = 0;        /* f2 */
    h[14][0] = -sx * sz + cx * sy * cz;h[14][1] = -cx * sy * sz - sx * cz;h[14][2] = 0;        /* f3 */
}

static inline double dot3(const double a[3], const double b[3]) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
static inline void mat3_vec(const double M[9], const double v[3], double o[3]) {
    for (int r = 0; r < 3; ++r) o[r] = (M[r * 3 + 0] * v[0] + M[r * 3 + 1] * v[1]) + M[r * 3 + 2] * v[2];
}

double orc_ndt_derivatives(const orc_grid *g, const orc_params *prm, orc_cloud src,
                           const float *trans_xyz, const double p[6], int compute_hessian,
                           double grad[6], double H[36], long long *pairs_out) {
    double d1, d2;
    orc_gauss_constants(prm->outlier_ratio, prm->res, &d1, &d2);
    ang_t A;
    angle_derivatives(p, &A);
    for (int i = 0; i < 6; ++i) grad[i] = 0.0;
    for (int i = 0; i < 36; ++i) H[i] = 0.0;
    double score = 0.0;
    long long pairs = 0;
    /* point_gradient_ (3x6) and point_hessian_ (18x6): constant parts set once */
    double PG[3][6];
    double PH[18][6];
    memset(PG, 0, sizeof(PG));
    memset(PH, 0, sizeof(PH));
    PG[0][0] = PG[1][1] = PG[2][2] = 1.0;
    int32_t slots[64];
    float d2s[64];
    for (size_t idx = 0; idx < src.n; ++idx) {
        const float *xt = &trans_xyz[3 * idx];
        int nn = (prm->intree_compat & 2)
                     ? grid_radius_search_intree(g, xt[0], xt[1], xt[2], prm->res, slots, 64)
                     : orc_grid_radius_search(g, xt[0], xt[1], xt[2], (double)prm->res, slots, d2s, 64);
        for (int k = 0; k < nn; ++k) {
            const orc_leaf *lf = &g->leaves[slots[k]];
            const float *xo = pt_xyz(src, idx);
            double x[3] = {(double)xo[0], (double)xo[1], (double)xo[2]};
            double xtr[3] = {(double)xt[0] - lf->mean[0], (double)xt[1] - lf->mean[1], (double)xt[2] - lf->mean[2]};
            const double *ci = lf->icov;
            /* computePointDerivatives :448-482 */
            PG[1][3] = dot3(x, A.j[0]); PG[2][3] = dot3(x, A.j[1]);
            PG[0][4] = dot3(x, A.j[2]); PG[1][4] = dot3(x, A.j[3]); PG[2][4] = dot3(x, A.j[4]);
            PG[0][5] = dot3(x, A.j[5]); PG[1][5] = dot3(x, A.j[6]); PG[2][5] = dot3(x, A.j[7]);
            if (compute_hessian) {
                double a[3] = {0, dot3(x, A.h[0]), dot3(x, A.h[1])};
                double b[3] = {0, dot3(x, A.h[2]), dot3(x, A.h[3])};
                double c[3] = {0, dot3(x, A.h[4]), dot3(x, A.h[5])};
                double d[3] = {dot3(x, A.h[6]), dot3(x, A.h[7]), dot3(x, A.h[8])};
                double e[3] = {dot3(x, A.h[9]), dot3(x, A.h[10]), dot3(x, A.h[11])};
                double f[3] = {dot3(x, A.h[12]), dot3(x, A.h[13]), dot3(x, A.h[14])};
                for (int r = 0; r < 3; ++r) {
                    PH[9 + r][3] = a[r];  PH[12 + r][3] = b[r]; PH[15 + r][3] = c[r];
                    PH[9 + r][4] = b[r];  PH[12 + r][4] = d[r]; PH[15 + r][4] = e[r];
                    PH[9 + r][5] = c[r];  PH[12 + r][5] = e[r]; PH[15 + r][5] = f[r];
                }
            }
            /* updateDerivatives :485-520 */
            ++pairs;
            double cx[3];
            mat3_vec(ci, xtr, cx);
            double e_x_cov_x = exp(-d2 * dot3(xtr, cx) / 2);
            double score_inc = -d1 * e_x_cov_x;
            e_x_cov_x = d2 * e_x_cov_x;
            if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) continue;
            e_x_cov_x *= d1;
            for (int i = 0; i < 6; ++i) {
                double pgi[3] = {PG[0][i], PG[1][i], PG[2][i]};
                double cov_dxd_pi[3];
                mat3_vec(ci, pgi, cov_dxd_pi);
                grad[i] += dot3(xtr, cov_dxd_pi) * e_x_cov_x;
                if (compute_hessian) {
                    for (int j = 0; j < 6; ++j) {
                        double pgj[3] = {PG[0][j], PG[1][j], PG[2][j]};
                        double phij[3] = {PH[3 * i + 0][j], PH[3 * i + 1][j], PH[3 * i + 2][j]};
                        double t1[3], t2[3];
                        mat3_vec(ci, pgj, t1);
                        mat3_vec(ci, phij, t2);
                        H[j * 6 + i] += e_x_cov_x * (-d2 * dot3(xtr, cov_dxd_pi) * dot3(xtr, t1) +
                                                     dot3(xtr, t2) + dot3(pgj, cov_dxd_pi));
                    }
                }
            }
            score += (prm->intree_compat & 1) ? lf->weight * score_inc : score_inc;
        }
    }
    if (pairs_out) *pairs_out = pairs;
    return score;
}

/* ------------------------------------------------------------------------------------------ */
/* More-Thuente helpers (NDTM/NormalDistributionsTransform.cpp:69-75, 760-873; "Copied from      */
/* ndt.hpp")                                                                                   */
/* ------------------------------------------------------------------------------------------ */
static double psi_mt(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
static double dpsi_mt(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

static double trial_value_selection_mt(double a_l, double f_l, double g_l, double a_u, double f_u,
                                       double g_u, double a_t, double f_t, double g_t) {
    if (f_t > f_l) { /* case 1 */
        double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
        double w = sqrt(z * z - g_t * g_l);
        double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
        if (fabs(a_c - a_l) < fabs(a_q - a_l)) return a_c;
        return 0.5 * (a_q + a_c);
    } else if (g_t * g_l < 0) { /* case 2 */
        double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
        double w = sqrt(z * z - g_t * g_l);
        double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        if (fabs(a_c - a_t) >= fabs(a_s - a_t)) return a_c;
        return a_s;
    } else if (fabs(g_t) <= fabs(g_l)) { /* case 3 */
        double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
        double w = sqrt(z * z - g_t * g_l);
        double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        double a_t_next = (fabs(a_c - a_t) < fabs(a_s - a_t)) ? a_c : a_s;
        if (a_t > a_l) return fmin(a_t + 0.66 * (a_u - a_t), a_t_next);
        return fmax(a_t + 0.66 * (a_u - a_t), a_t_next);
    } else { /* case 4 */
        double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
        double w = sqrt(z * z - g_t * g_u);
        return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
    }
}

static int update_interval_mt(double *a_l, double *f_l, double *g_l, double *a_u, double *f_u,
                              double *g_u, double a_t, double f_t, double g_t) {
    if (f_t > *f_l) { *a_u = a_t; *f_u = f_t; *g_u = g_t; return 0; }
    else if (g_t * (*a_l - a_t) > 0) { *a_l = a_t; *f_l = f_t; *g_l = g_t; return 0; }
    else if (g_t * (*a_l - a_t) < 0) {
        *a_u = *a_l; *f_u = *f_l; *g_u = *g_l;
        *a_l = a_t; *f_l = f_t; *g_l = g_t;
        return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* align: Registration::align prologue (NDTM/Registration.cpp:76-95) + computeTransformation   */
/* (NDTM/NormalDistributionsTransform.cpp:310-389) + computeStepLengthMT (:648-756)            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const orc_grid *g;
    const orc_params *prm;
    orc_cloud src;
    float *trans;   /* n*3 */
    float final_T[16];
    orc_result *res;
} align_ctx;

static void transform_cloud(align_ctx *c, const float T[16]) {
    for (size_t i = 0; i < c->src.n; ++i) {
        const float *p = pt_xyz(c->src, i);
        if (!finite3(p)) { c->trans[3 * i] = p[0]; c->trans[3 * i + 1] = p[1]; c->trans[3 * i + 2] = p[2]; continue; }
        orc_transform_point_f32(T, p[0], p[1], p[2], &c->trans[3 * i]);
    }
}

static double derivs(align_ctx *c, const double p[6], int hess, double grad[6], double H[36]) {
    long long pr = 0;
    double s = orc_ndt_derivatives(c->g, c->prm, c->src, c->trans, p, hess, grad, H, &pr);
    c->res->passes++;
    c->res->pairs += pr;
    return s;
}

static double step_length_mt(align_ctx *c, const double x[6], double step_dir[6], double step_init,
                             double step_max, double step_min, double *score, double grad[6],
                             double H[36]) {
    double phi_0 = -(*score);
    double d_phi_0 = 0.0;
    for (int i = 0; i < 6; ++i) d_phi_0 += grad[i] * step_dir[i];
    d_phi_0 = -d_phi_0;
    double x_t[6];
    if (d_phi_0 >= 0) {
        if (d_phi_0 == 0) return 0;
        d_phi_0 *= -1;
        for (int i = 0; i < 6; ++i) step_dir[i] *= -1;
    }
    const int max_step_iterations = 10;
    int step_iterations = 0;
    const double mu = 1.e-4, nu = 0.9;
    double a_l = 0, a_u = 0;
    double f_l = psi_mt(a_l, phi_0, phi_0, d_phi_0, mu);
    double g_l = dpsi_mt(d_phi_0, d_phi_0, mu);
    double f_u = psi_mt(a_u, phi_0, phi_0, d_phi_0, mu);
    double g_u = dpsi_mt(d_phi_0, d_phi_0, mu);
    /* PCL 1.7: bool interval_converged = (step_max - step_min) > 0  (:683) */
    int interval_converged = c->prm->pcl17_compat ? ((step_max - step_min) > 0) : ((step_max - step_min) < 0);
    int open_interval = 1;
    double a_t = step_init;
    a_t = fmin(a_t, step_max);
    a_t = fmax(a_t, step_min);
    for (int i = 0; i < 6; ++i) x_t[i] = x[i] + step_dir[i] * a_t;
    orc_pose_to_matrix_f32(x_t, c->final_T);
    transform_cloud(c, c->final_T);
    *score = derivs(c, x_t, 1, grad, H);
    double phi_t = -(*score);
    double d_phi_t = 0.0;
    for (int i = 0; i < 6; ++i) d_phi_t += grad[i] * step_dir[i];
    d_phi_t = -d_phi_t;
    double psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu);
    double d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
    while (!interval_converged && step_iterations < max_step_iterations &&
           !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
        if (open_interval) a_t = trial_value_selection_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
        else               a_t = trial_value_selection_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
        a_t = (a_t < step_max) ? a_t : step_max;
        a_t = (a_t > step_min) ? a_t : step_min;
        for (int i = 0; i < 6; ++i) x_t[i] = x[i] + step_dir[i] * a_t;
        orc_pose_to_matrix_f32(x_t, c->final_T);
        transform_cloud(c, c->final_T);
        *score = derivs(c, x_t, 0, grad, H);
        c->res->mt_trials++;
        /* PCL: phi_t = -score; d_phi_t = -(g . dir)   (in-tree copy accumulates, :726-727; A.4) */
        {
            double gd = 0.0;
            for (int i = 0; i < 6; ++i) gd += grad[i] * step_dir[i];
            if (c->prm->intree_compat & 4) { phi_t -= *score; d_phi_t -= gd; }
            else { phi_t = -(*score); d_phi_t = -gd; }
        }
        psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu);
        d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
        if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
            open_interval = 0;
            f_l += phi_0 - mu * d_phi_0 * a_l;
            g_l += mu * d_phi_0;
            f_u += phi_0 - mu * d_phi_0 * a_u;
            g_u += mu * d_phi_0;
        }
        if (open_interval) interval_converged = update_interval_mt(&a_l, &f_l, &g_l, &a_u, &f_u, &g_u, a_t, psi_t, d_psi_t);
        else               interval_converged = update_interval_mt(&a_l, &f_l, &g_l, &a_u, &f_u, &g_u, a_t, phi_t, d_phi_t);
        step_iterations++;
    }
    if (step_iterations) {
        /* computeHessian(hessian, trans_cloud, x_t) (:901-936): same per-pair Hessian terms */
        double gtmp[6];
        (void)derivs(c, x_t, 1, gtmp, H);
    }
    return a_t;
}

int orc_ndt_align(const orc_grid *g, const orc_params *prm, orc_cloud src, const float guess[16],
                  float pose_out[16], float *result_xyz, orc_result *res, double *trace,
                  int trace_cap) {
    static const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    align_ctx c;
    memset(res, 0, sizeof(*res));
    c.g = g; c.prm = prm; c.src = src; c.res = res;
    c.trans = (float *)malloc((src.n ? src.n : 1) * 3 * sizeof(float));
    memcpy(c.final_T, I4, sizeof(I4));
    /* align(): output = copy of source; data[3] = 1 */
    for (size_t i = 0; i < src.n; ++i) {
        const float *p = pt_xyz(src, i);
        c.trans[3 * i] = p[0]; c.trans[3 * i + 1] = p[1]; c.trans[3 * i + 2] = p[2];
    }
    int nr_iterations = 0, converged = 0;
    if (memcmp(guess, I4, sizeof(I4)) != 0) {
        int differs = 0;
        for (int i = 0; i < 16; ++i) if (guess[i] != I4[i]) differs = 1;
        if (differs) {
            memcpy(c.final_T, guess, sizeof(I4));
            transform_cloud(&c, guess);
        }
    }
    float ang[3];
    orc_euler_angles_012_f32(c.final_T, ang);
    double p[6] = {c.final_T[12], c.final_T[13], c.final_T[14], ang[0], ang[1], ang[2]};
    double delta_p[6], grad[6], H[36];
    double score = derivs(&c, p, 1, grad, H);
    const double npts = (double)src.n;
    int early = 0;
    while (!converged) {
        double neg_g[6];
        for (int i = 0; i < 6; ++i) neg_g[i] = -grad[i];
        orc_jacobi_svd_solve6(H, neg_g, delta_p, NULL);
        double nrm2 = 0.0;
        for (int i = 0; i < 6; ++i) nrm2 += delta_p[i] * delta_p[i];
        double delta_p_norm = sqrt(nrm2);
        if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
            res->trans_probability = score / npts;
            converged = (delta_p_norm == delta_p_norm);
            early = 1;
            break;
        }
        for (int i = 0; i < 6; ++i) delta_p[i] /= delta_p_norm; /* normalize(): *this /= norm() */
        delta_p_norm = step_length_mt(&c, p, delta_p, delta_p_norm, prm->step_size, prm->trans_eps / 2,
                                      &score, grad, H);
        for (int i = 0; i < 6; ++i) delta_p[i] *= delta_p_norm;
        for (int i = 0; i < 6; ++i) p[i] = p[i] + delta_p[i];
        if (trace && nr_iterations < trace_cap) {
            double *t = &trace[8 * nr_iterations];
            for (int i = 0; i < 6; ++i) t[i] = p[i];
            t[6] = score; t[7] = delta_p_norm;
        }
        if (nr_iterations > prm->max_iter || (nr_iterations && (fabs(delta_p_norm) < prm->trans_eps)))
            converged = 1;
        nr_iterations++;
    }
    if (!early) res->trans_probability = (src.n > 0) ? score / npts : 0.0;
    res->iterations = nr_iterations;
    res->converged = converged;
    res->score = score;
    memcpy(res->p, p, sizeof(p));
    memcpy(pose_out, c.final_T, sizeof(I4));
    if (result_xyz) memcpy(result_xyz, c.trans, src.n * 3 * sizeof(float));
    free(c.trans);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* pcl::Registration::getFitnessScore(max_range)  (ndt_registration.cpp:63-66; Appendix A.4)   */
/* exact 1-NN over the finite target points with a bucket grid + ring expansion               */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int64_t key; uint32_t start, count; } cell_t;

static inline int64_t cell_key(int64_t i, int64_t j, int64_t k) {
    return ((i + (1 << 20)) << 42) | ((j + (1 << 20)) << 21) | (k + (1 << 20));
}
static int cmp_u64pair(const void *a, const void *b) {
    const int64_t *x = (const int64_t *)a, *y = (const int64_t *)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1]);
}

double orc_fitness_score(orc_cloud tgt, orc_cloud src, const float pose[16], double max_range) {
    /* bucket the target */
    size_t nt = 0;
    int64_t *kv = (int64_t *)malloc((tgt.n ? tgt.n : 1) * 2 * sizeof(int64_t));
    const double cs = 1.0; /* bucket edge (m) */
    for (size_t i = 0; i < tgt.n; ++i) {
        const float *p = pt_xyz(tgt, i);
        if (!finite3(p)) continue;
        kv[2 * nt] = cell_key((int64_t)floor(p[0] / cs), (int64_t)floor(p[1] / cs), (int64_t)floor(p[2] / cs));
        kv[2 * nt + 1] = (int64_t)i;
        ++nt;
    }
    if (nt == 0 || src.n == 0) { free(kv); return DBL_MAX; }
    qsort(kv, nt, 2 * sizeof(int64_t), cmp_u64pair);
    size_t ncell = 0;
    for (size_t s = 0; s < nt;) { size_t e = s + 1; while (e < nt && kv[2 * e] == kv[2 * s]) ++e; ++ncell; s = e; }
    size_t hcap = 16;
    while (hcap < 2 * ncell + 16) hcap <<= 1;
    cell_t *ht = (cell_t *)malloc(hcap * sizeof(cell_t));
    for (size_t i = 0; i < hcap; ++i) ht[i].key = -1;
    float *tp = (float *)malloc(nt * 3 * sizeof(float));
    for (size_t s = 0; s < nt;) {
        size_t e = s + 1;
        while (e < nt && kv[2 * e] == kv[2 * s]) ++e;
        uint64_t hk = (uint64_t)kv[2 * s] * 0x9E3779B97F4A7C15ULL;
        size_t h = (size_t)(hk >> 20) & (hcap - 1);
        while (ht[h].key != -1) h = (h + 1) & (hcap - 1);
        ht[h].key = kv[2 * s]; ht[h].start = (uint32_t)s; ht[h].count = (uint32_t)(e - s);
        s = e;
    }
    for (size_t i = 0; i < nt; ++i) {
        const float *p = pt_xyz(tgt, (size_t)kv[2 * i + 1]);
        tp[3 * i] = p[0]; tp[3 * i + 1] = p[1]; tp[3 * i + 2] = p[2];
    }
    double total = 0.0;
    long long nr = 0;
    for (size_t i = 0; i < src.n; ++i) {
        const float *p = pt_xyz(src, i);
        float q[3];
        orc_transform_point_f32(pose, p[0], p[1], p[2], q);
        if (!(isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2]))) continue;
        int64_t ci = (int64_t)floor(q[0] / cs), cj = (int64_t)floor(q[1] / cs), ck = (int64_t)floor(q[2] / cs);
        float best = FLT_MAX;
        int found_ring = -1;
        const int RMAX = 48;
        for (int r = 0; r <= RMAX; ++r) {
            for (int64_t dk = -r; dk <= r; ++dk)
                for (int64_t dj = -r; dj <= r; ++dj)
                    for (int64_t di = -r; di <= r; ++di) {
                        int64_t ad = di < 0 ? -di : di, bd = dj < 0 ? -dj : dj, cd = dk < 0 ? -dk : dk;
                        int64_t ch = ad > bd ? ad : bd; if (cd > ch) ch = cd;
                        if (ch != r) continue;
                        int64_t key = cell_key(ci + di, cj + dj, ck + dk);
                        uint64_t hk = (uint64_t)key * 0x9E3779B97F4A7C15ULL;
                        size_t h = (size_t)(hk >> 20) & (hcap - 1);
                        while (ht[h].key != -1 && ht[h].key != key) h = (h + 1) & (hcap - 1);
                        if (ht[h].key == -1) continue;
                        for (uint32_t k = ht[h].start; k < ht[h].start + ht[h].count; ++k) {
                            float dx = q[0] - tp[3 * k], dy = q[1] - tp[3 * k + 1], dz = q[2] - tp[3 * k + 2];
                            float d2 = (dx * dx + dy * dy) + dz * dz;
                            if (d2 < best) best = d2;
                        }
                    }
            if (best < FLT_MAX) {
                if (found_ring < 0) found_ring = r;
                /* everything not yet visited is farther than r*cs from q */
                double lim = (double)r * cs;
                if ((double)best * (1.0 + 1e-5) <= lim * lim) break;
            }
            if (r == RMAX) { /* brute force fallback */
                for (size_t k = 0; k < nt; ++k) {
                    float dx = q[0] - tp[3 * k], dy = q[1] - tp[3 * k + 1], dz = q[2] - tp[3 * k + 2];
                    float d2 = (dx * dx + dy * dy) + dz * dz;
                    if (d2 < best) best = d2;
                }
            }
        }
        if ((double)best <= max_range) { total += (double)best; ++nr; }
    }
    free(kv); free(ht); free(tp);
    if (nr > 0) return total / (double)nr;
    return DBL_MAX;
}
