// b2_hostcheck.cpp -- TEST SHIM (not part of the product library).
// Compiles csrc/b2_ndt_math.cuh -- the exact source thread 0 of every match CTA runs on the GPU -- for
// the host, so the Newton / More-Thuente controller, the 6x6 solve, the float pose composition and the
// per-leaf covariance finish can be checked on a machine without a GPU (tests/test_hostcheck.py drives
// the controller with derivative passes evaluated by the CPU oracle).
#include "../b2_ndt_math.cuh"

#include <cstring>

using namespace b2;

extern "C" {
void hc_sincos(double x, double *s, double *c) { b2_sincos(x, *s, *c); }
void hc_pose_to_matrix(const double *p, float *T) { pose_to_matrix_f32(p, T); }
void hc_euler(const float *T, float *out) { euler_from_matrix_f32(T, out); }
void hc_newton_solve6(const double *H, const double *b, double *x, int force_svd) { newton_solve6(H, b, x, force_svd); }
int hc_svd_solve6(const double *H, const double *b, double *x) { return svd_solve6(H, b, x); }
double hc_lu_solve6(const double *H, const double *b, double *x) { return lu_solve6(H, b, x); }
int hc_leaf_finish(const double *sum, const double *acc, int n, int min_pts, double eig_mult, double *mean, double *cov,
                   double *icov, double *evals) {
    return leaf_finish(sum, acc, n, min_pts, eig_mult, mean, cov, icov, evals);
}
void hc_angle_derivatives(const double *p, double *j24, double *h45) {
    AngTab A;
    angle_derivatives(p, A);
    std::memcpy(j24, A.j, sizeof(A.j));
    std::memcpy(h45, A.h, sizeof(A.h));
}

struct hc_ctl { Ctl c; NdtConst k; };
void *hc_ctl_new(double d1, double d2, double step_size, double trans_eps, int max_iter, int pcl17_compat, int force_svd, float res) {
    hc_ctl *h = new hc_ctl();
    std::memset(h, 0, sizeof(*h));
    h->k.d1 = d1; h->k.d2 = d2; h->k.step_size = step_size; h->k.trans_eps = trans_eps; h->k.max_iter = max_iter;
    h->k.pcl17_compat = pcl17_compat; h->k.force_svd = force_svd; h->k.res = res;
    return h;
}
void hc_ctl_free(void *h) { delete (hc_ctl *)h; }
void hc_ctl_start(void *h, const float *guess, double npoints) {
    ctl_start(((hc_ctl *)h)->c, ((hc_ctl *)h)->k, guess, npoints);
    ctl_finish_request_serial(((hc_ctl *)h)->c);
}
int hc_ctl_step(void *h, const double *acc29) {
    int go = ctl_step(((hc_ctl *)h)->c, ((hc_ctl *)h)->k, acc29);
    if (go) ctl_finish_request_serial(((hc_ctl *)h)->c);
    return go;
}
// request of the next pass
void hc_ctl_request(void *h, float *T16, double *x6, int *hess) {
    Ctl &c = ((hc_ctl *)h)->c;
    std::memcpy(T16, c.T, 64);
    // pose the angle tables were built for: p before the first Newton step, x_t afterwards
    for (int i = 0; i < 6; ++i) x6[i] = c.x_req[i];
    *hess = c.hess;
}
void hc_ctl_result(void *h, float *finalT, double *p6, double *score, double *tp, int *iters, int *conv, int *passes, int *mt) {
    Ctl &c = ((hc_ctl *)h)->c;
    std::memcpy(finalT, c.finalT, 64);
    for (int i = 0; i < 6; ++i) p6[i] = c.p[i];
    *score = c.score; *tp = c.trans_probability; *iters = c.nr_iter; *conv = c.converged; *passes = c.passes; *mt = c.mt_trials;
}
}
