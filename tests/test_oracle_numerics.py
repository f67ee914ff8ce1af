"""Oracle (plain-C restatement) and product host-compiled math vs the REAL Eigen 3.2.92 of the reference.

Golden vectors in tests/golden/eigen_numerics.npz were produced by tools/make_golden.py from
oracle/_ref/libeigen_ref.so (the vendored Eigen compiled where it lies).  When that library is present
(build container) the same checks also run live on fresh random inputs.
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests.hostcheck_util import hc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "eigen_numerics.npz")


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _relerr(a, b):
    s = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / s)


def test_svd_solve_matches_eigen(oracle, gold):
    for t in range(len(gold["svd_rank"])):
        x, sv, rank = oracle.svd_solve6(gold["svd_H"][t].reshape(6, 6, order="F"), gold["svd_b"][t])
        assert rank == gold["svd_rank"][t]
        assert _relerr(sv, gold["svd_sv"][t]) < 1e-13
        assert _relerr(x, gold["svd_x"][t]) < 1e-9 * max(1.0, gold["svd_sv"][t][0] / max(gold["svd_sv"][t][rank - 1], 1e-300) * 1e-6) if rank else np.all(x == 0)


def test_product_newton_solve_matches_eigen(gold):
    """csrc/b2_ndt_math.cuh: svd_solve6 == Eigen; LU fast path == Eigen whenever it is taken."""
    L = hc()
    taken = 0
    for t in range(len(gold["svd_rank"])):
        H = np.ascontiguousarray(gold["svd_H"][t]); b = np.ascontiguousarray(gold["svd_b"][t])
        x = np.zeros(6)
        rank = L.hc_svd_solve6(dp(H), dp(b), dp(x))
        assert rank == gold["svd_rank"][t]
        ref = gold["svd_x"][t]
        sv = gold["svd_sv"][t]
        cond = sv[0] / max(sv[rank - 1], 1e-300) if rank else 1.0
        assert _relerr(x, ref) < 1e-13 * max(cond, 1.0) * 10 if rank else np.all(x == 0)
        x2 = np.zeros(6)
        rc = L.hc_lu_solve6(dp(H), dp(b), dp(x2))
        L.hc_newton_solve6(dp(H), dp(b), dp(x), 0)
        if rc > 1e-9 and rank == 6:
            taken += 1
            assert _relerr(x, ref) < 1e-15 * cond * 100 + 1e-12
        else:
            # falls back to the SVD: rank-truncated solution, identical to Eigen's
            assert _relerr(x, ref) < 1e-13 * max(cond, 1.0) * 10 if rank else np.all(x == 0)
    assert taken > 50


def test_pose_matrix_matches_eigen(oracle, gold):
    L = oracle.lib()
    L.orc_set_f32_trig_libm(1)     # float sin/cos through libm, as the reference build calls them
    try:
        exact = 0
        for t in range(len(gold["pose_p"])):
            T = np.zeros(16, np.float32)
            L.orc_pose_to_matrix_f32(dp(np.ascontiguousarray(gold["pose_p"][t])), fp(T))
            exact += np.array_equal(T, gold["pose_T"][t])
            assert np.max(np.abs(T - gold["pose_T"][t])) <= 1.2e-7
        # identical composition; only libm version differences in sinf/cosf can flip a last bit
        assert exact >= 0.98 * len(gold["pose_p"])
    finally:
        L.orc_set_f32_trig_libm(0)
    # default mode (float trig through double, shared with the CUDA path): at most 1 ulp from Eigen+libm
    H = hc()
    for t in range(len(gold["pose_p"])):
        T = np.zeros(16, np.float32); T2 = np.zeros(16, np.float32)
        p = np.ascontiguousarray(gold["pose_p"][t])
        L.orc_pose_to_matrix_f32(dp(p), fp(T))
        H.hc_pose_to_matrix(dp(p), fp(T2))
        assert np.array_equal(T, T2)                       # oracle == product math, bit for bit
        assert np.max(np.abs(T[:12] - gold["pose_T"][t][:12])) <= 1.2e-7
        assert np.array_equal(T[12:], gold["pose_T"][t][12:])


def test_euler_matches_eigen(oracle, gold):
    H = hc()
    for t in range(len(gold["pose_p"])):
        T = np.ascontiguousarray(gold["pose_T"][t])
        e = np.zeros(3, np.float32); e2 = np.zeros(3, np.float32)
        oracle.lib().orc_euler_angles_012_f32(fp(T), fp(e))
        H.hc_euler(fp(T), fp(e2))
        assert np.array_equal(e, e2)                       # oracle == product math
        d = np.abs(e - gold["pose_euler"][t])
        d = np.minimum(d, np.abs(d - 2 * np.pi))           # +-pi wrap of the same angle
        assert np.max(d) < 5e-6


def test_leaf_finish_matches_eigen(oracle, gold):
    H = hc()
    for t in range(len(gold["leaf_ret"])):
        off, n = gold["leaf_meta"][t]
        pts = gold["leaf_pts"][off:off + n]
        cloud = np.zeros((n, 4), np.float32); cloud[:, :3] = pts
        g = oracle.Grid(cloud, res=1e5)
        lv = g.leaves()
        assert len(lv) == 1 and lv[0]["n_raw"] == n
        assert lv[0]["nr_points"] == gold["leaf_ret"][t]
        assert np.array_equal(lv[0]["mean"], gold["leaf_mean"][t])
        if gold["leaf_ret"][t] > 0:
            assert _relerr(lv[0]["cov"], gold["leaf_cov"][t]) < 1e-12
            assert _relerr(lv[0]["icov"], gold["leaf_icov"][t]) < 1e-11
        # product math (same source as the GPU kernel) == oracle bit for bit
        s = np.zeros(3); acc = np.eye(3)
        for v in pts.astype(np.float64):
            s = s + v; acc = acc + np.outer(v, v)
        mean = np.zeros(3); cov = np.zeros(9); icov = np.zeros(9); ev = np.zeros(3)
        ret = H.hc_leaf_finish(dp(s), dp(np.ascontiguousarray(acc.flatten())), int(n), 6, 0.01, dp(mean), dp(cov), dp(icov), dp(ev))
        assert ret == lv[0]["nr_points"]
        assert np.array_equal(mean, lv[0]["mean"])
        if ret > 0:
            assert np.array_equal(icov, lv[0]["icov"])


def test_live_against_vendored_eigen(oracle):
    R = oracle.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built here (no /root/reference)")
    rng = np.random.default_rng(99)
    for t in range(300):
        A = rng.standard_normal((6, 6))
        Hm = -(A @ A.T) if t % 2 else A + A.T
        Hc = np.ascontiguousarray(Hm.flatten(order="F")); b = rng.standard_normal(6)
        x2 = np.zeros(6); s2 = np.zeros(6)
        r2 = R.ref_svd_solve6(dp(Hc), dp(b), dp(x2), dp(s2))
        x1, s1, r1 = oracle.svd_solve6(Hm, b)
        assert r1 == r2 and _relerr(x1, x2) < 1e-10
    for t in range(300):
        T = np.zeros(16, np.float32)
        p = np.concatenate([rng.uniform(-100, 100, 3), rng.uniform(-1, 1, 3)])
        R.ref_pose_matrix(dp(p), fp(T))
        q = rng.uniform(-100, 100, 3).astype(np.float32)
        o1 = np.zeros(3, np.float32); o2 = np.zeros(3, np.float32)
        R.ref_transform_point(fp(T), fp(q), fp(o2))
        oracle.lib().orc_transform_point_f32(fp(T), q[0], q[1], q[2], fp(o1))
        assert np.array_equal(o1, o2)
