"""Oracle NDT restatement: analytic derivatives vs finite differences, convergence on synthetic data,
the regression pin tests/golden/ndt_small.npz, and the fitness score vs brute force."""
import os

import numpy as np

from lidar_slam_b200 import synth
from tests.conftest import f32

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ndt_small.npz")


def test_gradient_and_hessian_match_finite_differences(oracle, small_map, scans):
    """Analytic score gradient / Hessian (Magnusson eq. 6.12/6.13 as coded in NDTM:448-520) against
    finite differences of an independent numpy evaluation of the score over a FIXED pair list (the
    radius search makes the true score piecewise smooth)."""
    truth, scan = scans[1]
    src, _, _, _ = oracle.voxel_filter(scan, 1.3, 1.3, 1.3)
    src = src[::2]
    grid = oracle.Grid(small_map, 1.0)
    lv = grid.leaves()
    prm = oracle.params()
    d1, d2 = oracle.gauss_constants(prm.outlier_ratio, prm.res)
    p = truth + np.array([0.15, -0.1, 0.05, 0.004, -0.003, 0.006])

    def transform(q):
        T = synth.pose6_to_matrix(q)
        return src[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]

    tr = transform(p).astype(np.float32)
    s0, g0, H0, pairs = oracle.derivatives(grid, prm, src, p, trans_xyz=tr)
    pi, ps = [], []
    for i, q in enumerate(tr):
        slots, _ = grid.radius_search(q)
        pi += [i] * len(slots); ps += slots.tolist()
    pi = np.array(pi); ps = np.array(ps)
    assert len(pi) == pairs and pairs > 500
    mu = lv["mean"][ps]; ic = lv["icov"][ps].reshape(-1, 3, 3)

    def score(q):
        x = transform(q)[pi] - mu
        m = np.einsum("ni,nij,nj->n", x, ic, x)
        return float(np.sum(-d1 * np.exp(-d2 * m / 2)))

    assert abs(score(p) - s0) <= 1e-5 * abs(s0)          # float rounding of the transformed points only
    hs = [1e-4] * 3 + [1e-5] * 3
    fd_g = np.zeros(6); fd_H = np.zeros((6, 6))
    for i in range(6):
        ei = np.zeros(6); ei[i] = hs[i]
        fd_g[i] = (score(p + ei) - score(p - ei)) / (2 * hs[i])
        for j in range(6):
            ej = np.zeros(6); ej[j] = hs[j]
            fd_H[i, j] = (score(p + ei + ej) - score(p + ei - ej) - score(p - ei + ej) + score(p - ei - ej)) / (4 * hs[i] * hs[j])
    assert np.allclose(fd_g, g0, rtol=0, atol=2e-4 * np.abs(g0).max())
    assert np.allclose(fd_H, H0, rtol=0, atol=2e-3 * np.abs(H0).max())
    assert np.allclose(H0, H0.T, rtol=0, atol=1e-9 * np.abs(H0).max())


def test_align_converges_to_truth(oracle, small_map, scans):
    grid = oracle.Grid(small_map, 1.0)
    rng = np.random.default_rng(11)
    for compat in (1, 0):
        prm = oracle.params(step_size=f32(0.1), trans_eps=f32(0.01), pcl17_compat=compat)
        for truth, scan in scans[:3]:
            src, _, _, _ = oracle.voxel_filter(scan, 1.3, 1.3, 1.3)
            guess = synth.pose6_to_matrix(synth.perturb_pose(truth, rng, 0.3, 1.0)).astype(np.float32)
            r = oracle.align(grid, prm, src, guess, want_cloud=True)
            assert r["converged"] and 1 <= r["iterations"] <= 32
            Tt = synth.pose6_to_matrix(truth)
            # NDT stops once the Newton step is < trans_eps; on a sparse map the optimum of the score is
            # within a few decimetres of the ground truth, and always an improvement over the guess
            assert np.max(np.abs(r["pose"][:3, 3] - Tt[:3, 3])) < 0.25
            assert np.max(np.abs(r["pose"][:3, :3] - Tt[:3, :3])) < 1e-2
            s_guess = oracle.derivatives(grid, prm, src, np.concatenate([guess[:3, 3], oracle.euler_angles(guess)]))[0]
            assert r["score"] > s_guess
            # align's output cloud is the source under the final pose, in PCL's float arithmetic
            assert np.array_equal(r["cloud"], oracle.transform_points(r["pose"], src[:, :3]))
            if compat:
                assert r["mt_trials"] == 0 and r["passes"] == r["iterations"] + 1   # dead More-Thuente loop of PCL 1.7


def test_degenerate_inputs(oracle, small_map):
    prm = oracle.params()
    grid = oracle.Grid(small_map, 1.0)
    # source far away from every voxel: score 0, H = 0 -> delta = 0 -> immediate converged exit
    far = np.array([[1e4, 1e4, 50, 0], [1e4 + 1, 1e4, 50, 0]], np.float32)
    r = oracle.align(grid, prm, far, np.eye(4, dtype=np.float32))
    assert r["converged"] and r["iterations"] == 0 and r["score"] == 0
    assert np.array_equal(r["pose"], np.eye(4, dtype=np.float32))
    # empty target
    g0 = oracle.Grid(np.zeros((0, 4), np.float32), 1.0)
    r = oracle.align(g0, prm, far, np.eye(4, dtype=np.float32))
    assert r["converged"] and r["iterations"] == 0


def test_golden_regression(oracle):
    G = np.load(GOLD)
    out, idx, cnt, _ = oracle.voxel_filter(G["raw"], 1.3, 1.3, 1.3)
    assert np.array_equal(idx, G["raw_idx"]) and np.array_equal(cnt, G["raw_cnt"]) and np.array_equal(out, G["raw_filt"])
    grid = oracle.Grid(G["target"], 1.0)
    lv = grid.leaves()
    assert np.array_equal(lv["idx"], G["leaf_idx"]) and np.array_equal(lv["n_raw"], G["leaf_n"])
    assert np.array_equal(lv["centroid"], G["leaf_centroid"]) and np.array_equal(lv["mean"], G["leaf_mean"])
    assert np.allclose(lv["icov"], G["leaf_icov"], rtol=1e-12, atol=0)
    prm = oracle.params(step_size=f32(0.1), trans_eps=f32(0.01))
    for k, q in enumerate(G["deriv_pose"]):
        s, g, H, pairs = oracle.derivatives(grid, prm, G["src"], q)
        assert pairs == G["deriv_pairs"][k]
        assert abs(s - G["deriv_score"][k]) <= 1e-11 * abs(s)
        assert np.allclose(g, G["deriv_grad"][k], rtol=1e-10, atol=1e-10 * np.abs(g).max())
        assert np.allclose(H, G["deriv_hess"][k], rtol=1e-10, atol=1e-10 * np.abs(H).max())
    for compat in (1, 0):
        prm = oracle.params(step_size=f32(0.1), trans_eps=f32(0.01), pcl17_compat=compat)
        tag = "c%d_" % compat
        for k, guess in enumerate(G["guesses"]):
            r = oracle.align(grid, prm, G["src"], guess)
            assert r["iterations"] == G[tag + "iterations"][k] and r["converged"] == G[tag + "converged"][k]
            assert np.allclose(r["p"], G[tag + "p"][k], rtol=0, atol=1e-9)
            assert np.allclose(r["pose"], G[tag + "pose"][k], rtol=0, atol=1e-6)
    for k in range(len(G["guesses"])):
        f = oracle.fitness_score(G["target"], G["src"], G["c1_pose"][k])
        assert abs(f - G["fitness"][k]) <= 1e-12 * f


def test_fitness_score_vs_bruteforce(oracle, small_map, scans):
    truth, scan = scans[2]
    src = scan[::97]
    T = synth.pose6_to_matrix(truth).astype(np.float32)
    got = oracle.fitness_score(small_map, src, T)
    q = oracle.transform_points(T, src[:, :3])
    tgt = small_map[:, :3]
    tot = 0.0
    for p in q:
        d = p[None, :] - tgt
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        tot += float(d2.min())
    assert abs(got - tot / len(q)) <= 1e-12 * got
    # max_range drops far correspondences
    lim = float(np.median([0.05]))
    got2 = oracle.fitness_score(small_map, src, T, max_range=lim)
    assert got2 <= lim
