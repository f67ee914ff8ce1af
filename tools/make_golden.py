"""Generate the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists).

  eigen_numerics.npz  outputs of the REAL Eigen 3.2.92 vendored by the reference
                      (/root/reference/lidar_localization/third_party/eigen3, via oracle/_ref/libeigen_ref.so)
                      for the expressions PCL's NDT uses: JacobiSVD<6x6>.solve, Translation*AngleAxis^3,
                      Transform::rotation().eulerAngles(0,1,2), and the per-leaf covariance finish with
                      SelfAdjointEigenSolver.  These pin the plain-C oracle AND the product's device math.
  ndt_small.npz       a small scan-to-map case with the oracle's own outputs (regression pin of the
                      restatement; the reference itself ships no golden vectors).
  intree_ndt.npz      outputs of the reference's OWN in-tree NDT source (compiled where it lies against
                      oracle/ref_stubs) on that case: voxel statistics, angle tables, computeDerivatives,
                      align.  True reference outputs; pin the oracle's NDT core and grid build.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from lidar_slam_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def eigen_numerics():
    O.build(ref=True)
    R = O.ref_lib()
    assert R is not None, "oracle/_ref/libeigen_ref.so missing (needs /root/reference)"
    rng = np.random.default_rng(20261018)
    n = 240
    H = np.zeros((n, 36)); b = np.zeros((n, 6)); x = np.zeros((n, 6)); sv = np.zeros((n, 6)); rank = np.zeros(n, np.int32)
    for t in range(n):
        A = rng.standard_normal((6, 6))
        kind = t % 6
        if kind == 0: M = -(A @ A.T) * 10 ** rng.uniform(-2, 4)          # negative definite (typical NDT Hessian)
        elif kind == 1: M = A + A.T                                      # symmetric indefinite
        elif kind == 2: Bm = rng.standard_normal((6, 4)); M = Bm @ Bm.T   # rank 4
        elif kind == 3: M = A                                            # general
        elif kind == 4: Bm = rng.standard_normal((6, 1)); M = Bm @ Bm.T   # rank 1
        else: M = np.diag(10.0 ** rng.uniform(-8, 8, 6)) @ (A + A.T) @ np.diag(10.0 ** rng.uniform(-3, 3, 6)); M = M + M.T
        if t == n - 1: M = np.zeros((6, 6))
        H[t] = M.flatten(order="F"); b[t] = rng.standard_normal(6)
        rank[t] = R.ref_svd_solve6(dp(H[t]), dp(b[t]), dp(x[t]), dp(sv[t]))
    m = 400
    P = np.zeros((m, 6)); T = np.zeros((m, 16), np.float32); E = np.zeros((m, 3), np.float32)
    for t in range(m):
        scale = [1.0, 0.05, 1e-3, 1e-5][t % 4]
        P[t] = np.concatenate([rng.uniform(-800, 800, 3), rng.uniform(-3.1, 3.1, 3) * scale])
        if t == 0: P[t] = 0
        R.ref_pose_matrix(dp(P[t]), fp(T[t]))
        R.ref_euler012(fp(T[t]), fp(E[t]))
    k = 300
    pts_list = []; meta = np.zeros((k, 2), np.int32)
    mean = np.zeros((k, 3)); cov = np.zeros((k, 9)); icov = np.zeros((k, 9)); ev = np.zeros((k, 3)); ret = np.zeros(k, np.int32)
    for t in range(k):
        npts = int(rng.integers(6, 120))
        c = rng.uniform(100, 1500, 3)
        kind = t % 4
        sig = np.array([0.3, 0.3, 0.3 if kind == 0 else (0.01 if kind == 1 else 1e-4)])
        q = rng.standard_normal((npts, 3)) * sig
        if kind == 2: q[:, 1] *= 1e-3
        if kind == 3: q[:] = q[0]                                         # all points identical
        Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        q = (q @ Q.T + c).astype(np.float32)
        s = np.zeros(3); acc = np.eye(3)
        for v in q.astype(np.float64):
            s += v; acc += np.outer(v, v)
        accf = np.ascontiguousarray(acc.flatten())
        ret[t] = R.ref_leaf_finish(dp(s), dp(accf), npts, 0.01, dp(mean[t]), dp(cov[t]), dp(icov[t]), dp(ev[t]))
        meta[t] = (len(pts_list) and sum(len(p) for p in pts_list), npts)
        pts_list.append(q)
    pts = np.concatenate(pts_list, 0)
    np.savez_compressed(os.path.join(OUT, "eigen_numerics.npz"), svd_H=H, svd_b=b, svd_x=x, svd_sv=sv, svd_rank=rank,
                        pose_p=P, pose_T=T, pose_euler=E, leaf_pts=pts, leaf_meta=meta, leaf_mean=mean, leaf_cov=cov,
                        leaf_icov=icov, leaf_evals=ev, leaf_ret=ret)
    print("eigen_numerics.npz:", n, m, k)


def ndt_small():
    scene = synth.Scene(leg=40.0)
    target = scene.make_map(24000, 2.0)
    truth = scene.path_pose(17.0)
    scan = scene.scan(5, truth)
    filt, idx, cnt, _ = O.voxel_filter(scan, 1.3, 1.3, 1.3)
    src = filt[::2].copy()
    raw = scan[::16].copy()
    r_filt, r_idx, r_cnt, _ = O.voxel_filter(raw, 1.3, 1.3, 1.3)
    grid = O.Grid(target, 1.0)
    lv = grid.leaves()
    f32 = lambda v: float(np.float32(v))
    out = dict(target=target, src=src, raw=raw, raw_filt=r_filt, raw_idx=r_idx, raw_cnt=r_cnt,
               leaf_idx=lv["idx"], leaf_n=lv["n_raw"], leaf_centroid=lv["centroid"], leaf_mean=lv["mean"], leaf_icov=lv["icov"],
               truth=truth)
    rng = np.random.default_rng(7)
    guesses = [synth.pose6_to_matrix(synth.perturb_pose(truth, rng)).astype(np.float32) for _ in range(6)]
    guesses.append(np.eye(4, dtype=np.float32))
    out["guesses"] = np.stack(guesses)
    for compat in (1, 0):
        prm = O.params(step_size=f32(0.1), trans_eps=f32(0.01), pcl17_compat=compat)
        poses, ps, its, conv, sc, tp, passes = [], [], [], [], [], [], []
        for G in guesses:
            r = O.align(grid, prm, src, G)
            poses.append(r["pose"]); ps.append(r["p"]); its.append(r["iterations"]); conv.append(r["converged"])
            sc.append(r["score"]); tp.append(r["trans_probability"]); passes.append(r["passes"])
        tag = "c%d_" % compat
        out[tag + "pose"] = np.stack(poses); out[tag + "p"] = np.stack(ps); out[tag + "iterations"] = np.array(its)
        out[tag + "converged"] = np.array(conv); out[tag + "score"] = np.array(sc); out[tag + "trans_probability"] = np.array(tp)
        out[tag + "passes"] = np.array(passes)
    prm = O.params(step_size=f32(0.1), trans_eps=f32(0.01))
    dposes = np.stack([synth.perturb_pose(truth, rng) for _ in range(4)])
    ds, dg, dH, dpairs = [], [], [], []
    for q in dposes:
        s, g, H, pairs = O.derivatives(grid, prm, src, q)
        ds.append(s); dg.append(g); dH.append(H); dpairs.append(pairs)
    out.update(deriv_pose=dposes, deriv_score=np.array(ds), deriv_grad=np.stack(dg), deriv_hess=np.stack(dH), deriv_pairs=np.array(dpairs))
    out["fitness"] = np.array([O.fitness_score(target, src, out["c1_pose"][i]) for i in range(len(guesses))])
    np.savez_compressed(os.path.join(OUT, "ndt_small.npz"), **out)
    print("ndt_small.npz: target", target.shape, "src", src.shape, "raw", raw.shape, "leaves", len(lv),
          "iterations", out["c1_iterations"], out["c0_iterations"])


def intree_ndt():
    """Outputs of the reference's OWN in-tree NDT (ndt_registration_manual/*.cpp compiled where it lies
    against oracle/ref_stubs, oracle/_ref/libndt_manual_ref.so) on the ndt_small case: per-voxel mean /
    inverse covariance / static value, computeAngleDerivatives tables, computeDerivatives at fixed poses,
    and align() results.  These are true reference outputs (not oracle outputs)."""
    O.build(ref=True)
    R = O.refndt_lib()
    assert R is not None, "oracle/_ref/libndt_manual_ref.so missing (needs /root/reference)"
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    G = np.load(os.path.join(OUT, "ndt_small.npz"))
    target = np.ascontiguousarray(G["target"]); src = np.ascontiguousarray(G["src"])
    f32 = lambda v: float(np.float32(v))
    h = R.refndt_new(1.0, f32(0.1), f32(0.01), 30, 0.55)
    R.refndt_set_target(h, fp(target), len(target))
    info = np.zeros(12, np.int32); R.refndt_grid_info(h, ip(info))
    # enumerate the reference's searchable voxels (points_per_voxel >= 6) over its occupied index range
    ijk, cen, icov, sv, npv = [], [], [], [], []
    for ix in range(info[6], info[9] + 1):
        for iy in range(info[7], info[10] + 1):
            for iz in range(info[8], info[11] + 1):
                c = np.zeros(3); ic = np.zeros(9); s = C.c_double()
                n = R.refndt_voxel(h, ix, iy, iz, dp(c), dp(ic), C.byref(s))
                if n >= 6:
                    ijk.append((ix, iy, iz)); cen.append(c); icov.append(ic); sv.append(s.value); npv.append(n)
    out = dict(grid_info=info, vox_ijk=np.array(ijk, np.int32), vox_mean=np.array(cen), vox_icov=np.array(icov),
               vox_static=np.array(sv), vox_n=np.array(npv, np.int32))
    R.refndt_set_source(h, fp(src), len(src))
    poses = G["deriv_pose"]
    ds, dg, dH, tabs_j, tabs_h, trans = [], [], [], [], [], []
    for q in poses:
        q = np.ascontiguousarray(q)
        tr = np.ascontiguousarray(O.transform_points(O.pose_to_matrix(q), src[:, :3]))
        g6 = np.zeros(6); H = np.zeros(36)
        ds.append(R.refndt_derivatives(h, fp(tr), dp(q), 1, dp(g6), dp(H))); dg.append(g6); dH.append(H.reshape(6, 6, order="F")); trans.append(tr)
        j = np.zeros(24); hh = np.zeros(45)
        R.refndt_angle_tables(h, dp(q), dp(j), dp(hh)); tabs_j.append(j); tabs_h.append(hh)
    out.update(deriv_pose=poses, deriv_trans=np.stack(trans), deriv_score=np.array(ds), deriv_grad=np.stack(dg), deriv_hess=np.stack(dH),
               ang_j=np.stack(tabs_j), ang_h=np.stack(tabs_h))
    al_pose, al_it, al_conv, al_tp, al_cloud = [], [], [], [], []
    for guess in G["guesses"]:
        Gc = np.ascontiguousarray(guess.flatten(order="F")); pose = np.zeros(16, np.float32)
        it = C.c_int(); cv = C.c_int(); tp = C.c_double(); tc = np.zeros((len(src), 3), np.float32)
        R.refndt_align(h, fp(Gc), fp(pose), C.byref(it), C.byref(cv), C.byref(tp), fp(tc))
        al_pose.append(pose.reshape(4, 4, order="F")); al_it.append(it.value); al_conv.append(cv.value); al_tp.append(tp.value); al_cloud.append(tc[::16])
    out.update(guesses=G["guesses"], align_pose=np.stack(al_pose), align_iterations=np.array(al_it), align_converged=np.array(al_conv),
               align_trans_probability=np.array(al_tp), align_cloud_every16=np.stack(al_cloud))
    R.refndt_free(h)
    np.savez_compressed(os.path.join(OUT, "intree_ndt.npz"), **out)
    print("intree_ndt.npz: voxels", len(ijk), "iterations", al_it)


def _ref_voxels(R, h):
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    info = np.zeros(12, np.int32); R.refndt_grid_info(h, ip(info))
    ijk, cen, icov, npv = [], [], [], []
    for ix in range(info[6], info[9] + 1):
        for iy in range(info[7], info[10] + 1):
            for iz in range(info[8], info[11] + 1):
                c = np.zeros(3); ic = np.zeros(9); s = C.c_double()
                n = R.refndt_voxel(h, ix, iy, iz, dp(c), dp(ic), C.byref(s))
                if n >= 6:
                    ijk.append((ix, iy, iz)); cen.append(c); icov.append(ic); npv.append(n)
    return np.array(ijk, np.int32), np.array(cen), np.array(icov), np.array(npv, np.int32)


def intree_update():
    """Outputs of the reference's OWN incremental grid update (NormalDistributionsTransform::updateVoxelGrid ->
    VoxelGrid::update / updateVoxelContent, VoxelGrid.cpp:545-584,736-809) on the ndt_small target split in three:
    setInputTarget(first 60 %) then two updateVoxelGrid calls (next 25 %, last 15 % -- the last part reaches outside the
    first grid, so the reference's updateBoundaries path runs too).  Stored: the searchable voxels after the updates,
    and how far they are from the reference's own setInputTarget over the whole cloud."""
    O.build(ref=True)
    R = O.refndt_lib()
    assert R is not None, "oracle/_ref/libndt_manual_ref.so missing (needs /root/reference)"
    G = np.load(os.path.join(OUT, "ndt_small.npz"))
    target = np.ascontiguousarray(G["target"])
    # order the cloud so that the last part extends the bounding box (sorted by x within the last 15 %)
    n = len(target)
    a, b = int(0.6 * n), int(0.85 * n)
    order = np.argsort(target[:, 0], kind="stable")
    tail = order[-(n - b):]
    rest = np.setdiff1d(np.arange(n), tail)           # ascending: input order kept
    target = np.ascontiguousarray(np.concatenate([target[rest], target[tail]]))
    f32 = lambda v: float(np.float32(v))
    h = R.refndt_new(1.0, f32(0.1), f32(0.01), 30, 0.55)
    R.refndt_set_target(h, fp(target[:a]), a)
    R.refndt_update(h, fp(np.ascontiguousarray(target[a:b])), b - a)
    R.refndt_update(h, fp(np.ascontiguousarray(target[b:])), n - b)
    ijk, cen, icov, npv = _ref_voxels(R, h)
    R.refndt_free(h)
    h = R.refndt_new(1.0, f32(0.1), f32(0.01), 30, 0.55)
    R.refndt_set_target(h, fp(target), n)
    ijk2, cen2, icov2, npv2 = _ref_voxels(R, h)
    R.refndt_free(h)
    same = np.array_equal(ijk, ijk2) and np.array_equal(npv, npv2)
    d_mean = float(np.max(np.abs(cen - cen2))) if same else np.inf
    d_icov = float(np.max(np.abs(icov - icov2).max(1) / np.abs(icov2).max(1))) if same else np.inf
    np.savez_compressed(os.path.join(OUT, "intree_update.npz"), target=target, split=np.array([a, b]), vox_ijk=ijk, vox_mean=cen,
                        vox_icov=icov, vox_n=npv, full_build_same_voxels=np.array(same), full_build_max_dmean=np.array(d_mean),
                        full_build_max_dicov_rel=np.array(d_icov))
    print("intree_update.npz: voxels", len(ijk), "vs the reference's full build: same voxel set", same, "max |dmean|", d_mean, "max rel dicov", d_icov)


def intree_matching():
    """Outputs of the reference's OWN position-only initialisation (matching.cpp: generateGauss2DMapCells :344-394,
    getInitialYawAngle :267-308, compiled where it lies into oracle/_ref/libmatching_ref.so) on a seeded local map (with
    a few non-finite points) and two scans: the Gaussian height grid (mu / sigma / count per cell) and the winning yaw."""
    from lidar_slam_b200 import synth
    O.build(ref=True)
    scene = synth.Scene(leg=60.0)
    m = scene.make_map(20000, 2.0)
    m[::501, 0] = np.nan
    origin = np.array([10.0, 2.0, 0.5], np.float32)
    scans = []
    for k, s in enumerate((20.0, 33.0)):
        sc = scene.scan(5 + k, scene.path_pose(s))[::40].copy()
        sc[::97, 1] = np.nan
        # the scan as the matching node sees it: sensor frame rotated by an unknown heading about the local-map origin
        scans.append(sc)
    ref, yaws = O.gauss2d_map_cells_reference(m, origin, 0.8, scans)
    np.savez_compressed(os.path.join(OUT, "intree_matching.npz"), local_map=m, origin=origin, grid_resolution=np.array(0.8),
                        scan0=scans[0], scan1=scans[1], width=np.array(ref["width"]), height=np.array(ref["height"]),
                        min_xyz=ref["min_xyz"], max_xyz=ref["max_xyz"], mu=ref["mu"], sigma=ref["sigma"], cnt=ref["cnt"],
                        yaw=np.array(yaws))
    print("intree_matching.npz: grid", ref["width"], "x", ref["height"], "occupied", int((ref["cnt"] > 0).sum()), "yaw", yaws)


def deskew():
    """Outputs of the reference's OWN DistortionAdjust (oracle/_ref/libdeskew_ref.so = distortion_adjust.cpp compiled
    where it lies against oracle/ref_stubs + the vendored Eigen) on a seeded synthetic sweep."""
    from lidar_slam_b200 import synth
    scene = synth.Scene(leg=60.0)
    scan = scene.scan(31, scene.path_pose(14.0))[::16].copy()
    cases = []
    for lin, ang, period in (([8.0, 0.3, -0.1], [0.02, -0.01, 0.35], 0.1), ([0.0, 0.0, 0.0], [0.0, 0.0, 0.0], 0.1),
                             ([-3.0, 1.5, 0.2], [0.3, 0.2, -0.6], 0.05)):
        out = O.distortion_adjust_reference(scan, period, lin, ang)
        cases.append((np.array(lin), np.array(ang), period, out))
    np.savez_compressed(os.path.join(OUT, "deskew.npz"), scan=scan,
                        lin=np.stack([c[0] for c in cases]), ang=np.stack([c[1] for c in cases]),
                        period=np.array([c[2] for c in cases]), n_out=np.array([len(c[3]) for c in cases]),
                        **{"out%d" % i: c[3] for i, c in enumerate(cases)})
    print("deskew.npz: scan", len(scan), "outputs", [len(c[3]) for c in cases])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    eigen_numerics()
    ndt_small()
    intree_ndt()
    intree_update()
    intree_matching()
    deskew()
