"""Oracle VoxelGrid / VoxelGridCovariance restatement: known answers, an independent numpy
re-derivation of the PCL index formula, and the edge cases PCL handles."""
import numpy as np


def np_voxel_index(pts, leaf):
    """Independent float32 re-derivation: floor(x * (1/leaf)) - min_b, idx = i + j*dx + k*dx*dy."""
    inv = (np.float32(1.0) / np.float32(leaf)).astype(np.float32)
    fin = np.all(np.isfinite(pts[:, :3]), axis=1)
    p = pts[fin, :3].astype(np.float32)
    mn = p.min(0); mx = p.max(0)
    min_b = np.floor(mn * inv).astype(np.int64); max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = (np.floor(p * inv) - min_b.astype(np.float32)).astype(np.int64)
    return ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1], fin, min_b, div


def test_known_answer_tiny(oracle):
    pts = np.array([[0.1, 0.1, 0.1, 1.0], [0.2, 0.2, 0.2, 3.0], [1.5, 0.1, 0.1, 5.0], [-0.5, 0.1, 0.1, 7.0]], np.float32)
    out, idx, cnt, ov = oracle.voxel_filter(pts, 1.0, 1.0, 1.0)
    # min_b.x = -1 -> x cells: -0.5 -> 0, 0.1/0.2 -> 1, 1.5 -> 2
    assert not ov and list(idx) == [0, 1, 2] and list(cnt) == [1, 2, 1]
    assert np.array_equal(out[0], pts[3])
    exp = (pts[0] + pts[1]) / np.float32(2.0)
    assert np.array_equal(out[1], exp)
    assert np.array_equal(out[2], pts[2])


def test_index_formula_and_stable_sums(oracle, scans):
    _, scan = scans[0]
    for leaf in (1.3, 0.6, 0.25):
        out, idx, cnt, ov = oracle.voxel_filter(scan, leaf, leaf, leaf)
        keys, fin, min_b, div = np_voxel_index(scan, leaf)
        uk, ucnt = np.unique(keys, return_counts=True)
        assert not ov
        assert np.array_equal(idx, uk.astype(np.int32)) and np.array_equal(cnt, ucnt.astype(np.int32))
        assert cnt.sum() == fin.sum() and np.all(np.diff(idx) > 0)
        # centroid = sequential float32 sum in input order / float32(count), checked on a few voxels
        order = np.argsort(keys, kind="stable")
        sk = keys[order]
        for v in (0, len(uk) // 2, len(uk) - 1):
            members = order[sk == uk[v]]
            acc = np.zeros(4, np.float32)
            for m in members:
                acc = acc + scan[m]
            assert np.array_equal(out[v], acc / np.float32(len(members)))
        lay = oracle.vox_layout(scan, leaf, leaf, leaf)
        assert list(lay.min_b) == list(min_b) and list(lay.div_b) == list(div)


def test_edge_cases(oracle):
    # empty cloud
    out, idx, cnt, ov = oracle.voxel_filter(np.zeros((0, 4), np.float32), 1, 1, 1)
    assert len(out) == 0
    # non-finite points are skipped (getMinMax3D / applyFilter on non-dense clouds)
    pts = np.array([[0.5, 0.5, 0.5, 1], [np.nan, 0, 0, 2], [0.6, 0.6, np.inf, 3], [0.7, 0.5, 0.5, 5]], np.float32)
    out, idx, cnt, ov = oracle.voxel_filter(pts, 1, 1, 1)
    assert len(out) == 1 and cnt[0] == 2 and np.array_equal(out[0], (pts[0] + pts[3]) / np.float32(2))
    # all non-finite
    out, idx, cnt, ov = oracle.voxel_filter(np.full((3, 4), np.nan, np.float32), 1, 1, 1)
    assert len(out) == 0
    # single point
    out, idx, cnt, ov = oracle.voxel_filter(np.array([[3, 4, 5, 9]], np.float32), 0.5, 0.5, 0.5)
    assert len(out) == 1 and idx[0] == 0 and np.array_equal(out[0], [3, 4, 5, 9])
    # PCL's int32 index guard: leaf too small for the extent -> output = input
    pts = np.array([[0, 0, 0, 1], [5000, 5000, 5000, 2]], np.float32)
    out, idx, cnt, ov = oracle.voxel_filter(pts, 0.01, 0.01, 0.01)
    assert ov and np.array_equal(out, pts)
    # PointXYZI (stride 32) layout gives the same result as the packed one
    rng = np.random.default_rng(3)
    c4 = rng.uniform(-20, 20, (5000, 4)).astype(np.float32)
    c8 = np.zeros((5000, 8), np.float32); c8[:, :3] = c4[:, :3]; c8[:, 3] = 1; c8[:, 4] = c4[:, 3]
    a = oracle.voxel_filter(c4, 0.8, 0.9, 1.1); b = oracle.voxel_filter(c8, 0.8, 0.9, 1.1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_filter_is_idempotent_on_counts(oracle, scans):
    """filtering the filtered cloud with the same leaf keeps one point per voxel"""
    _, scan = scans[1]
    out, idx, cnt, _ = oracle.voxel_filter(scan, 1.3, 1.3, 1.3)
    out2, idx2, cnt2, _ = oracle.voxel_filter(out, 1.3, 1.3, 1.3)
    assert len(out2) <= len(out) and cnt2.sum() == len(out)


def test_target_grid_semantics(oracle, small_map):
    g = oracle.Grid(small_map, 1.0)
    lv = g.leaves()
    keys, fin, min_b, div = np_voxel_index(small_map, 1.0)
    uk, ucnt = np.unique(keys, return_counts=True)
    assert np.array_equal(lv["idx"], uk.astype(np.int32)) and np.array_equal(lv["n_raw"], ucnt.astype(np.int32))
    tree = lv["n_raw"] >= 6
    assert np.array_equal(lv["in_tree"].astype(bool), tree)
    # identity-initialised accumulator => cov = (n-1)/n (S_biased + I/n): eigenvalues >= (n-1)/n^2 > 0
    assert np.all(lv["nr_points"][tree] == lv["n_raw"][tree])
    ev = lv["evals"][tree]
    assert np.all(ev[:, 0] > 0) and np.all(ev[:, 0] >= 0.01 * ev[:, 2] * (1 - 1e-12))
    # icov * cov = I
    for j in np.nonzero(tree)[0][:200]:
        P = lv["icov"][j].reshape(3, 3) @ lv["cov"][j].reshape(3, 3)
        assert np.allclose(P, np.eye(3), atol=1e-8)
    # mean in double vs float centroid
    assert np.allclose(lv["mean"][tree], lv["centroid"][tree][:, :3], atol=2e-3)
    # radius search = brute force over tree centroids with float L2 < r^2, sorted by distance
    cen = lv["centroid"][:, :3]
    rng = np.random.default_rng(5)
    for q in small_map[rng.integers(0, len(small_map), 200), :3]:
        q = (q + rng.normal(0, 0.3, 3)).astype(np.float32)
        d = q[None, :] - cen
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        want = np.nonzero(tree & (d2 < np.float32(1.0)))[0]
        slots, dd = g.radius_search(q)
        assert sorted(slots.tolist()) == sorted(want.tolist())
        assert np.all(np.diff(dd) >= 0)


def test_box_filter_oracle_semantics():
    """pcl::CropBox as BoxFilter uses it: inclusive bounds, order kept, non-finite points dropped."""
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    c = (rng.uniform(-10, 10, size=(5000, 4))).astype(np.float32)
    c[5, 1] = np.nan
    edge = [-3.0, 4.0, -2.0, 2.0, -1.0, 8.0]
    out = O.box_filter(c, edge)
    keep = [i for i in range(len(c)) if np.isfinite(c[i, :3]).all() and edge[0] <= c[i, 0] <= edge[1]
            and edge[2] <= c[i, 1] <= edge[3] and edge[4] <= c[i, 2] <= edge[5]]
    assert np.array_equal(out, c[keep])
    on_edge = np.array([[-3.0, -2.0, -1.0, 0.5], [4.0, 2.0, 8.0, 0.25]], np.float32)
    assert np.array_equal(O.box_filter(on_edge, edge), on_edge)
    assert O.box_filter(np.zeros((0, 4), np.float32), edge).shape == (0, 4)


def test_transform_cloud_oracle_keeps_intensity_and_nonfinite():
    from oracle import oracle as O
    c = np.array([[1, 2, 3, 0.5], [np.nan, 0, 0, 0.7]], np.float32)
    T = np.eye(4, dtype=np.float32); T[:3, 3] = [1, 1, 1]
    out = O.transform_cloud(c, T)
    assert np.array_equal(out[0], np.array([2, 3, 4, 0.5], np.float32))
    assert np.isnan(out[1, 0]) and out[1, 3] == np.float32(0.7)


def test_gauss2d_cells_and_yaw_oracle_small_case():
    """Matching::generateGauss2DMapCells / getInitialYawAngle restatement: hand-checkable grid, heading recovered."""
    from oracle import oracle as O
    pts = np.array([[0, 0, 1, 0], [0.1, 0.1, 3, 0], [4, 0, 5, 0], [4, 4, 7, 0], [0, 4, 2, 0], [0.2, 0, 2, 0]], np.float32)
    c = O.gauss2d_map_cells(pts, [0, 0, 0], 0.8)
    assert c["width"] == 5 and c["height"] == 5           # round(4 / 0.8)
    assert c["cnt"][0, 0] == 3 and c["mu"][0, 0] == np.float32(2.0)          # z = 1, 3, 2
    assert c["cnt"].sum() == 3                             # the points on the max edge fall outside [0, width)
    # the reference's recurrence for z = 1, 3, 2: after z=3: (0 + 4 + 2*1 + 2*1*1)/1 = 8; after z=2: (1*8 + 0 + 0 + 0)/2 = 4
    assert c["sigma"][0, 0] == np.float32(4.0)
    # an L-shaped wall: rotating the scan by the right yaw maximises the score
    def world(seed, n):
        rng = np.random.default_rng(seed)
        w = np.concatenate([np.stack([rng.uniform(0, 20, n), np.zeros(n) + 10, rng.uniform(0, 3, n)], 1),
                            np.stack([np.zeros(n) + 15, rng.uniform(-10, 10, n), rng.uniform(0, 3, n)], 1),
                            np.stack([rng.uniform(-20, 20, 2 * n), rng.uniform(-20, 20, 2 * n), rng.normal(0, 0.02, 2 * n)], 1)])
        return np.concatenate([w, np.zeros((len(w), 1))], 1).astype(np.float32)

    cells = O.gauss2d_map_cells(world(0, 4000), [0, 0, 0], 0.8)
    yaw = 1.0
    R = np.array([[np.cos(-yaw), -np.sin(-yaw)], [np.sin(-yaw), np.cos(-yaw)]])
    scan = world(1, 800)                                  # other samples of the same surfaces
    scan[:, :2] = scan[:, :2] @ R.T                       # what the sensor sees when the vehicle is yawed by +1 rad
    best, probs = O.initial_yaw_angle(cells, scan, 90)
    assert abs(best - yaw) <= 2 * np.pi / 90


def test_synthetic_drive_stays_on_the_streets():
    """The L-shaped drive of the synthetic scene (street canyon) must never enter the footprint of
    a building or a parked car, on either leg or in the corner -- a scan taken from inside a box is not a street
    scene and the NDT (oracle and GPU alike) loses track there."""
    from lidar_slam_b200 import synth
    scene = synth.Scene(leg=500.0)
    s = np.arange(0.0, scene.path_length, 0.25)
    poses = np.stack([scene.path_pose(v) for v in s])
    inside = [v for v, p in zip(s, poses) if scene.point_in_box(p[0], p[1], margin=0.5)]
    assert not inside, "drive enters a box at s = %s" % inside[:5]
    # both legs are street centre lines of the 50 m lattice (x or y = 25 + 50 k) up to the 0.6 m lane weave; the
    # quarter turn (r = 12 m) cuts the intersection by at most r (1 - cos 45 deg) = 3.5 m, inside the 5 m clearance
    on_street = np.minimum(np.abs((poses[:, 0] - 25.0 + 25.0) % 50.0 - 25.0), np.abs((poses[:, 1] - 25.0 + 25.0) % 50.0 - 25.0))
    assert on_street.max() < 3.6 and np.percentile(on_street, 95) < 0.7
    # the heading is continuous (no jump at the ends of the quarter turn)
    assert np.abs(np.diff(poses[:, 5])).max() < 0.25 / 12.0 + 1e-3
