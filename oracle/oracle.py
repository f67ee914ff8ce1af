"""ctypes binding of the CPU oracle (oracle/libndt_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this
module.  The product package (lidar_slam_b200) never does.

Clouds are numpy float32 arrays of shape (n, 4) = packed {x, y, z, intensity} (stride 16) or
(n, 8) = pcl::PointXYZI memory layout (stride 32, intensity in column 4; cloud_data.hpp:35).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(ref=True):
    """Compile the C restatement (always) and, where /root/reference exists, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    if ref and os.path.isdir("/root/reference/lidar_localization/third_party/eigen3"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


class Cloud(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("stride", C.c_size_t), ("ioff", C.c_size_t)]


class VoxLayout(C.Structure):
    _fields_ = [("ok", C.c_int), ("inv", C.c_float * 3), ("min_p", C.c_float * 3), ("max_p", C.c_float * 3),
                ("min_b", C.c_int32 * 3), ("max_b", C.c_int32 * 3), ("div_b", C.c_int32 * 3),
                ("divb_mul", C.c_int32 * 3), ("n_finite", C.c_size_t)]


class Leaf(C.Structure):
    _fields_ = [("idx", C.c_int32), ("n_raw", C.c_int32), ("nr_points", C.c_int32), ("in_tree", C.c_int32),
                ("centroid", C.c_float * 4), ("mean", C.c_double * 3), ("cov", C.c_double * 9),
                ("icov", C.c_double * 9), ("evals", C.c_double * 3), ("weight", C.c_double)]


LEAF_DTYPE = np.dtype([("idx", "<i4"), ("n_raw", "<i4"), ("nr_points", "<i4"), ("in_tree", "<i4"),
                       ("centroid", "<f4", (4,)), ("mean", "<f8", (3,)), ("cov", "<f8", (9,)),
                       ("icov", "<f8", (9,)), ("evals", "<f8", (3,)), ("weight", "<f8")], align=True)


class Params(C.Structure):
    _fields_ = [("res", C.c_float), ("step_size", C.c_double), ("trans_eps", C.c_double),
                ("outlier_ratio", C.c_double), ("max_iter", C.c_int), ("min_pts", C.c_int),
                ("eig_mult", C.c_double), ("pcl17_compat", C.c_int), ("intree_compat", C.c_int)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("score", C.c_double),
                ("trans_probability", C.c_double), ("p", C.c_double * 6), ("passes", C.c_int),
                ("pairs", C.c_longlong), ("mt_trials", C.c_int)]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "libndt_oracle.so")
    if not os.path.exists(path):
        build(ref=False)
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int32)
    L.orc_vox_layout_compute.argtypes = [Cloud, C.c_float, C.c_float, C.c_float, C.POINTER(VoxLayout)]
    L.orc_vox_layout_compute.restype = C.c_int
    L.orc_vox_index.argtypes = [C.POINTER(VoxLayout), C.c_float, C.c_float, C.c_float]
    L.orc_vox_index.restype = C.c_int32
    L.orc_voxel_filter.argtypes = [Cloud, C.c_float, C.c_float, C.c_float, fp, ip, ip, C.POINTER(C.c_int)]
    L.orc_voxel_filter.restype = C.c_size_t
    L.orc_grid_build.argtypes = [Cloud, C.c_float, C.c_int, C.c_double]
    L.orc_grid_build.restype = C.c_void_p
    L.orc_grid_from_leaves.argtypes = [C.c_size_t, ip, dp, dp, dp, C.c_float]
    L.orc_grid_from_leaves.restype = C.c_void_p
    L.orc_grid_free.argtypes = [C.c_void_p]
    L.orc_grid_num_leaves.argtypes = [C.c_void_p]
    L.orc_grid_num_leaves.restype = C.c_size_t
    L.orc_grid_leaves.argtypes = [C.c_void_p]
    L.orc_grid_leaves.restype = C.POINTER(Leaf)
    L.orc_grid_layout.argtypes = [C.c_void_p]
    L.orc_grid_layout.restype = C.POINTER(VoxLayout)
    L.orc_grid_radius_search.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_double, ip, fp, C.c_int]
    L.orc_grid_radius_search.restype = C.c_int
    L.orc_euler_angles_012_f32.argtypes = [fp, fp]
    L.orc_pose_to_matrix_f32.argtypes = [dp, fp]
    L.orc_jacobi_svd_solve6.argtypes = [dp, dp, dp, dp]
    L.orc_jacobi_svd_solve6.restype = C.c_int
    L.orc_eig3_sym.argtypes = [dp, dp, dp]
    L.orc_inverse3.argtypes = [dp, dp]
    L.orc_gauss_constants.argtypes = [C.c_double, C.c_float, dp, dp]
    L.orc_transform_point_f32.argtypes = [fp, C.c_float, C.c_float, C.c_float, fp]
    L.orc_params_default.argtypes = [C.POINTER(Params)]
    L.orc_ndt_derivatives.argtypes = [C.c_void_p, C.POINTER(Params), Cloud, fp, dp, C.c_int, dp, dp,
                                      C.POINTER(C.c_longlong)]
    L.orc_ndt_derivatives.restype = C.c_double
    L.orc_ndt_align.argtypes = [C.c_void_p, C.POINTER(Params), Cloud, fp, fp, fp, C.POINTER(Result), dp, C.c_int]
    L.orc_ndt_align.restype = C.c_int
    L.orc_fitness_score.argtypes = [Cloud, Cloud, fp, C.c_double]
    L.orc_fitness_score.restype = C.c_double
    _LIB = L
    return L


def ref_lib():
    """oracle/_ref/libeigen_ref.so (real vendored Eigen) or None when not built."""
    global _REF
    if _REF is not None:
        return _REF
    path = os.path.join(_HERE, "_ref", "libeigen_ref.so")
    if not os.path.exists(path):
        return None
    R = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    fp = C.POINTER(C.c_float)
    R.ref_svd_solve6.argtypes = [dp, dp, dp, dp]
    R.ref_svd_solve6.restype = C.c_int
    R.ref_euler012.argtypes = [fp, fp]
    R.ref_pose_matrix.argtypes = [dp, fp]
    R.ref_leaf_finish.argtypes = [dp, dp, C.c_int, C.c_double, dp, dp, dp, dp]
    R.ref_leaf_finish.restype = C.c_int
    R.ref_inverse3.argtypes = [dp, dp]
    R.ref_transform_point.argtypes = [fp, fp, fp]
    _REF = R
    return R


_REFNDT = None


def refndt_lib():
    """oracle/_ref/libndt_manual_ref.so: the reference's own in-tree NDT compiled against stub headers."""
    global _REFNDT
    if _REFNDT is not None:
        return _REFNDT
    path = os.path.join(_HERE, "_ref", "libndt_manual_ref.so")
    if not os.path.exists(path):
        return None
    R = C.CDLL(path)
    dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
    R.refndt_new.argtypes = [C.c_float, C.c_double, C.c_double, C.c_int, C.c_double]
    R.refndt_new.restype = C.c_void_p
    R.refndt_free.argtypes = [C.c_void_p]
    R.refndt_set_target.argtypes = [C.c_void_p, fp, C.c_size_t]
    R.refndt_update.argtypes = [C.c_void_p, fp, C.c_size_t]
    R.refndt_grid_info.argtypes = [C.c_void_p, ip]
    R.refndt_voxel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp, dp]
    R.refndt_voxel.restype = C.c_int
    R.refndt_radius_search.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, ip, C.c_int]
    R.refndt_radius_search.restype = C.c_int
    R.refndt_voxel_coords.argtypes = [C.c_void_p, C.c_int, ip]
    R.refndt_set_source.argtypes = [C.c_void_p, fp, C.c_size_t]
    R.refndt_derivatives.argtypes = [C.c_void_p, fp, dp, C.c_int, dp, dp]
    R.refndt_derivatives.restype = C.c_double
    R.refndt_angle_tables.argtypes = [C.c_void_p, dp, dp, dp]
    R.refndt_align.argtypes = [C.c_void_p, fp, fp, ip, ip, dp, fp]
    _REFNDT = R
    return R


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def as_cloud(a):
    """numpy (n,4) packed or (n,8) PCL-layout float32 -> (Cloud struct, keepalive array)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (4, 8):
        raise ValueError("cloud must be (n,4) or (n,8) float32")
    ioff = 12 if a.shape[1] == 4 else 16
    return Cloud(a.ctypes.data, a.shape[0], a.shape[1] * 4, ioff), a


def params(res=1.0, step_size=0.1, trans_eps=0.01, max_iter=30, outlier_ratio=0.55, min_pts=6,
           eig_mult=0.01, pcl17_compat=1, intree_compat=0):
    return Params(res, step_size, trans_eps, outlier_ratio, max_iter, min_pts, eig_mult, pcl17_compat, intree_compat)


def vox_layout(cloud, lx, ly, lz):
    c, keep = as_cloud(cloud)
    L = VoxLayout()
    lib().orc_vox_layout_compute(c, lx, ly, lz, C.byref(L))
    return L


def voxel_indices(cloud, lx, ly, lz):
    """Per-point PCL voxel index (int32; -1 for non-finite points) + layout."""
    c, a = as_cloud(cloud)
    L = VoxLayout()
    ok = lib().orc_vox_layout_compute(c, lx, ly, lz, C.byref(L))
    out = np.full(a.shape[0], -1, np.int32)
    if ok:
        f = lib().orc_vox_index
        for i in range(a.shape[0]):
            if np.all(np.isfinite(a[i, :3])):
                out[i] = f(C.byref(L), a[i, 0], a[i, 1], a[i, 2])
    return out, L


def voxel_filter(cloud, lx, ly, lz):
    """-> (out (M,4) float32, idx (M,) int32, counts (M,) int32, overflow flag)."""
    c, a = as_cloud(cloud)
    n = a.shape[0]
    out = np.zeros((max(n, 1), 4), np.float32)
    idx = np.zeros(max(n, 1), np.int32)
    cnt = np.zeros(max(n, 1), np.int32)
    ov = C.c_int(0)
    m = lib().orc_voxel_filter(c, lx, ly, lz, _fp(out), _ip(idx), _ip(cnt), C.byref(ov))
    return out[:m].copy(), idx[:m].copy(), cnt[:m].copy(), bool(ov.value)


class Grid:
    """NDT target cells (pcl::VoxelGridCovariance restatement)."""

    def __init__(self, target, res=1.0, min_pts=6, eig_mult=0.01):
        if target is None:
            self.h = None
            self.res = res
            return
        c, self._keep = as_cloud(target)
        self.h = lib().orc_grid_build(c, res, min_pts, eig_mult)
        self.res = res

    @classmethod
    def from_leaves(cls, ijk, mean, icov, weight=None, res=1.0):
        """grid from externally supplied per-voxel statistics (absolute cell coordinates)."""
        g = cls(None, res)
        ijk = np.ascontiguousarray(ijk, np.int32); mean = np.ascontiguousarray(mean, np.float64)
        icov = np.ascontiguousarray(icov, np.float64).reshape(-1, 9)
        w = None if weight is None else np.ascontiguousarray(weight, np.float64)
        g.h = lib().orc_grid_from_leaves(len(ijk), _ip(ijk), _dp(mean), _dp(icov), _dp(w) if w is not None else None, res)
        return g

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_grid_free(self.h)
            self.h = None

    @property
    def layout(self):
        return lib().orc_grid_layout(self.h).contents

    def leaves(self):
        n = lib().orc_grid_num_leaves(self.h)
        if n == 0:
            return np.zeros(0, LEAF_DTYPE)
        assert C.sizeof(Leaf) == LEAF_DTYPE.itemsize, (C.sizeof(Leaf), LEAF_DTYPE.itemsize)
        ptr = lib().orc_grid_leaves(self.h)
        buf = (C.c_char * (n * C.sizeof(Leaf))).from_address(C.addressof(ptr.contents))
        return np.frombuffer(buf, dtype=LEAF_DTYPE, count=n).copy()

    def radius_search(self, q, radius=None, cap=64):
        slots = np.zeros(cap, np.int32)
        d2 = np.zeros(cap, np.float32)
        k = lib().orc_grid_radius_search(self.h, float(q[0]), float(q[1]), float(q[2]),
                                         float(self.res if radius is None else radius), _ip(slots), _fp(d2), cap)
        return slots[:k].copy(), d2[:k].copy()


def transform_points(T, xyz):
    """pcl::transformPointCloud float semantics, vectorised (numpy float32 ops round like C)."""
    T = np.asarray(T, np.float32).reshape(4, 4, order="F")
    x = xyz[:, 0].astype(np.float32)
    y = xyz[:, 1].astype(np.float32)
    z = xyz[:, 2].astype(np.float32)
    out = np.empty((xyz.shape[0], 3), np.float32)
    for r in range(3):
        out[:, r] = ((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3]
    return out


def pose_to_matrix(p):
    p = np.ascontiguousarray(p, np.float64)
    T = np.zeros(16, np.float32)
    lib().orc_pose_to_matrix_f32(_dp(p), _fp(T))
    return T.reshape(4, 4, order="F").copy()


def euler_angles(T):
    Tc = np.ascontiguousarray(np.asarray(T, np.float32).reshape(4, 4).flatten(order="F"))
    out = np.zeros(3, np.float32)
    lib().orc_euler_angles_012_f32(_fp(Tc), _fp(out))
    return out


def svd_solve6(H, b):
    Hc = np.ascontiguousarray(np.asarray(H, np.float64).reshape(6, 6).flatten(order="F"))
    bc = np.ascontiguousarray(b, np.float64)
    x = np.zeros(6)
    sv = np.zeros(6)
    rank = lib().orc_jacobi_svd_solve6(_dp(Hc), _dp(bc), _dp(x), _dp(sv))
    return x, sv, rank


def gauss_constants(outlier_ratio=0.55, res=1.0):
    d1 = C.c_double()
    d2 = C.c_double()
    lib().orc_gauss_constants(outlier_ratio, res, C.byref(d1), C.byref(d2))
    return d1.value, d2.value


def derivatives(grid, prm, src, pose6, trans_xyz=None, compute_hessian=True):
    """computeDerivatives at 6-vector pose; trans_xyz defaults to T(pose)*src in float."""
    c, a = as_cloud(src)
    p = np.ascontiguousarray(pose6, np.float64)
    if trans_xyz is None:
        trans_xyz = transform_points(pose_to_matrix(p), a[:, :3])
    t = np.ascontiguousarray(trans_xyz, np.float32)
    g = np.zeros(6)
    H = np.zeros(36)
    pairs = C.c_longlong(0)
    s = lib().orc_ndt_derivatives(grid.h, C.byref(prm), c, _fp(t), _dp(p), int(compute_hessian), _dp(g), _dp(H),
                                  C.byref(pairs))
    return s, g, H.reshape(6, 6, order="F").copy(), pairs.value


def align(grid, prm, src, guess, want_cloud=False, trace_cap=0):
    c, a = as_cloud(src)
    G = np.ascontiguousarray(np.asarray(guess, np.float32).reshape(4, 4).flatten(order="F"))
    pose = np.zeros(16, np.float32)
    res = Result()
    out = np.zeros((a.shape[0], 3), np.float32) if want_cloud else None
    trace = np.zeros((max(trace_cap, 1), 8)) if trace_cap else None
    lib().orc_ndt_align(grid.h, C.byref(prm), c, _fp(G), _fp(pose), _fp(out) if want_cloud else None,
                        C.byref(res), _dp(trace) if trace_cap else None, trace_cap)
    r = dict(pose=pose.reshape(4, 4, order="F").copy(), iterations=res.iterations, converged=bool(res.converged),
             score=res.score, trans_probability=res.trans_probability, p=np.array(res.p[:]), passes=res.passes,
             pairs=res.pairs, mt_trials=res.mt_trials)
    if want_cloud:
        r["cloud"] = out
    if trace_cap:
        r["trace"] = trace[:min(res.iterations, trace_cap)].copy()
    return r


def fitness_score(target, src, pose, max_range=np.finfo(np.float64).max):
    ct, kt = as_cloud(target)
    cs, ks = as_cloud(src)
    P = np.ascontiguousarray(np.asarray(pose, np.float32).reshape(4, 4).flatten(order="F"))
    return lib().orc_fitness_score(ct, cs, _fp(P), max_range)


def box_filter(cloud, edge):
    """pcl::CropBox as BoxFilter::Filter configures it (lidar_localization/src/models/cloud_filter/box_filter.cpp:27-37):
    identity box pose, min = (edge[0], edge[2], edge[4]), max = (edge[1], edge[3], edge[5]); a finite point is kept
    unless a coordinate is < min or > max (PCL 1.7 crop_box.hpp applyFilter), input order kept.  TEST ORACLE."""
    c = np.ascontiguousarray(cloud, dtype=np.float32)
    e = np.asarray(edge, np.float32)
    xyz = c[:, :3]
    fin = np.isfinite(xyz).all(axis=1)
    lo = np.array([e[0], e[2], e[4]], np.float32)
    hi = np.array([e[1], e[3], e[5]], np.float32)
    with np.errstate(invalid="ignore"):
        keep = fin & ~((xyz < lo).any(axis=1) | (xyz > hi).any(axis=1))
    return c[keep].copy()


def transform_cloud(cloud, T):
    """pcl::transformPointCloud on PointXYZI with a float 4x4 (front_end.cpp:402-407): xyz through transform_points
    (the float expression pinned against the vendored Eigen), intensity copied, non-finite points copied unchanged.
    TEST ORACLE."""
    c = np.array(cloud, dtype=np.float32, copy=True)
    fin = np.isfinite(c[:, :3]).all(axis=1)
    if fin.any():
        c[fin, :3] = transform_points(np.asarray(T, np.float32).reshape(4, 4), c[fin, :3])
    return c


def gauss2d_map_cells(local_map, origin, grid_resolution=0.8):
    """Matching::generateGauss2DMapCells + resetMapRange (lidar_localization/src/matching/matching.cpp:344-424) with the
    reference's C++ arithmetic (float unless an operand is double; grid_map_resolution_ is double, std::pow(float, int)
    is double).  Sequential over the points in input order.  TEST ORACLE (pure-Python loop: small maps only).
    -> dict(width, height, min_xyz, max_xyz, mu[w,h], sigma[w,h], cnt[w,h])"""
    f32 = np.float32
    res = max(0.1, float(grid_resolution))
    c = np.ascontiguousarray(local_map, dtype=np.float32)
    fin = np.isfinite(c[:, :3]).all(axis=1)
    pts = (c[fin, :3] - np.asarray(origin, f32)[None, :]).astype(f32)
    mn = np.full(3, np.finfo(f32).max, f32)
    mx = np.full(3, -np.finfo(f32).max, f32)
    if len(pts):
        mn = np.minimum(mn, pts.min(axis=0)); mx = np.maximum(mx, pts.max(axis=0))

    def cround(v):                      # std::round: half away from zero
        return float(np.floor(abs(v) + 0.5) * (1.0 if v >= 0 else -1.0))

    w = int(cround(float(f32(mx[0] - mn[0])) / res)) if len(pts) else 0
    h = int(cround(float(f32(mx[1] - mn[1])) / res)) if len(pts) else 0
    mu = np.zeros((w, h), f32); sg = np.zeros((w, h), f32); cnt = np.zeros((w, h), np.int32)
    cx = np.array([int(cround(float(f32(p - mn[0])) / res)) for p in pts[:, 0]], np.int64)
    cy = np.array([int(cround(float(f32(p - mn[1])) / res)) for p in pts[:, 1]], np.int64)
    for i in range(len(pts)):
        x, y = cx[i], cy[i]
        if x < 0 or y < 0 or x >= w or y >= h:
            continue
        z = f32(pts[i, 2]); m = mu[x, y]; s = sg[x, y]; n = int(cnt[x, y])
        if n == 0:
            mu[x, y] = z; sg[x, y] = 0; cnt[x, y] = 1
        else:
            mn_new = f32(f32(f32(f32(n) * m) + z) / f32(n + 1))
            a = float(f32(f32(n - 1) * s))
            b = float(f32(z - m)) ** 2
            cc = float(n + 1) * (float(f32(mn_new - m)) ** 2)
            d = float(f32(f32(f32(2) * f32(mn_new - m)) * f32(z - mn_new)))
            tot = ((a + b) + cc) + d
            sg[x, y] = f32(f32(tot) / f32(n))
            mu[x, y] = mn_new
            cnt[x, y] = n + 1
    return dict(width=w, height=h, min_xyz=mn, max_xyz=mx, mu=mu, sigma=sg, cnt=cnt, res=res)


def initial_yaw_angle(cells, scan, angle_size=270):
    """Matching::getInitialYawAngle (matching.cpp:267-308) on the grid of gauss2d_map_cells.  float sin/cos here are
    numpy's (the reference's are libm's): scores agree to float round-off, not bit for bit.  TEST ORACLE.
    -> (best yaw angle, probs[angle_size])"""
    f32 = np.float32
    s = np.ascontiguousarray(scan, dtype=np.float32)
    delta = f32(2 * np.pi / angle_size)
    w, h, res = cells["width"], cells["height"], cells["res"]
    mn = cells["min_xyz"]
    probs = np.zeros(angle_size, np.float64)
    x0, y0, z0 = s[:, 0], s[:, 1], s[:, 2]
    for i in range(angle_size):
        a = f32(delta * f32(i))
        c, sn = f32(np.cos(np.float64(a))), f32(np.sin(np.float64(a)))       # correctly rounded float cos / sin
        m22 = f32(f32(f32(1) - c) + c)
        with np.errstate(invalid="ignore"):
            x = ((c * x0 + (-sn) * y0) + f32(0) * z0) + f32(0)
            y = ((sn * x0 + c * y0) + f32(0) * z0) + f32(0)
            z = ((f32(0) * x0 + f32(0) * y0) + m22 * z0) + f32(0)
        fin = np.isfinite(x) & np.isfinite(y) & np.isfinite(z)
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            qx = (x - mn[0]).astype(np.float64) / res
            qy = (y - mn[1]).astype(np.float64) / res
            cx = (np.floor(np.abs(qx) + 0.5) * np.sign(qx)); cy = (np.floor(np.abs(qy) + 0.5) * np.sign(qy))
            ok = fin & (cx >= 0) & (cy >= 0) & (cx < w) & (cy < h)
            ix, iy = cx[ok].astype(np.int64), cy[ok].astype(np.int64)
            occ = cells["cnt"][ix, iy] > 0
            ix, iy = ix[occ], iy[occ]
            d = (z[ok][occ] - cells["mu"][ix, iy]).astype(np.float64)
            term = np.exp(-(d * d) / (f32(2) * cells["sigma"][ix, iy]).astype(np.float64))
        acc = 0.0
        for t in term:                      # sequential double sum in point order
            acc += t
        probs[i] = acc
    max_prob = -np.finfo(f32).max
    best = 0.0
    for it in range(angle_size):
        if probs[it] > max_prob:
            max_prob = f32(probs[it]); best = float(f32(f32(it) * delta))
    return best, probs


_REFMATCH = None


def refmatch_lib():
    """oracle/_ref/libmatching_ref.so: the reference's own matching.cpp (height grid + yaw scan) compiled against stub headers."""
    global _REFMATCH
    if _REFMATCH is not None:
        return _REFMATCH
    path = os.path.join(_HERE, "_ref", "libmatching_ref.so")
    if not os.path.exists(path):
        return None
    R = C.CDLL(path)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    R.refmatch_new.argtypes = [C.c_double]; R.refmatch_new.restype = C.c_void_p
    R.refmatch_free.argtypes = [C.c_void_p]
    R.refmatch_build.argtypes = [C.c_void_p, fp, C.c_size_t, fp]
    R.refmatch_info.argtypes = [C.c_void_p, ip, fp, fp]
    R.refmatch_cells.argtypes = [C.c_void_p, fp, fp, ip]
    R.refmatch_yaw.argtypes = [C.c_void_p, fp, C.c_size_t]; R.refmatch_yaw.restype = C.c_double
    _REFMATCH = R
    return R


def gauss2d_map_cells_reference(local_map, origin, grid_resolution=0.8, scans=()):
    """Run the reference's own Matching::generateGauss2DMapCells (+ getInitialYawAngle for every cloud in `scans`);
    only where oracle/_ref was built.  -> (cells dict like gauss2d_map_cells, [yaw angle per scan])"""
    R = refmatch_lib()
    if R is None:
        raise RuntimeError("oracle/_ref/libmatching_ref.so is not built (needs /root/reference)")
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    m = np.ascontiguousarray(local_map, np.float32)
    o = np.ascontiguousarray(origin, np.float32)
    h = R.refmatch_new(float(grid_resolution))
    try:
        R.refmatch_build(h, m.ctypes.data_as(fp), len(m), o.ctypes.data_as(fp))
        wh = np.zeros(2, np.int32); mn = np.zeros(3, np.float32); mx = np.zeros(3, np.float32)
        R.refmatch_info(h, wh.ctypes.data_as(ip), mn.ctypes.data_as(fp), mx.ctypes.data_as(fp))
        w, hh = int(wh[0]), int(wh[1])
        mu = np.zeros((w, hh), np.float32); sg = np.zeros((w, hh), np.float32); cnt = np.zeros((w, hh), np.int32)
        R.refmatch_cells(h, mu.ctypes.data_as(fp), sg.ctypes.data_as(fp), cnt.ctypes.data_as(ip))
        yaws = []
        for s in scans:
            s = np.ascontiguousarray(s, np.float32)
            yaws.append(R.refmatch_yaw(h, s.ctypes.data_as(fp), len(s)))
    finally:
        R.refmatch_free(h)
    return dict(width=w, height=hh, min_xyz=mn, max_xyz=mx, mu=mu, sigma=sg, cnt=cnt, res=max(0.1, float(grid_resolution))), yaws


# ------------------------------------------------------------------ scan de-skew (data_pretreat) -------------
_DESKEW = None


def deskew_ref_lib():
    """oracle/_ref/libdeskew_ref.so: the reference's own distortion_adjust.cpp compiled against stub headers."""
    global _DESKEW
    if _DESKEW is not None:
        return _DESKEW
    path = os.path.join(_HERE, "_ref", "libdeskew_ref.so")
    if not os.path.exists(path):
        return None
    R = C.CDLL(path)
    R.ref_distortion_adjust.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.POINTER(C.c_float), C.c_int]
    R.ref_distortion_adjust.restype = C.c_int
    _DESKEW = R
    return R


def distortion_adjust_reference(cloud, scan_period, linear_velocity, angular_velocity):
    """Run the reference's own DistortionAdjust::AdjustCloud (only where oracle/_ref was built)."""
    R = deskew_ref_lib()
    c = np.ascontiguousarray(cloud, dtype=np.float32)
    out = np.zeros_like(c)
    lin = np.ascontiguousarray(linear_velocity, np.float64); ang = np.ascontiguousarray(angular_velocity, np.float64)
    m = R.ref_distortion_adjust(_fp(c), len(c), float(scan_period), _dp(lin), _dp(ang), _fp(out), len(c))
    return out[:m].copy()


def distortion_adjust(cloud, scan_period, linear_velocity, angular_velocity):
    """DistortionAdjust::AdjustCloud (lidar_localization/src/models/scan_adjust/distortion_adjust.cpp:16-69) restated in
    float64 numpy (the reference computes in float through Eigen: agreement is to float round-off, ~1e-5 m):
      * the scan is rotated about z so that its first point has azimuth 0 (:22-29); velocity and angular rate are
        rotated by the SAME (not the inverse) matrix (:31-32);
      * point 0 is skipped (:34); a point with azimuth o in [0, 2 pi) is dropped inside the 5 degree sector around
        0 (:39-40); its time is o / (2 pi) * period - period / 2 (:42);
      * p' = Rz(wz t) Ry(wy t) Rx(wx t) p + v t (:48-51, UpdateMatrix :60-68); intensity is NOT carried over (:52-56);
      * the result is rotated back (:59).
    The sector test is evaluated in float like the reference.  TEST ORACLE.  -> (n_out, 4) float32"""
    c = np.ascontiguousarray(cloud, dtype=np.float32)
    if len(c) == 0:
        return np.zeros((0, 4), np.float32)
    f32 = np.float32
    start = np.arctan2(f32(c[0, 1]), f32(c[0, 0])).astype(f32)
    cs, sn = np.cos(np.float64(start)), np.sin(np.float64(start))
    R = np.array([[cs, -sn, 0.0], [sn, cs, 0.0], [0.0, 0.0, 1.0]])
    Rinv = R.T
    p = c[:, :3].astype(np.float64) @ Rinv.T
    vel = R @ np.asarray(linear_velocity, np.float64).astype(f32).astype(np.float64)
    rate = R @ np.asarray(angular_velocity, np.float64).astype(f32).astype(np.float64)
    pf = p.astype(f32)
    o = np.arctan2(pf[:, 1], pf[:, 0]).astype(f32)
    o = np.where(o < 0, (o.astype(np.float64) + 2.0 * np.pi).astype(f32), o)
    delete_space = f32(5.0 * np.pi / 180.0)
    keep = ~((o < delete_space) | ((2.0 * np.pi - o.astype(np.float64)) < np.float64(delete_space)))
    keep[0] = False
    t = (np.abs(o.astype(np.float64)) / np.float64(f32(2.0 * np.pi)) * np.float64(f32(scan_period)) - np.float64(f32(scan_period)) / 2.0)
    t = t.astype(f32).astype(np.float64)[keep]
    q = p[keep]
    ax, ay, az = rate[0] * t, rate[1] * t, rate[2] * t
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    # Rz Ry Rx applied to q
    x1, y1, z1 = q[:, 0], cx * q[:, 1] - sx * q[:, 2], sx * q[:, 1] + cx * q[:, 2]
    x2, y2, z2 = cy * x1 + sy * z1, y1, -sy * x1 + cy * z1
    x3, y3, z3 = cz * x2 - sz * y2, sz * x2 + cz * y2, z2
    adj = np.stack([x3, y3, z3], 1) + vel[None, :] * t[:, None]
    out = np.zeros((len(adj), 4), np.float32)
    out[:, :3] = (adj @ R.T).astype(f32)
    return out
