"""Drive the product's host-compiled Newton / More-Thuente controller (csrc/b2_ndt_math.cuh via
libb2hostcheck.so) with derivative passes evaluated by the CPU oracle.  Test helper only."""
import ctypes as C
import os

import numpy as np

from lidar_slam_b200 import build
from oracle import oracle as O

_HC = None


def hc():
    global _HC
    if _HC is None:
        path = build.build_hostcheck()
        L = C.CDLL(path)
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.hc_pose_to_matrix.argtypes = [dp, fp]
        L.hc_euler.argtypes = [fp, fp]
        L.hc_newton_solve6.argtypes = [dp, dp, dp, C.c_int]
        L.hc_svd_solve6.argtypes = [dp, dp, dp]
        L.hc_svd_solve6.restype = C.c_int
        L.hc_lu_solve6.argtypes = [dp, dp, dp]
        L.hc_lu_solve6.restype = C.c_double
        L.hc_leaf_finish.argtypes = [dp, dp, C.c_int, C.c_int, C.c_double, dp, dp, dp, dp]
        L.hc_leaf_finish.restype = C.c_int
        L.hc_angle_derivatives.argtypes = [dp, dp, dp]
        L.hc_ctl_new.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_float]
        L.hc_ctl_new.restype = C.c_void_p
        L.hc_ctl_free.argtypes = [C.c_void_p]
        L.hc_ctl_start.argtypes = [C.c_void_p, fp, C.c_double]
        L.hc_ctl_step.argtypes = [C.c_void_p, dp]
        L.hc_ctl_step.restype = C.c_int
        L.hc_ctl_request.argtypes = [C.c_void_p, fp, dp, ip]
        L.hc_ctl_result.argtypes = [C.c_void_p, fp, dp, dp, dp, ip, ip, ip, ip]
        _HC = L
    return _HC


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def pack_acc(score, g, H, pairs):
    acc = np.zeros(29)
    acc[0] = score
    acc[1:7] = g
    k = 7
    for i in range(6):
        for j in range(i, 6):
            acc[k] = H[i, j]
            k += 1
    acc[28] = pairs
    return acc


def controller_align(grid, prm, src, guess, force_svd=0, max_passes=2000):
    """Run the product controller; passes evaluated with oracle.derivatives.  Returns dict like oracle.align."""
    L = hc()
    d1, d2 = O.gauss_constants(prm.outlier_ratio, prm.res)
    h = L.hc_ctl_new(d1, d2, prm.step_size, prm.trans_eps, prm.max_iter, prm.pcl17_compat, force_svd, prm.res)
    G = np.ascontiguousarray(np.asarray(guess, np.float32).reshape(4, 4).flatten(order="F"))
    src = np.ascontiguousarray(src, np.float32)
    L.hc_ctl_start(h, _fp(G), float(src.shape[0]))
    T = np.zeros(16, np.float32)
    x = np.zeros(6)
    hess = C.c_int(0)
    npass = 0
    while True:
        L.hc_ctl_request(h, _fp(T), _dp(x), C.byref(hess))
        trans = O.transform_points(T.reshape(4, 4, order="F"), src[:, :3])
        s, g, H, pairs = O.derivatives(grid, prm, src, x, trans_xyz=trans, compute_hessian=bool(hess.value))
        acc = pack_acc(s, g, H, pairs)
        npass += 1
        if not L.hc_ctl_step(h, _dp(acc)) or npass >= max_passes:
            break
    fT = np.zeros(16, np.float32)
    p = np.zeros(6)
    score, tp = C.c_double(), C.c_double()
    it, conv, passes, mt = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    L.hc_ctl_result(h, _fp(fT), _dp(p), C.byref(score), C.byref(tp), C.byref(it), C.byref(conv), C.byref(passes), C.byref(mt))
    L.hc_ctl_free(h)
    return dict(pose=fT.reshape(4, 4, order="F").copy(), p=p, score=score.value, trans_probability=tp.value,
                iterations=it.value, converged=bool(conv.value), passes=passes.value, mt_trials=mt.value)
