"""Synthetic HDL-64-shaped workload: ctypes wrapper of csrc/synth_hdl64.c.

Workload generation only (tests + bench); not part of the registration path.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None
RAYS = 64 * 2083
SCENE_SEED = 0xB200


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_build.LIBDIR, "libb2synth.so")
        if not os.path.exists(path):
            _build.build_synth()
        L = C.CDLL(path)
        L.synth_scene_create.argtypes = [C.c_uint64, C.c_double]
        L.synth_scene_create.restype = C.c_void_p
        L.synth_scene_free.argtypes = [C.c_void_p]
        L.synth_path_pose.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        L.synth_path_length.argtypes = [C.c_void_p]
        L.synth_path_length.restype = C.c_double
        L.synth_pose_to_matrix.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.synth_scan.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_double), C.c_double, C.c_int,
                                 C.POINTER(C.c_float)]
        L.synth_scan.restype = C.c_size_t
        L.synth_scans.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.POINTER(C.c_size_t), C.c_int]
        L.synth_map.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.POINTER(C.c_float), C.c_int]
        L.synth_map.restype = C.c_size_t
        L.synth_point_in_box.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.synth_point_in_box.restype = C.c_int
        _LIB = L
    return _LIB


def pose6_to_matrix(p):
    """(x,y,z,roll,pitch,yaw) -> 4x4 float64, R = Rx(roll) Ry(pitch) Rz(yaw) (the reference's convention)."""
    p = np.ascontiguousarray(p, np.float64)
    T = np.zeros(16)
    _lib().synth_pose_to_matrix(p.ctypes.data_as(C.POINTER(C.c_double)), T.ctypes.data_as(C.POINTER(C.c_double)))
    return T.reshape(4, 4)


class Scene:
    """Procedural street scene + L-shaped drive of 2*leg metres."""

    def __init__(self, seed=SCENE_SEED, leg=500.0):
        self.h = _lib().synth_scene_create(seed, leg)
        self.leg = leg

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                _lib().synth_scene_free(h)
            except Exception:       # interpreter shutdown: module globals may already be gone
                pass

    @property
    def path_length(self):
        return _lib().synth_path_length(self.h)

    def path_pose(self, s):
        p = np.zeros(6)
        _lib().synth_path_pose(self.h, float(s), p.ctypes.data_as(C.POINTER(C.c_double)))
        return p

    def point_in_box(self, x, y, margin=0.0):
        """True when (x, y) is inside the footprint of a building or car grown by `margin` metres."""
        return bool(_lib().synth_point_in_box(self.h, float(x), float(y), float(margin)))

    def scan(self, frame_id, pose6, world=False):
        """One scan in the sensor frame -> (n,4) float32 {x,y,z,intensity}."""
        out = np.empty((RAYS, 4), np.float32)
        p = np.ascontiguousarray(pose6, np.float64)
        n = _lib().synth_scan(self.h, int(frame_id), p.ctypes.data_as(C.POINTER(C.c_double)), 1.0, int(world),
                              out.ctypes.data_as(C.POINTER(C.c_float)))
        return out[:n].copy()

    def scans(self, frame_ids, poses6, nthreads=None):
        """Many scans -> list of (n_i,4) arrays (generated on worker threads)."""
        ids = np.ascontiguousarray(frame_ids, np.uint64)
        P = np.ascontiguousarray(poses6, np.float64).reshape(-1, 6)
        n = len(ids)
        buf = np.empty((n, RAYS, 4), np.float32)
        cnt = np.zeros(n, np.uint64)
        nt = nthreads or min(64, os.cpu_count() or 1)
        _lib().synth_scans(self.h, ids.ctypes.data_as(C.POINTER(C.c_uint64)), P.ctypes.data_as(C.POINTER(C.c_double)),
                           n, RAYS, buf.ctypes.data_as(C.POINTER(C.c_float)),
                           cnt.ctypes.data_as(C.POINTER(C.c_size_t)), nt)
        return [buf[i, :int(cnt[i])].copy() for i in range(n)]

    def make_map(self, n_points=1_000_000, spacing=2.0):
        out = np.empty((n_points, 4), np.float32)
        n = _lib().synth_map(self.h, n_points, spacing, out.ctypes.data_as(C.POINTER(C.c_float)), 1)
        return out[:n].copy()


def perturb_pose(pose6, rng, dt=0.5, dr_deg=2.0):
    """truth o perturbation: U[-dt,dt] m per axis, U[-dr,dr] deg per angle (config 4)."""
    d = np.concatenate([rng.uniform(-dt, dt, 3), np.deg2rad(rng.uniform(-dr_deg, dr_deg, 3))])
    return np.asarray(pose6) + d
