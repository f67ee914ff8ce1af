"""ctypes binding of the C ABI declared in include/b2ndt.h (lidar_slam_b200/_lib/libb2ndt.so).

The library is the product: there is no Python / CPU fallback.  Loading fails loudly when the shared
object is missing, and every entry point returns B2_ERR_CUDA when no sm_100 device is usable.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

B2_OK, B2_ERR_INVALID, B2_ERR_CUDA, B2_ERR_STATE, B2_ERR_CAPACITY = 0, -1, -2, -3, -4

# every symbol include/b2ndt.h declares (tests check that the library exports all of them)
EXPORTS = [
    "b2_last_error", "b2_kernel_launch_count", "b2_device_count",
    "b2ndt_params_default", "b2ndt_create", "b2ndt_destroy", "b2ndt_set_stream", "b2ndt_synchronize",
    "b2ndt_set_cluster", "b2ndt_set_target", "b2ndt_set_target_device", "b2ndt_target_info_get",
    "b2ndt_update_target", "b2ndt_update_target_device", "b2ndt_update_target_cloud",
    "b2ndt_target_leaves", "b2ndt_align", "b2ndt_align_ex", "b2ndt_align_batch", "b2ndt_align_batch_device",
    "b2ndt_derivatives", "b2ndt_fitness", "b2ndt_fitness_ex",
    "b2vf_create", "b2vf_destroy", "b2vf_set_stream", "b2vf_filter", "b2vf_filter_batch_device", "b2vf_filter_batch_append_device",
    "b2cloud_create", "b2cloud_destroy", "b2cloud_upload", "b2cloud_download", "b2cloud_size", "b2cloud_clear",
    "b2cloud_device_ptr", "b2cloud_append_transformed", "b2cloud_assemble", "b2cloud_box_filter", "b2cloud_remove_nan", "b2cloud_distortion_adjust", "b2vf_filter_cloud", "b2vf_ingest_filter_cloud",
    "b2ndt_set_target_cloud", "b2ndt_align_cloud",
    "b2_pcd_read", "b2_pcd_free", "b2_pcd_write_binary", "b2cloud_load_pcd", "b2cloud_save_pcd",
    "b2hmap_create", "b2hmap_destroy", "b2hmap_build", "b2hmap_info", "b2hmap_cells", "b2hmap_yaw_search", "b2hmap_pose_search",
]


class B2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libb2ndt error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("res", C.c_float), ("step_size", C.c_double), ("trans_eps", C.c_double),
                ("outlier_ratio", C.c_double), ("max_iter", C.c_int), ("min_pts", C.c_int),
                ("eig_mult", C.c_double), ("pcl17_compat", C.c_int)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("score", C.c_double),
                ("trans_probability", C.c_double), ("p", C.c_double * 6), ("passes", C.c_int32),
                ("mt_trials", C.c_int32), ("pairs", C.c_int64)]


RESULT_DTYPE = np.dtype([("iterations", "<i4"), ("converged", "<i4"), ("score", "<f8"),
                         ("trans_probability", "<f8"), ("p", "<f8", (6,)), ("passes", "<i4"),
                         ("mt_trials", "<i4"), ("pairs", "<i8")], align=True)


class TargetInfo(C.Structure):
    _fields_ = [("ok", C.c_int32), ("min_b", C.c_int32 * 3), ("div_b", C.c_int32 * 3), ("n_points", C.c_uint32),
                ("n_leaves", C.c_uint32), ("n_tree", C.c_uint32), ("inv_leaf", C.c_float),
                ("updates_incremental", C.c_uint32), ("updates_rebuilt", C.c_uint32)]


def lib_path():
    # B2NDT_LIB selects a tuning variant built by build.build_cuda(variant=...) (kernel-shape sweeps)
    return os.environ.get("B2NDT_LIB") or os.path.join(_build.LIBDIR, "libb2ndt.so")


def lib():
    """Load libb2ndt.so (built in-tree by lidar_slam_b200.build.build_cuda)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the registration path)" % path)
    L = C.CDLL(path)
    vp, sz, fp, dp = C.c_void_p, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_double)
    u32p, i32p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    L.b2_last_error.restype = C.c_char_p
    L.b2_kernel_launch_count.restype = C.c_uint64
    L.b2_device_count.restype = C.c_int
    L.b2ndt_params_default.argtypes = [C.POINTER(Params)]
    L.b2ndt_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
    L.b2ndt_destroy.argtypes = [vp]
    L.b2ndt_destroy.restype = None
    L.b2ndt_set_stream.argtypes = [vp, vp]
    L.b2ndt_synchronize.argtypes = [vp]
    L.b2ndt_set_cluster.argtypes = [vp, C.c_int, C.c_int]
    L.b2ndt_set_target.argtypes = [vp, vp, sz, sz, sz]
    L.b2ndt_set_target_device.argtypes = [vp, vp, sz]
    L.b2ndt_update_target.argtypes = [vp, vp, sz, sz, sz]
    L.b2ndt_update_target_device.argtypes = [vp, vp, sz]
    L.b2ndt_update_target_cloud.argtypes = [vp, vp]
    L.b2ndt_target_info_get.argtypes = [vp, C.POINTER(TargetInfo)]
    L.b2ndt_target_leaves.argtypes = [vp, i32p, i32p, fp, dp, dp]
    L.b2ndt_align.argtypes = [vp, vp, sz, sz, sz, fp, fp, C.POINTER(Result)]
    L.b2ndt_align_ex.argtypes = [vp, vp, sz, sz, sz, fp, fp, C.POINTER(Result), vp, sz, sz]
    L.b2ndt_align_batch.argtypes = [vp, vp, sz, sz, sz, u32p, sz, fp, fp, vp]
    L.b2ndt_align_batch_device.argtypes = [vp, vp, sz, vp, sz, vp, vp, vp]
    L.b2ndt_derivatives.argtypes = [vp, vp, sz, sz, sz, dp, dp, dp, dp, C.POINTER(C.c_int64)]
    L.b2ndt_fitness.argtypes = [vp, C.c_double, dp]
    L.b2ndt_fitness_ex.argtypes = [vp, vp, sz, sz, sz, fp, C.c_double, dp]
    L.b2vf_create.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.POINTER(vp)]
    L.b2vf_destroy.argtypes = [vp]
    L.b2vf_destroy.restype = None
    L.b2vf_set_stream.argtypes = [vp, vp]
    L.b2vf_filter.argtypes = [vp, vp, sz, sz, sz, vp, sz, sz, sz, C.POINTER(sz), i32p, i32p]
    L.b2vf_filter_batch_device.argtypes = [vp, vp, sz, u32p, sz, vp, vp]
    L.b2vf_filter_batch_append_device.argtypes = [vp, vp, sz, u32p, sz, vp, sz, vp, sz, vp]
    L.b2cloud_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.b2cloud_destroy.argtypes = [vp]
    L.b2cloud_destroy.restype = None
    L.b2cloud_upload.argtypes = [vp, vp, sz, sz, sz]
    L.b2cloud_download.argtypes = [vp, vp, sz, sz, sz, C.POINTER(sz)]
    L.b2cloud_size.argtypes = [vp, C.POINTER(sz)]
    L.b2cloud_clear.argtypes = [vp]
    L.b2cloud_device_ptr.argtypes = [vp, C.POINTER(vp)]
    L.b2cloud_append_transformed.argtypes = [vp, vp, fp]
    L.b2cloud_assemble.argtypes = [vp, C.POINTER(vp), fp, sz]
    L.b2cloud_box_filter.argtypes = [vp, fp, vp]
    L.b2cloud_remove_nan.argtypes = [vp, vp]
    L.b2cloud_distortion_adjust.argtypes = [vp, C.c_float, dp, dp, vp]
    L.b2vf_filter_cloud.argtypes = [vp, vp, vp]
    L.b2vf_ingest_filter_cloud.argtypes = [vp, vp, C.c_float, dp, dp, vp, vp]
    L.b2ndt_set_target_cloud.argtypes = [vp, vp]
    L.b2ndt_align_cloud.argtypes = [vp, vp, fp, fp, C.POINTER(Result), vp]
    L.b2_pcd_read.argtypes = [C.c_char_p, C.POINTER(fp), C.POINTER(sz)]
    L.b2_pcd_free.argtypes = [fp]
    L.b2_pcd_free.restype = None
    L.b2_pcd_write_binary.argtypes = [C.c_char_p, fp, sz]
    L.b2cloud_load_pcd.argtypes = [vp, C.c_char_p]
    L.b2cloud_save_pcd.argtypes = [vp, C.c_char_p]
    L.b2hmap_create.argtypes = [C.c_int, C.c_double, C.POINTER(vp)]
    L.b2hmap_destroy.argtypes = [vp]
    L.b2hmap_destroy.restype = None
    L.b2hmap_build.argtypes = [vp, vp, fp]
    L.b2hmap_info.argtypes = [vp, i32p, i32p, fp, fp]
    L.b2hmap_cells.argtypes = [vp, fp, fp, i32p]
    L.b2hmap_yaw_search.argtypes = [vp, vp, C.c_int, dp, dp]
    L.b2hmap_pose_search.argtypes = [vp, vp, C.c_int, fp, C.c_int, dp]
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise B2Error(rc, lib().b2_last_error().decode("utf-8", "replace"))


def launches():
    return int(lib().b2_kernel_launch_count())


def cloud_args(a):
    """numpy (n,4) packed or (n,8) PointXYZI-layout float32 -> (array, ptr, n, stride, ioff)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (4, 8):
        raise ValueError("cloud must be float32 of shape (n,4) or (n,8)")
    return a, a.ctypes.data, a.shape[0], a.shape[1] * 4, (12 if a.shape[1] == 4 else 16)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pose_to_colmajor(T):
    return np.ascontiguousarray(np.asarray(T, np.float32).reshape(4, 4).flatten(order="F"))


def colmajor_to_pose(v):
    return np.asarray(v, np.float32).reshape(4, 4, order="F").copy()


def pcd_read(path):
    """pcl::io::loadPCDFile for PointXYZI clouds -> numpy (n,4) float32 {x,y,z,intensity} (host only)."""
    p = C.POINTER(C.c_float)()
    n = C.c_size_t(0)
    check(lib().b2_pcd_read(os.fsencode(path), C.byref(p), C.byref(n)))
    try:
        out = np.ctypeslib.as_array(p, shape=(max(n.value, 1) * 4,))[:n.value * 4].reshape(-1, 4).copy() if n.value else np.zeros((0, 4), np.float32)
    finally:
        lib().b2_pcd_free(p)
    return out


def pcd_write_binary(path, cloud):
    """pcl::io::savePCDFileBinary of a PointXYZI cloud given as numpy (n,4) or (n,8) (host only)."""
    a = np.ascontiguousarray(cloud, dtype=np.float32)
    if a.ndim == 2 and a.shape[1] == 8:
        a = np.ascontiguousarray(np.concatenate([a[:, :3], a[:, 4:5]], axis=1))
    check(lib().b2_pcd_write_binary(os.fsencode(path), _fp(a), a.shape[0]))
