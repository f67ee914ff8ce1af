"""Python mirror of the reference's plug-in interface for the registration hot path.

Same class / method names, argument meaning and error behaviour as
  lidar_localization/include/lidar_localization/models/registration/registration_interface.hpp:14-24
  lidar_localization/src/models/registration/ndt_registration.cpp:12-66
  lidar_localization/include/lidar_localization/models/cloud_filter/cloud_filter_interface.hpp:13-18
  lidar_localization/src/models/cloud_filter/voxel_filter.cpp:12-41
so the parity tests read like the reference's call sites.  All work is done by libb2ndt.so through
the C ABI of include/b2ndt.h; nothing here computes on the CPU except the final float transform of the
source cloud that fills `result_cloud` (pcl::Registration::align's output), and there is no fallback.

Clouds are numpy float32 arrays: (n, 8) in pcl::PointXYZI memory layout
{x, y, z, 1, intensity, 0, 0, 0} (cloud_data.hpp:35) or (n, 4) packed {x, y, z, intensity}.
Poses are 4x4 float32 (row/column indexed like Eigen::Matrix4f).
"""
import ctypes as C

import numpy as np

from . import capi

DBL_MAX = float(np.finfo(np.float64).max)


def to_xyzi8(cloud4):
    """(n,4) packed -> (n,8) pcl::PointXYZI layout."""
    c = np.asarray(cloud4, np.float32)
    out = np.zeros((c.shape[0], 8), np.float32)
    out[:, :3] = c[:, :3]
    out[:, 3] = 1.0
    out[:, 4] = c[:, 3]
    return out


def transform_cloud(cloud, T):
    """pcl::transformPointCloud float semantics: x' = ((m00 x + m01 y) + m02 z) + m03."""
    c = np.array(cloud, dtype=np.float32, copy=True)
    T = np.asarray(T, np.float32).reshape(4, 4)
    x, y, z = c[:, 0].copy(), c[:, 1].copy(), c[:, 2].copy()
    for r in range(3):
        c[:, r] = ((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3]
    return c


class DeviceCloud:
    """A point cloud resident in HBM (b2cloud of include/b2ndt.h): what CloudData::CLOUD_PTR becomes when the
    callers either side of the hot path keep their clouds on the device."""

    def __init__(self, cloud=None, device=0):
        self._h = C.c_void_p()
        capi.check(capi.lib().b2cloud_create(int(device), C.byref(self._h)))
        if cloud is not None:
            self.Upload(cloud)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and capi is not None and getattr(capi, "_LIB", None) is not None:
            try:
                capi._LIB.b2cloud_destroy(h)
            except Exception:
                pass
            self._h = None

    def __len__(self):
        n = C.c_size_t(0)
        capi.check(capi.lib().b2cloud_size(self._h, C.byref(n)))
        return int(n.value)

    def Upload(self, cloud):
        a, ptr, n, stride, ioff = capi.cloud_args(cloud)
        capi.check(capi.lib().b2cloud_upload(self._h, ptr, n, stride, ioff))
        return self

    def Download(self, layout=4):
        """-> numpy (n,4) packed or (n,8) PointXYZI layout."""
        n = len(self)
        out = np.zeros((n, layout), np.float32)
        m = C.c_size_t(0)
        capi.check(capi.lib().b2cloud_download(self._h, out.ctypes.data, n, layout * 4, 12 if layout == 4 else 16, C.byref(m)))
        return out[:m.value]

    def Clear(self):
        capi.check(capi.lib().b2cloud_clear(self._h))

    def DevicePtr(self):
        p = C.c_void_p()
        capi.check(capi.lib().b2cloud_device_ptr(self._h, C.byref(p)))
        return p.value

    def LoadPCD(self, path):
        """pcl::io::loadPCDFile(path, *cloud) (matching.cpp:155) straight into HBM."""
        import os
        capi.check(capi.lib().b2cloud_load_pcd(self._h, os.fsencode(path)))
        return self

    def SavePCD(self, path):
        """pcl::io::savePCDFileBinary(path, *cloud) (back_end.cpp:194)."""
        import os
        capi.check(capi.lib().b2cloud_save_pcd(self._h, os.fsencode(path)))

    def RemoveNaN(self, dst=None):
        """pcl::removeNaNFromPointCloud (front_end.cpp:92): finite points of *this, order kept -> dst (new cloud if
        None; dst may be this cloud: in place, the device pointer changes)."""
        if dst is None:
            dst = DeviceCloud()
        capi.check(capi.lib().b2cloud_remove_nan(self._h, dst._h))
        return dst

    def AppendTransformed(self, src, pose):
        """*this += pcl::transformPointCloud(src, pose)  (front_end.cpp:402-407)."""
        T = capi.pose_to_colmajor(pose)
        capi.check(capi.lib().b2cloud_append_transformed(self._h, src._h, capi._fp(T)))
        return self


    def Assemble(self, clouds, poses):
        """*this = concat_k pcl::transformPointCloud(clouds[k], poses[k])  (the local-map loop of front_end.cpp:398-407)
        in one launch; same points and order as Clear() + AppendTransformed per key frame."""
        K = len(clouds)
        hs = (C.c_void_p * max(K, 1))(*[c._h for c in clouds])
        P = np.ascontiguousarray(np.concatenate([capi.pose_to_colmajor(p) for p in poses]) if K else np.zeros(16, np.float32), np.float32)
        capi.check(capi.lib().b2cloud_assemble(self._h, hs, capi._fp(P), K))
        return self


class RegistrationInterface:
    def SetInputTarget(self, input_target):
        raise NotImplementedError

    def ScanMatch(self, input_source, predict_pose):
        raise NotImplementedError

    def GetFitnessScore(self):
        raise NotImplementedError


class CloudFilterInterface:
    def Filter(self, input_cloud):
        raise NotImplementedError


class NDTRegistration(RegistrationInterface):
    """NDTRegistration(res, step_size, trans_eps, max_iter) or NDTRegistration(node) with a dict holding
    the YAML keys res / step_size / trans_eps / max_iter (ndt_registration.cpp:12-27)."""

    def __init__(self, res, step_size=None, trans_eps=None, max_iter=None, device=0, pcl17_compat=True,
                 outlier_ratio=0.55, min_pts=6, eig_mult=0.01):
        if isinstance(res, dict):
            node = res
            res, step_size, trans_eps, max_iter = (float(node["res"]), float(node["step_size"]),
                                                   float(node["trans_eps"]), int(node["max_iter"]))
        L = capi.lib()
        p = capi.Params()
        L.b2ndt_params_default(C.byref(p))
        # the reference passes float values into setStepSize(double) / setTransformationEpsilon(double)
        p.res = float(np.float32(res))
        p.step_size = float(np.float32(step_size))
        p.trans_eps = float(np.float32(trans_eps))
        p.max_iter = int(max_iter)
        p.outlier_ratio = outlier_ratio
        p.min_pts = min_pts
        p.eig_mult = eig_mult
        p.pcl17_compat = 1 if pcl17_compat else 0
        self.params = p
        self._h = C.c_void_p()
        capi.check(L.b2ndt_create(C.byref(p), int(device), C.byref(self._h)))
        self.device = device
        self.last_result = None
        self._keep = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and capi is not None and getattr(capi, "_LIB", None) is not None:
            try:
                capi._LIB.b2ndt_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- reference surface -------------------------------------------------------------------
    def SetInputTarget(self, input_target):
        a, ptr, n, stride, ioff = capi.cloud_args(input_target)
        capi.check(capi.lib().b2ndt_set_target(self._h, ptr, n, stride, ioff))
        return True

    def ScanMatch(self, input_source, predict_pose, want_cloud=True):
        """-> (True, result_cloud, result_pose).  result_cloud = source transformed by the final pose."""
        a, ptr, n, stride, ioff = capi.cloud_args(input_source)
        g = capi.pose_to_colmajor(predict_pose)
        out = np.zeros(16, np.float32)
        res = capi.Result()
        cloud = None
        if want_cloud == "device":
            # result cloud filled by the library from the device-resident source (b2ndt_align_ex), packed x y z intensity
            cloud = np.zeros((n, 4), np.float32)
            capi.check(capi.lib().b2ndt_align_ex(self._h, ptr, n, stride, ioff, capi._fp(g), capi._fp(out), C.byref(res),
                                                 cloud.ctypes.data_as(C.c_void_p), 16, 12))
        else:
            capi.check(capi.lib().b2ndt_align(self._h, ptr, n, stride, ioff, capi._fp(g), capi._fp(out), C.byref(res)))
        pose = capi.colmajor_to_pose(out)
        self.last_result = dict(iterations=res.iterations, converged=bool(res.converged), score=res.score,
                                trans_probability=res.trans_probability, p=np.array(res.p[:]), passes=res.passes,
                                mt_trials=res.mt_trials, pairs=res.pairs)
        if want_cloud is True:
            cloud = transform_cloud(a, pose)
        return True, cloud, pose

    def GetFitnessScore(self, max_range=DBL_MAX):
        v = C.c_double()
        capi.check(capi.lib().b2ndt_fitness(self._h, max_range, C.byref(v)))
        return float(np.float32(v.value)) if max_range == DBL_MAX else v.value

    # ---- extensions over the same C ABI --------------------------------------------------------
    def GetFitnessScoreFor(self, source, pose, max_range=DBL_MAX):
        a, ptr, n, stride, ioff = capi.cloud_args(source)
        v = C.c_double()
        P = capi.pose_to_colmajor(pose)
        capi.check(capi.lib().b2ndt_fitness_ex(self._h, ptr, n, stride, ioff, capi._fp(P), max_range, C.byref(v)))
        return v.value

    def ScanMatchBatch(self, sources, predict_poses):
        """Many independent ScanMatch calls in one launch.  sources: list of clouds (same layout) or one
        cloud shared by all guesses (relocalisation hypotheses).  -> (poses (B,4,4), results recarray)."""
        poses = np.asarray(predict_poses, np.float32).reshape(-1, 4, 4)
        B = poses.shape[0]
        g = np.ascontiguousarray(poses.transpose(0, 2, 1).reshape(B, 16))
        out = np.zeros((B, 16), np.float32)
        res = np.zeros(B, capi.RESULT_DTYPE)
        L = capi.lib()
        if isinstance(sources, (list, tuple)):
            assert len(sources) == B
            cat = np.ascontiguousarray(np.concatenate([np.asarray(s, np.float32) for s in sources], axis=0))
            off = np.zeros(B + 1, np.uint32)
            off[1:] = np.cumsum([len(s) for s in sources])
            a, ptr, n, stride, ioff = capi.cloud_args(cat)
            capi.check(L.b2ndt_align_batch(self._h, ptr, n, stride, ioff, off.ctypes.data_as(C.POINTER(C.c_uint32)), B,
                                           capi._fp(g), capi._fp(out), res.ctypes.data))
        else:
            a, ptr, n, stride, ioff = capi.cloud_args(sources)
            capi.check(L.b2ndt_align_batch(self._h, ptr, n, stride, ioff, None, B, capi._fp(g), capi._fp(out),
                                           res.ctypes.data))
        return out.reshape(B, 4, 4).transpose(0, 2, 1).copy(), res

    def Derivatives(self, source, pose6):
        """One computeDerivatives pass at a 6-vector pose -> (score, grad(6), H(6,6), pairs)."""
        a, ptr, n, stride, ioff = capi.cloud_args(source)
        p = np.ascontiguousarray(pose6, np.float64)
        score = C.c_double()
        g = np.zeros(6)
        H = np.zeros(36)
        pairs = C.c_int64()
        capi.check(capi.lib().b2ndt_derivatives(self._h, ptr, n, stride, ioff, capi._dp(p), C.byref(score), capi._dp(g),
                                                capi._dp(H), C.byref(pairs)))
        return score.value, g, H.reshape(6, 6, order="F").copy(), pairs.value

    def UpdateInputTarget(self, new_cloud):
        """NormalDistributionsTransform::updateVoxelGrid(new_cloud) of the reference's in-tree NDT
        (ndt_registration_manual/NormalDistributionsTransform.cpp:968-972): add a cloud to the target without a full
        rebuild.  Same target as SetInputTarget(old ++ new), bit for bit.  Accepts a host cloud or a DeviceCloud."""
        if hasattr(new_cloud, "_h"):
            capi.check(capi.lib().b2ndt_update_target_cloud(self._h, new_cloud._h))
        else:
            a, ptr, n, stride, ioff = capi.cloud_args(new_cloud)
            capi.check(capi.lib().b2ndt_update_target(self._h, ptr, n, stride, ioff))
        return True

    def TargetInfo(self):
        info = capi.TargetInfo()
        capi.check(capi.lib().b2ndt_target_info_get(self._h, C.byref(info)))
        return dict(ok=bool(info.ok), min_b=list(info.min_b), div_b=list(info.div_b), n_points=info.n_points,
                    n_leaves=info.n_leaves, n_tree=info.n_tree, updates_incremental=info.updates_incremental,
                    updates_rebuilt=info.updates_rebuilt)

    def TargetLeaves(self):
        V = self.TargetInfo()["n_leaves"]
        idx = np.zeros(V, np.int32)
        n = np.zeros(V, np.int32)
        cen = np.zeros((V, 4), np.float32)
        mean = np.zeros((V, 3))
        icov = np.zeros((V, 9))
        if V:
            capi.check(capi.lib().b2ndt_target_leaves(self._h, idx.ctypes.data_as(C.POINTER(C.c_int32)),
                                                      n.ctypes.data_as(C.POINTER(C.c_int32)), capi._fp(cen),
                                                      capi._dp(mean), capi._dp(icov)))
        return dict(idx=idx, n=n, centroid=cen, mean=mean, icov=icov)

    def SetCluster(self, single_match_ctas, batch_ctas=0):
        capi.check(capi.lib().b2ndt_set_cluster(self._h, single_match_ctas, batch_ctas))

    def SetStream(self, cuda_stream_ptr):
        capi.check(capi.lib().b2ndt_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def Synchronize(self):
        capi.check(capi.lib().b2ndt_synchronize(self._h))

    # device-resident clouds (DeviceCloud): same semantics as SetInputTarget / ScanMatch, no host round trip
    def SetInputTargetCloud(self, target):
        capi.check(capi.lib().b2ndt_set_target_cloud(self._h, target._h))
        return True

    def ScanMatchCloud(self, source, predict_pose, result_cloud=None):
        """-> (True, result_cloud (DeviceCloud or None), result_pose)."""
        g = capi.pose_to_colmajor(predict_pose)
        out = np.zeros(16, np.float32)
        res = capi.Result()
        capi.check(capi.lib().b2ndt_align_cloud(self._h, source._h, capi._fp(g), capi._fp(out), C.byref(res),
                                                result_cloud._h if result_cloud is not None else None))
        pose = capi.colmajor_to_pose(out)
        self.last_result = dict(iterations=res.iterations, converged=bool(res.converged), score=res.score,
                                trans_probability=res.trans_probability, p=np.array(res.p[:]), passes=res.passes,
                                mt_trials=res.mt_trials, pairs=res.pairs)
        return True, result_cloud, pose

    # device-resident entry points (torch tensors' data_ptr()); asynchronous on the handle's stream
    def SetInputTargetDevice(self, d_ptr, n):
        capi.check(capi.lib().b2ndt_set_target_device(self._h, C.c_void_p(d_ptr), n))
        return True

    def ScanMatchBatchDevice(self, d_src, n_total, d_offsets, B, d_guesses, d_poses_out, d_results=None):
        capi.check(capi.lib().b2ndt_align_batch_device(self._h, C.c_void_p(d_src), n_total,
                                                       C.c_void_p(d_offsets) if d_offsets else None, B,
                                                       C.c_void_p(d_guesses), C.c_void_p(d_poses_out),
                                                       C.c_void_p(d_results) if d_results else None))


def src_device(cloud):
    return getattr(cloud, "device", 0)


class BoxFilter(CloudFilterInterface):
    """BoxFilter(node) with node["box_filter_size"] = [min_x, max_x, min_y, max_y, min_z, max_z] or BoxFilter(size)
    (box_filter.cpp:12-24); SetSize / SetOrigin / GetEdge as box_filter.cpp:39-75; Filter = pcl::CropBox."""

    def __init__(self, size=None, device=0):
        if isinstance(size, dict):
            size = size["box_filter_size"]
        self.size_ = [0.0] * 6 if size is None else [float(np.float32(v)) for v in size]
        self.origin_ = [0.0, 0.0, 0.0]
        self.edge_ = [0.0] * 6
        self.device = int(device)
        self.CalculateEdge()

    def SetSize(self, size):
        self.size_ = [float(np.float32(v)) for v in size]
        self.CalculateEdge()

    def SetOrigin(self, origin):
        self.origin_ = [float(np.float32(v)) for v in origin]
        self.CalculateEdge()

    def CalculateEdge(self):
        for i in range(3):
            self.edge_[2 * i] = float(np.float32(self.size_[2 * i]) + np.float32(self.origin_[i]))
            self.edge_[2 * i + 1] = float(np.float32(self.size_[2 * i + 1]) + np.float32(self.origin_[i]))

    def GetEdge(self):
        return list(self.edge_)

    def FilterCloud(self, src, dst=None):
        """Device clouds in, device cloud out (order kept); dst may be src (in place)."""
        if dst is None:
            dst = DeviceCloud(device=self.device)
        e = np.asarray(self.edge_, np.float32)
        capi.check(capi.lib().b2cloud_box_filter(src._h, capi._fp(e), dst._h))
        return dst

    def Filter(self, input_cloud):
        """-> (True, cropped cloud) in the layout of the input."""
        a = np.asarray(input_cloud)
        out = self.FilterCloud(DeviceCloud(input_cloud, device=self.device)).Download(layout=a.shape[1])
        return True, out


class VoxelFilter(CloudFilterInterface):
    """VoxelFilter(lx, ly, lz) or VoxelFilter(node) with node["leaf_size"] = [lx, ly, lz]
    (voxel_filter.cpp:12-23)."""

    def __init__(self, leaf_size_x, leaf_size_y=None, leaf_size_z=None, device=0):
        if isinstance(leaf_size_x, dict):
            leaf_size_x, leaf_size_y, leaf_size_z = [float(v) for v in leaf_size_x["leaf_size"]]
        self._h = C.c_void_p()
        capi.check(capi.lib().b2vf_create(float(leaf_size_x), float(leaf_size_y), float(leaf_size_z), int(device),
                                          C.byref(self._h)))
        self.leaf = (float(leaf_size_x), float(leaf_size_y), float(leaf_size_z))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and capi is not None and getattr(capi, "_LIB", None) is not None:
            try:
                capi._LIB.b2vf_destroy(h)
            except Exception:
                pass
            self._h = None

    def Filter(self, input_cloud, with_info=False):
        """-> (True, filtered_cloud) in the layout of the input; with_info adds (voxel idx, counts)."""
        a, ptr, n, stride, ioff = capi.cloud_args(input_cloud)
        out = np.empty_like(a)                  # the library writes every byte of the records it returns
        m = C.c_size_t(0)
        ip = C.POINTER(C.c_int32)
        idx = np.empty(max(n, 1), np.int32) if with_info else None
        cnt = np.empty(max(n, 1), np.int32) if with_info else None
        capi.check(capi.lib().b2vf_filter(self._h, ptr, n, stride, ioff, out.ctypes.data, n, stride, ioff, C.byref(m),
                                          idx.ctypes.data_as(ip) if with_info else None,
                                          cnt.ctypes.data_as(ip) if with_info else None))
        M = m.value
        if with_info:
            return True, out[:M].copy(), idx[:M].copy(), cnt[:M].copy()
        return True, out[:M].copy()

    def SetStream(self, cuda_stream_ptr):
        capi.check(capi.lib().b2vf_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def FilterCloud(self, src, dst=None):
        """Device cloud in, device cloud out; dst may be src (in place, matching.cpp:158)."""
        if dst is None:
            dst = DeviceCloud(device=src_device(src))
        capi.check(capi.lib().b2vf_filter_cloud(self._h, src._h, dst._h))
        return dst

    def IngestFilterCloud(self, src, filtered=None, ingested=None, scan_period=0.1, linear_velocity=None, angular_velocity=None):
        """Fused ingest of a raw device scan: de-skew (when velocities are given) + NaN removal + this filter in one
        pipeline run (b2vf_ingest_filter_cloud).  -> (filtered, ingested); `ingested` keeps src's size with NaN points
        where a point was dropped."""
        if filtered is None:
            filtered = DeviceCloud(device=src_device(src))
        lin = ang = None
        if linear_velocity is not None:
            lin = np.ascontiguousarray(linear_velocity, np.float64)
            ang = np.ascontiguousarray(angular_velocity, np.float64)
        dp = C.POINTER(C.c_double)
        capi.check(capi.lib().b2vf_ingest_filter_cloud(self._h, src._h, float(scan_period),
                                                       lin.ctypes.data_as(dp) if lin is not None else None,
                                                       ang.ctypes.data_as(dp) if ang is not None else None,
                                                       filtered._h, ingested._h if ingested is not None else None))
        return filtered, ingested

    def FilterBatchDevice(self, d_in, n_total, h_offsets, d_out, d_out_offsets):
        off = np.ascontiguousarray(h_offsets, np.uint32)
        capi.check(capi.lib().b2vf_filter_batch_device(self._h, C.c_void_p(d_in), n_total,
                                                       off.ctypes.data_as(C.POINTER(C.c_uint32)), len(off) - 1,
                                                       C.c_void_p(d_out), C.c_void_p(d_out_offsets)))

    def FilterBatchAppendDevice(self, d_in, n_total, h_offsets, d_out, out_capacity, d_out_offsets, first_index, d_cursor):
        """Batched ingest: filter B device clouds and append the results behind the ones earlier calls left in d_out
        (device cursor, no host round trip); see b2vf_filter_batch_append_device."""
        off = np.ascontiguousarray(h_offsets, np.uint32)
        capi.check(capi.lib().b2vf_filter_batch_append_device(self._h, C.c_void_p(d_in), n_total,
                                                              off.ctypes.data_as(C.POINTER(C.c_uint32)), len(off) - 1,
                                                              C.c_void_p(d_out), out_capacity, C.c_void_p(d_out_offsets),
                                                              first_index, C.c_void_p(d_cursor)))


class InitialYawSearch:
    """The matching node's position-only initialisation (matching.cpp:327-342 SetInitPose with
    init_type OnlyPosition, :344-394 generateGauss2DMapCells, :267-308 getInitialYawAngle) on device clouds."""

    def __init__(self, grid_resolution=0.8, device=0):
        self._h = C.c_void_p()
        capi.check(capi.lib().b2hmap_create(int(device), float(grid_resolution), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and capi is not None and getattr(capi, "_LIB", None) is not None:
            try:
                capi._LIB.b2hmap_destroy(h)
            except Exception:
                pass
            self._h = None

    def GenerateGauss2DMapCells(self, local_map, origin):
        """local_map: DeviceCloud (the BoxFilter crop around `origin`, matching.cpp:166-183)."""
        o = np.asarray(origin, np.float32)
        capi.check(capi.lib().b2hmap_build(self._h, local_map._h, capi._fp(o)))
        return self.Info()

    def Info(self):
        w, h = C.c_int32(), C.c_int32()
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        capi.check(capi.lib().b2hmap_info(self._h, C.byref(w), C.byref(h), capi._fp(mn), capi._fp(mx)))
        return dict(width=w.value, height=h.value, min_xyz=mn, max_xyz=mx)

    def Cells(self):
        info = self.Info()
        n = info["width"] * info["height"]
        mu, sg, cnt = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        capi.check(capi.lib().b2hmap_cells(self._h, capi._fp(mu), capi._fp(sg), cnt.ctypes.data_as(C.POINTER(C.c_int32))))
        shp = (info["width"], info["height"])
        return mu.reshape(shp), sg.reshape(shp), cnt.reshape(shp)

    def GetInitialYawAngle(self, scan, angle_size=270):
        """scan: DeviceCloud in the sensor frame -> (best yaw in radians, per-bin scores)."""
        probs = np.zeros(angle_size, np.float64)
        best = C.c_double()
        capi.check(capi.lib().b2hmap_yaw_search(self._h, scan._h, int(angle_size), capi._dp(probs), C.byref(best)))
        return best.value, probs

    def PoseSearch(self, scan, offsets_xy, angle_size=270):
        """(x, y, yaw) hypothesis scores: offsets_xy (n,2) sensor positions relative to the grid origin -> (n, angle_size)."""
        off = np.ascontiguousarray(offsets_xy, np.float32).reshape(-1, 2)
        probs = np.zeros((len(off), angle_size), np.float64)
        capi.check(capi.lib().b2hmap_pose_search(self._h, scan._h, int(angle_size), capi._fp(off), len(off), capi._dp(probs)))
        return probs


class DistortionAdjust:
    """DistortionAdjust (lidar_localization/src/models/scan_adjust/distortion_adjust.cpp): SetMotionInfo(scan_period,
    velocity) then AdjustCloud(cloud).  velocity = (linear xyz, angular xyz) in the sensor frame.  Unlike the reference,
    whose AdjustCloud rotates its stored velocities in place (so SetMotionInfo has to precede every call, as
    data_pretreat_flow does), the stored motion is left untouched."""

    def __init__(self, device=0):
        self.device = int(device)
        self.scan_period_ = 0.1
        self.velocity_ = np.zeros(3, np.float64)
        self.angular_rate_ = np.zeros(3, np.float64)

    def SetMotionInfo(self, scan_period, linear_velocity, angular_velocity):
        self.scan_period_ = float(np.float32(scan_period))
        self.velocity_ = np.ascontiguousarray(linear_velocity, np.float64)
        self.angular_rate_ = np.ascontiguousarray(angular_velocity, np.float64)

    def AdjustCloudDevice(self, src, dst=None):
        if dst is None:
            dst = DeviceCloud(device=self.device)
        capi.check(capi.lib().b2cloud_distortion_adjust(src._h, self.scan_period_, capi._dp(self.velocity_),
                                                        capi._dp(self.angular_rate_), dst._h))
        return dst

    def AdjustCloud(self, input_cloud):
        """-> (True, adjusted cloud) in the layout of the input."""
        a = np.asarray(input_cloud)
        return True, self.AdjustCloudDevice(DeviceCloud(input_cloud, device=self.device)).Download(layout=a.shape[1])
