"""Host-side mirrors of the reference's CALLERS of the registration / filter plug-ins, backend-agnostic: the same loop
drives the B200 engine (host buffers or device-resident clouds) and the CPU oracle, so parity tests and benches
replay exactly the call pattern of the nodes (BASELINE.json configs 2, 3 and 5).

  FrontEnd / FrontEndDevice   lidar_localization/src/mapping/front_end/front_end.cpp:88-341 (Update), 348-424 (UpdateWithNewFrame)
  MatchingLoop                src/matching/matching.cpp:148-183 (InitGlobalMap, ResetLocalMap), 185-265 (Update, re-crop)
  hypothesis_lattice          the (x, y) lattice of config 5, generalising the yaw scan of matching.cpp:267-308

Nothing here computes: the numerics live behind the callables handed in.
"""
import time

import numpy as np

from .registration import DeviceCloud, transform_cloud


class FrontEnd:
    """The front end's use of the two plug-ins: frame filter -> ScanMatch against the local map -> constant-velocity
    prediction -> new key frame every `key_dist` metres (L1), local map = last `local_frames` key frames (unfiltered),
    down-sampled once 10 key frames exist (front_end.cpp:395-410)."""

    def __init__(self, filt, local_filt, set_target, scan_match, key_dist=2.0, local_frames=20, transform=transform_cloud):
        self.filt, self.local_filt, self.set_target, self.scan_match = filt, local_filt, set_target, scan_match
        self.key_dist, self.local_frames, self.transform = key_dist, local_frames, transform
        self.keyframes = []           # (pose, unfiltered cloud)
        self.pose = None; self.last = None; self.predict = None; self.last_key = None
        self.t_match, self.t_target, self.t_assemble = [], [], []

    def _new_keyframe(self, cloud, pose):
        self.keyframes.append((pose.copy(), cloud))
        if len(self.keyframes) > self.local_frames:
            self.keyframes.pop(0)
        t0 = time.perf_counter()
        local = np.concatenate([self.transform(c, T) for T, c in self.keyframes], axis=0)
        self.t_assemble.append(1e3 * (time.perf_counter() - t0))
        t = time.perf_counter()
        if len(self.keyframes) >= 10:
            local = self.local_filt(local)
        self.set_target(local)
        self.t_target.append(1e3 * (time.perf_counter() - t))
        self.last_key = pose.copy()

    def _advance(self, pose):
        step = np.linalg.inv(self.last.astype(np.float64)) @ pose.astype(np.float64)
        self.predict = (pose.astype(np.float64) @ step).astype(np.float32)
        self.last = pose.copy(); self.pose = pose

    def update(self, cloud, init_pose):
        filtered = self.filt(cloud)
        if self.pose is None:
            self.pose = init_pose.astype(np.float32).copy(); self.last = self.pose.copy(); self.predict = self.pose.copy()
            self._new_keyframe(cloud, self.pose)
            return self.pose
        t = time.perf_counter()
        pose = self.scan_match(filtered, self.predict)
        self.t_match.append(1e3 * (time.perf_counter() - t))
        self._advance(pose)
        if np.sum(np.abs(self.last_key[:3, 3] - pose[:3, 3])) > self.key_dist:
            self._new_keyframe(cloud, pose)
        return pose


class FrontEndDevice(FrontEnd):
    """The same front end with every cloud resident in HBM: the raw frame is uploaded once,
    key frames stay on the device, the local map is assembled (Assemble: one launch for the 20 key frames), filtered (FilterCloud, in
    place) and handed to SetInputTargetCloud without a host round trip."""

    def __init__(self, vf, lvf, reg, key_dist=2.0, local_frames=20, device=0):
        super().__init__(None, None, None, None, key_dist, local_frames)
        self.vf, self.lvf, self.reg, self.device = vf, lvf, reg, device
        self.local = DeviceCloud(device=device)
        self.filtered = DeviceCloud(device=device)
        self.frame = DeviceCloud(device=device)      # upload target, reused; a key frame takes it over and a new one is made
        self.t_upload = []

    def _new_keyframe(self, cloud, pose):
        self.keyframes.append((pose.copy(), cloud))
        if len(self.keyframes) > self.local_frames:
            self.keyframes.pop(0)
        t0 = time.perf_counter()
        self.local.Assemble([c for _, c in self.keyframes], [T for T, _ in self.keyframes])
        self.t_assemble.append(1e3 * (time.perf_counter() - t0))
        t = time.perf_counter()
        if len(self.keyframes) >= 10:
            self.lvf.FilterCloud(self.local, self.local)
        self.reg.SetInputTargetCloud(self.local)
        self.t_target.append(1e3 * (time.perf_counter() - t))
        self.last_key = pose.copy()

    def update(self, cloud, init_pose):
        t = time.perf_counter(); d_cloud = self.frame.Upload(cloud); self.t_upload.append(1e3 * (time.perf_counter() - t))
        if self.pose is None:
            self.pose = init_pose.astype(np.float32).copy(); self.last = self.pose.copy(); self.predict = self.pose.copy()
            self._new_keyframe(d_cloud, self.pose)
            self.frame = DeviceCloud(device=self.device)
            return self.pose
        t = time.perf_counter()
        self.vf.FilterCloud(d_cloud, self.filtered)
        pose = self.reg.ScanMatchCloud(self.filtered, self.predict)[2]
        self.t_match.append(1e3 * (time.perf_counter() - t))          # frame filter + ScanMatch, both on the device
        self._advance(pose)
        if np.sum(np.abs(self.last_key[:3, 3] - pose[:3, 3])) > self.key_dist:
            self._new_keyframe(d_cloud, pose)
            self.frame = DeviceCloud(device=self.device)
        return pose


class MatchingLoop:
    """The matching node's use of the plug-ins (matching.cpp): the global map is down-sampled once (InitGlobalMap,
    :148-163), a +-`size` m box around the pose is cropped out of it and becomes the NDT target (ResetLocalMap,
    :166-183); every frame is filtered and matched against it with a constant-velocity prediction (:185-253), and
    when the pose comes within `edge` m of a box face the box is re-centred and the target rebuilt (:255-262).

    crop(origin) -> the local map handed to set_target; both are callables so that host-buffer, device-resident and
    oracle back ends replay the same loop."""

    def __init__(self, filt, crop, set_target, scan_match, size=100.0, edge=50.0):
        self.filt, self.crop, self.set_target, self.scan_match = filt, crop, set_target, scan_match
        self.size, self.edge = size, edge
        self.origin = None
        self.last = None; self.predict = None
        self.recrops = 0
        self.t_match, self.t_reset = [], []

    def reset_local_map(self, origin):
        t = time.perf_counter()
        self.origin = np.asarray(origin, np.float64).copy()
        self.local = self.crop(self.origin)
        self.set_target(self.local)
        self.t_reset.append(1e3 * (time.perf_counter() - t))

    def set_init_pose(self, pose):
        pose = np.asarray(pose, np.float32)
        self.last = pose.copy(); self.predict = pose.copy()
        self.reset_local_map(pose[:3, 3])

    def update(self, cloud):
        src = self.filt(cloud)
        t = time.perf_counter()
        pose = self.scan_match(src, self.predict)
        self.t_match.append(1e3 * (time.perf_counter() - t))
        step = np.linalg.inv(self.last.astype(np.float64)) @ pose.astype(np.float64)
        self.predict = (pose.astype(np.float64) @ step).astype(np.float32)
        self.last = pose.copy()
        # matching.cpp:255-262: a box face (float edges of BoxFilter::GetEdge) within `edge` metres on any axis -> re-centre
        e = box_edges(self.origin, self.size)
        for i in range(3):
            if abs(np.float32(pose[i, 3]) - e[2 * i]) > self.edge and abs(np.float32(pose[i, 3]) - e[2 * i + 1]) > self.edge:
                continue
            self.recrops += 1
            self.reset_local_map(pose[:3, 3])
            break
        return pose


def box_edges(origin, size=100.0):
    """BoxFilter::CalculateEdge with its float arithmetic (box_filter.cpp:63-70): [min_x, max_x, min_y, max_y, min_z, max_z]."""
    o = np.asarray(origin, np.float32)
    s = np.float32(size)
    return [-s + o[0], s + o[0], -s + o[1], s + o[1], -s + o[2], s + o[2]]


def hypothesis_lattice(truth6, pose6_to_matrix, side=32, pitch=2.0, offset=(0.4, -0.3)):
    """config 5: side x side lattice of positions (pitch metres) around the true pose, shifted by `offset` so that no
    hypothesis is the truth itself; every hypothesis keeps the true orientation. -> (side*side, 4, 4) float32"""
    half = (side - 1) / 2.0
    gx, gy = np.meshgrid(np.arange(side) - half, np.arange(side) - half, indexing="ij")
    out = []
    for dx, dy in zip(gx.ravel(), gy.ravel()):
        p = np.array(truth6, np.float64)
        p[0] += pitch * dx + offset[0]; p[1] += pitch * dy + offset[1]
        out.append(pose6_to_matrix(p).astype(np.float32))
    return np.stack(out)
