import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lidar_slam_b200 import synth, capi
from lidar_slam_b200.registration import NDTRegistration, VoxelFilter
scene = synth.Scene(leg=500.0)
target = scene.make_map(1_000_000, 2.0)
B = 2000
vf = VoxelFilter(1.3, 1.3, 1.3)
s = 5 + (scene.path_length - 10) * (np.arange(B) + 0.5) / B
truth = np.stack([scene.path_pose(v) for v in s])
srcs = []
for c0 in range(0, B, 250):
    raws = scene.scans(np.arange(c0, min(B, c0 + 250)), truth[c0:c0 + 250])
    srcs += [vf.Filter(r)[1] for r in raws]
rng = np.random.default_rng(1)
guesses = np.stack([synth.pose6_to_matrix(synth.perturb_pose(truth[k], rng)).astype(np.float32) for k in range(B)])
cat = np.ascontiguousarray(np.concatenate(srcs)); off = np.zeros(B + 1, np.uint32); off[1:] = np.cumsum([len(x) for x in srcs])
g = np.ascontiguousarray(guesses.transpose(0, 2, 1).reshape(B, 16))
reg = NDTRegistration(1.0, 0.1, 0.01, 30); reg.SetInputTarget(target)
L = capi.lib(); out = np.zeros((B, 16), np.float32); res = np.zeros(B, capi.RESULT_DTYPE)
def run(buf):
    capi.check(L.b2ndt_align_batch(reg._h, buf.ctypes.data, len(buf), 16, 12, off.ctypes.data_as(C.POINTER(C.c_uint32)), B,
                                   g.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)), res.ctypes.data))
def t(f, n=5):
    f(); f()
    ts = []
    for _ in range(n):
        a = time.perf_counter(); f(); ts.append(1e3 * (time.perf_counter() - a))
    return np.median(ts)
pin = torch.from_numpy(cat).pin_memory(); catp = pin.numpy()
dev = torch.empty_like(pin, device="cuda")
def h2d():
    dev.copy_(pin, non_blocking=True); torch.cuda.synchronize()
print("bytes", cat.nbytes, "H2D pinned ms", t(h2d))
for ch in (1, 2, 4, 8, 16):
    os.environ["B2NDT_CHUNKS"] = str(ch)
    print("chunks", ch, "pageable ms", t(lambda: run(cat)), "pinned ms", t(lambda: run(catp)))
