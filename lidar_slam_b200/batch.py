"""Sharding of independent scan matches over the GPUs of one box (BASELINE.json configs 4/5).

One process per GPU (torchrun); the target map is replicated, the batch of (source, guess) pairs is
split into contiguous blocks, and there is NO collective on the match path (north_star): the only
communication is the final gather of (pose, score, iterations) per match, done here with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of n_items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(n_items, world):
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_rows(local_rows, n_items, dist=None, device=None):
    """All ranks contribute their (n_local, k) float64 rows; rank 0 gets the (n_items, k) table in batch
    order (other ranks get None).  Uses all_gather on padded blocks so it works with NCCL and gloo."""
    import torch
    local_rows = np.ascontiguousarray(local_rows, np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_rows
    world, rank = dist.get_world_size(), dist.get_rank()
    k = local_rows.shape[1]
    sizes = shard_sizes(n_items, world)
    pad = max(sizes)
    buf = torch.zeros((pad, k), dtype=torch.float64, device=device)
    if local_rows.shape[0]:
        buf[:local_rows.shape[0]] = torch.from_numpy(local_rows).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    if rank != 0:
        return None
    return np.concatenate([out[r][:sizes[r]].cpu().numpy() for r in range(world)], axis=0)


def best_hypothesis(rows, score_col=0):
    """config 5: index of the best-scoring hypothesis (highest NDT score)."""
    return int(np.argmax(rows[:, score_col]))


def pose6_to_matrix(p):
    """(x, y, z, roll, pitch, yaw) -> 4x4, R = Rx(roll) Ry(pitch) Rz(yaw): the convention of the reference's 6-vector
    (NormalDistributionsTransform.cpp:370-373: Translation * AngleAxis(X) * AngleAxis(Y) * AngleAxis(Z))."""
    cx, sx, cy, sy, cz, sz = np.cos(p[3]), np.sin(p[3]), np.cos(p[4]), np.sin(p[4]), np.cos(p[5]), np.sin(p[5])
    T = np.eye(4)
    T[:3, :3] = [[cy * cz, -cy * sz, sy],
                 [cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy],
                 [sx * sz - cx * sy * cz, sx * cz + cx * sy * sz, cx * cy]]
    T[:3, 3] = p[:3]
    return T


def scan_match_batch_multi_device(registrations, sources, predict_poses):
    """One process, several devices: `registrations` holds one NDTRegistration per device (each with the
    same target set); the batch is split into contiguous blocks, every block runs on its own host thread / handle /
    stream, and the results come back in batch order.  No communication between the devices.
    -> (poses (B,4,4) float32, results structured array)"""
    import threading
    B = len(predict_poses)
    shared = not isinstance(sources, (list, tuple))
    world = len(registrations)
    out = [None] * world
    err = [None] * world

    def work(r):
        lo, hi = shard_range(B, r, world)
        if hi == lo:
            return
        try:
            src = sources if shared else list(sources[lo:hi])
            out[r] = registrations[r].ScanMatchBatch(src, list(predict_poses[lo:hi]))
        except Exception as e:      # surfaced on the caller's thread
            err[r] = e

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in err:
        if e is not None:
            raise e
    parts = [o for o in out if o is not None]
    return np.concatenate([p[0] for p in parts], axis=0), np.concatenate([p[1] for p in parts], axis=0)


def relocalize(registration, yaw_search, scan_filtered, scan_device, position, lattice=9, pitch=1.0, top=16, angle_size=270):
    """Global relocalisation from a position prior (BASELINE.json config 5, generalising the reference's yaw-only scan,
    matching.cpp:199-232,267-308, to (x, y, yaw)): the Gaussian height grid (`yaw_search`: an InitialYawSearch built
    around `position`; `scan_device`: the scan as a DeviceCloud) scores lattice x lattice sensor positions (pitch
    metres) x angle_size yaw bins in one launch, the `top` best poses become the initial guesses of ONE batched NDT
    launch (`registration.ScanMatchBatch` with a shared source), and the best final NDT score wins.
    -> (best pose 4x4, index into the candidates, candidate poses after NDT (top,4,4), results)"""
    half = (lattice - 1) / 2.0
    offs = np.array([[(ix - half) * pitch, (iy - half) * pitch] for iy in range(lattice) for ix in range(lattice)], np.float32)
    probs = yaw_search.PoseSearch(scan_device, offs, angle_size)
    flat = np.where(np.isfinite(probs), probs, -np.inf).ravel()
    cand = np.argsort(-flat)[:max(1, int(top))]
    delta = np.float32(2 * np.pi / angle_size)
    guesses = []
    for c in cand:
        o, b = divmod(int(c), angle_size)
        pose6 = np.array([position[0] + offs[o, 0], position[1] + offs[o, 1], position[2], 0.0, 0.0, float(np.float32(b) * delta)])
        guesses.append(pose6_to_matrix(pose6).astype(np.float32))
    poses, res = registration.ScanMatchBatch(scan_filtered, guesses)
    score = np.where(res["converged"] > 0, res["score"], -np.inf)
    k = int(np.argmax(score))
    return poses[k], k, poses, res
