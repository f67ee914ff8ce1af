"""The product's Newton / More-Thuente controller (csrc/b2_ndt_math.cuh, the code thread 0 of every
match CTA runs) compiled for the host and driven by oracle derivative passes must reproduce the
oracle's align: same iterations, same number of passes and line-search trials, same pose."""
import numpy as np

from lidar_slam_b200 import synth
from tests.conftest import f32
from tests.hostcheck_util import controller_align


def test_controller_reproduces_oracle_align(oracle, small_map, scans):
    grid = oracle.Grid(small_map, 1.0)
    rng = np.random.default_rng(21)
    for compat in (1, 0):
        prm = oracle.params(step_size=f32(0.1), trans_eps=f32(0.01), pcl17_compat=compat)
        for truth, scan in scans[:2]:
            src, _, _, _ = oracle.voxel_filter(scan, 1.3, 1.3, 1.3)
            for trial in range(3):
                guess = synth.pose6_to_matrix(synth.perturb_pose(truth, rng)).astype(np.float32)
                if trial == 2:
                    guess = np.eye(4, dtype=np.float32)      # the `guess != Identity` branch (NDTM:323)
                ref = oracle.align(grid, prm, src, guess)
                for force_svd in (0, 1):
                    got = controller_align(grid, prm, src, guess, force_svd=force_svd)
                    assert got["iterations"] == ref["iterations"] and got["converged"] == ref["converged"]
                    assert got["passes"] == ref["passes"] and got["mt_trials"] == ref["mt_trials"]
                    assert np.max(np.abs(got["p"] - ref["p"])) < 1e-9
                    assert np.max(np.abs(got["pose"] - ref["pose"])) < 1e-6
                    assert abs(got["score"] - ref["score"]) <= 1e-9 * max(1.0, abs(ref["score"]))


def test_controller_degenerate(oracle, small_map):
    grid = oracle.Grid(small_map, 1.0)
    prm = oracle.params()
    far = np.array([[1e4, 1e4, 50, 0], [1e4 + 1, 1e4, 50, 0]], np.float32)
    got = controller_align(grid, prm, far, np.eye(4, dtype=np.float32))
    assert got["converged"] and got["iterations"] == 0 and got["passes"] == 1
