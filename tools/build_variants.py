"""Build kernel-shape variants of libb2ndt.so for a sweep on the GPU box (selected with B2NDT_LIB).
usage: python tools/build_variants.py name:-DA=1,-DB=2 [name:flags ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_slam_b200 import build as b

for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    out = b.build_cuda(variant=name, variant_flags=[f for f in flags.split(",") if f])
    log = os.path.join(b.LIBDIR, "nvcc_ptxas_%s.log" % name)
    lines = open(log).read().splitlines()
    for i, ln in enumerate(lines):
        if "ndt_match_kernel" in ln and "Function properties" in ln:
            print(name, "|", lines[i + 1].strip(), "|", lines[i + 2].strip())
    print("built", out)
