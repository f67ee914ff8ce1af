"""The drop-in boundary: the C-ABI library loads, exports every symbol include/b2ndt.h declares, and
fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from lidar_slam_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "b2ndt.h")).read()
    declared = set(re.findall(r"\b(b2(?:ndt|vf|cloud|hmap|_pcd)?_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("b2_status")
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    path = build.build_cuda()
    nm = subprocess.run(["nm", "-D", "--defined-only", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = set(re.findall(r" T (b2[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported
    lib = capi.lib()
    for s in declared:
        assert hasattr(lib, s)


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", capi.lib_path()], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_python_structs_match_header_layout():
    assert C.sizeof(capi.Result) == capi.RESULT_DTYPE.itemsize == 88
    assert C.sizeof(capi.Params) == 56
    assert C.sizeof(capi.TargetInfo) == 52


def test_no_cpu_fallback_without_gpu():
    lib = capi.lib()
    if lib.b2_device_count() > 0:
        pytest.skip("GPU present")
    from lidar_slam_b200.registration import NDTRegistration, VoxelFilter
    with pytest.raises(capi.B2Error) as e:
        NDTRegistration(1.0, 0.1, 0.01, 30)
    assert e.value.code == capi.B2_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.B2Error):
        VoxelFilter(1.3, 1.3, 1.3)
    from lidar_slam_b200.registration import DeviceCloud
    with pytest.raises(capi.B2Error):
        DeviceCloud()


def test_invalid_arguments_are_rejected_before_any_cuda_call():
    lib = capi.lib()
    h = C.c_void_p()
    assert lib.b2vf_create(0.0, 1.0, 1.0, 0, C.byref(h)) == capi.B2_ERR_INVALID
    assert b"leaf" in lib.b2_last_error()
    p = capi.Params()
    lib.b2ndt_params_default(C.byref(p))
    assert (p.res, p.step_size, p.trans_eps, p.outlier_ratio, p.max_iter, p.min_pts, p.eig_mult, p.pcl17_compat) == \
        (1.0, 0.1, 0.01, 0.55, 30, 6, 0.01, 1)
    p.res = -1.0
    assert lib.b2ndt_create(C.byref(p), 0, C.byref(h)) == capi.B2_ERR_INVALID


def test_cpp_dropin_headers_compile():
    """the C++ mirror of RegistrationInterface / CloudFilterInterface builds against the PCL-free shim"""
    assert build.build_host() is not None
    src = os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp")
    exe = os.path.join(build.LIBDIR, "test_dropin")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                           "-L", build.LIBDIR, "-lb2host", "-lb2ndt", "-Wl,-rpath," + build.LIBDIR])
    assert os.path.exists(exe)
