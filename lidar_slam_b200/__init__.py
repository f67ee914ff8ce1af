"""lidar_slam_b200: B200-native NDT scan matching + voxel filtering behind Lidar-SLAM's plug-in API.

Only the registration hot path lives here: CUDA kernels + C ABI in csrc/, the
host-side mirror of the reference's RegistrationInterface / CloudFilterInterface in registration.py
(Python) and include/lidar_localization/ (C++).
"""
__all__ = ["build", "capi", "registration", "synth"]
