"""instancefusion_b200 -- B200-native (sm_100a) dense frame-to-model tracker.

A from-scratch replacement for the CUDA hot path of InstanceFusion's ElasticFusion core
(RGBDOdometry::getIncrementalTransformation and its pyramid builders).  The product is the C-ABI
shared library ``libef_track.so`` (include/ef_track.h) built from csrc/; this package is the
Python-side mirror of the reference's operator interface used by the tests and bench.py.
There is no CPU fallback: importing works anywhere, but every compute call needs the CUDA
library and a GPU and raises otherwise.
"""
from .binding import EFError, lib, lib_path, build  # noqa: F401
from .rgbd_odometry import RGBDOdometry  # noqa: F401

__all__ = ["RGBDOdometry", "EFError", "lib", "lib_path", "build"]
