"""wave_tracer_b200 -- B200-native drop-in for wave_tracer's per-sample integrator hot path.

Python is the thin host layer (scene description, torch/NCCL plumbing); all compute lives in the native library
wave_tracer_b200/libwt_b200.so (hand-written sm_100a CUDA behind the C-ABI of include/wtgpu.h).
"""
from . import _abi
from .scene import *  # noqa: F401,F403
from .render import render, develop, GpuScene  # noqa: F401
