"""obj2voxel_b200 — B200-native surface voxelizer behind the obj2voxel C API.

The package holds only what the voxelization hot path needs:
  csrc/     hand-written sm_100a CUDA kernels + the C-ABI shared library (libobj2voxel_b200.so)
  _lib.py   ctypes declarations of include/obj2voxel.h and include/obj2voxel_b200.h
  api.py    host-side mirror of the reference interface (Instance/Texture) and the bulk/device Engine
  slabs.py  Z-slab partitioning of the sample grid across the GPUs of one box (torch.distributed plumbing)
  meshes.py synthetic meshes of BASELINE.json's configs
There is no CPU implementation in this package: without the built library or without a GPU, calls fail loudly.
"""
from . import _lib
from ._lib import (BLEND_STRATEGY, ERR_DEVICE, ERR_DOUBLE_VOXELIZATION, ERR_IO_WRITE, ERR_NO_INPUT, ERR_NO_OUTPUT,
                   ERR_NO_RESOLUTION, ERR_OK, MAX_STRATEGY, UV_CLAMP, UV_WRAP, build, load)
from .api import DeviceError, Engine, Instance, Texture, make_params, sort_voxels

__all__ = ["_lib", "build", "load", "Engine", "Instance", "Texture", "DeviceError", "make_params", "sort_voxels",
           "MAX_STRATEGY", "BLEND_STRATEGY", "UV_CLAMP", "UV_WRAP", "ERR_OK", "ERR_NO_INPUT", "ERR_NO_OUTPUT",
           "ERR_NO_RESOLUTION", "ERR_IO_WRITE", "ERR_DOUBLE_VOXELIZATION", "ERR_DEVICE"]
