"""voxplat_b200 -- B200-native (sm_100a) chunk-rebuild path of the Voxplat voxel engine.

The product is the CUDA C-ABI library (include/voxplat_b200.h, voxplat_b200/csrc/*.cu); this package is
the Python plumbing around it: ctypes binding (api), deterministic synthetic worlds (worldgen) and the
multi-GPU slab driver (slab).  Nothing here computes on the CPU what the kernels compute.
"""
from .api import (Context, MultiContext, VoxplatError, load_library, world_file_info, VP_REBUILD_SPLAT, VP_REBUILD_MESH, RESULT_DTYPE)  # noqa: F401
from . import worldgen  # noqa: F401
