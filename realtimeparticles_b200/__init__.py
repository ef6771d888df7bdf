"""realtimeparticles_b200 -- B200-native (sm_100a CUDA) backend for the RealTimeParticles hot path:
uniform-grid neighbour search (cell id, onesweep radix sort, cell start/end table) and the neighbour-loop
kernels it feeds (Reynolds boids, Position Based Fluids, the clouds thermodynamics extension).

The compute path is realtimeparticles_b200/lib/librtp_cuda.so behind the C ABI of include/rtp_cuda.h;
`models` mirrors the reference's Physics::Model interface on top of it. There is no CPU fallback.
"""
from . import _abi  # noqa: F401
from .models import (Boids, Boundary, Clouds, CreateModel, Dimension, Fluids, Model, ModelParams, ModelType,  # noqa: F401
                     PhysicsCase)

__all__ = ["Boids", "Fluids", "Clouds", "CreateModel", "Model", "ModelParams", "ModelType", "Boundary", "Dimension",
           "PhysicsCase", "_abi"]
