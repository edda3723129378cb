"""Mirror of the reference's ``nbody`` backend boundary (nbody/gpu_backend.py)."""
from .gpu_backend import (Backend, B200BarnesHutSimulation, CUDASimulation, create_gpu_simulation,  # noqa: F401
                          detect_backend, force_backend, get_backend)
