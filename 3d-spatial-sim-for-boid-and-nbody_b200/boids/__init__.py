"""Mirror of the reference's ``boids`` update seam (boids/flock.py:627-678)."""
from .flock import B200Flock, attach, make_params  # noqa: F401
