"""Parameter mirrors of the reference's config package."""
from . import boids, nbody  # noqa: F401
