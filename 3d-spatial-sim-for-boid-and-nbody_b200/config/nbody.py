"""N-body physics parameters: same keys and default values as the reference's
``config/nbody.py:57-73`` (NBODY dict); window/camera/grid settings are viewer-only and are
not mirrored.  There is no ``dt`` key in the reference: dt comes from the caller
(tools/record.py:749, nbody/simulation.py:802)."""

BODY_COUNT = 150_000   # config/nbody.py:16
THETA = 0.8            # config/nbody.py:17

NBODY = {
    "count": BODY_COUNT,
    "spawn_radius": 500.0,
    "G": 0.1,
    "theta": THETA,
    "softening": 2.0,
    "damping": 1.0,
    "distribution": "galaxy",
    "point_size": 1.5,
    "max_speed_color": 15.0,
}
