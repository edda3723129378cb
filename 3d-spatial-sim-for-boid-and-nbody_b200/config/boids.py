"""Boids parameters: same keys and default values as the reference's ``config/boids.py:30-46``."""

BOIDS = {
    "count": 500000,
    "bounds": 500.0,
    "max_speed": 25.0,
    "max_force": 60.0,
    "size": 1.2,
    "wall_margin": 3.0,
    "wall_weight": 10.0,
    "perception_radius": 5.0,
    "separation_radius": 3.0,
    "separation_weight": 2.5,
    "alignment_weight": 1.0,
    "cohesion_weight": 1.0,
    "color_blend_rate": 1.0,
}
