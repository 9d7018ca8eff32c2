"""The reference's own known-answer vectors (tests/known_answers.py) against the CPU oracle: part of what pins the checker."""
import known_answers as ka
from oracle_binding import oracle_context, oracle_scene


def test_tangent_frame_maps_y_to_normal():
    ka.tangent_frame_maps_y_to_normal(oracle_context())


def test_reflect_follows_the_source():
    ka.reflect_follows_the_source(oracle_context())


def test_one_pixel_environment_map():
    ka.one_pixel_environment_map(oracle_scene("test_scenes/environment_map_sampling.json", 32, 24))
