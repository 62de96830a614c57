"""Space data types -- the B200 mirror of ``foragax/base/space_classes.py:8-55``.

Same names and fields (``Point, Wall, Ray, Space, NetworkSpace``); leaves are float32 CUDA
tensors.  A set of W walls is the struct-of-arrays ``Wall(p1=Point(x[W], y[W]), p2=...)``
the reference builds with ``jax.vmap(Wall.create_wall)`` (16 bytes per wall).
"""
from .. import struct


@struct.dataclass
class Point:
    x: object
    y: object

    def create_point(x, y):
        return Point(x, y)


@struct.dataclass
class Wall:
    p1: Point
    p2: Point

    def create_wall(p1, p2):
        return Wall(p1, p2)


@struct.dataclass
class Ray:
    ray_origin: Point
    ray_direction: Point  # cosines and sines of the angle
    ray_length: object

    def create_ray(ray_origin, ray_direction, ray_length):
        return Ray(ray_origin, ray_direction, ray_length)


@struct.dataclass
class Space:
    x_min: object
    x_max: object
    y_min: object
    y_max: object
    torous: bool
    walls: Wall

    def create_space(x_min, x_max, y_min, y_max, torous, wall_begins, wall_ends):
        return Space(x_min, x_max, y_min, y_max, torous, Wall(wall_begins, wall_ends))


@struct.dataclass
class NetworkSpace:
    pass
