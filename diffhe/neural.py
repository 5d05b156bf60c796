"""Alias of ``difffe_physics_lab_b200.neural`` under the reference's module path ``diffhe.neural``."""
from difffe_physics_lab_b200.neural import *  # noqa: F401,F403
from difffe_physics_lab_b200 import neural as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
