"""Alias of ``difffe_physics_lab_b200.loss`` under the reference's module path ``diffhe.loss``."""
from difffe_physics_lab_b200.loss import *  # noqa: F401,F403
from difffe_physics_lab_b200 import loss as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
