"""Alias of ``difffe_physics_lab_b200.mesh`` under the reference's module path ``diffhe.mesh``."""
from difffe_physics_lab_b200.mesh import *  # noqa: F401,F403
from difffe_physics_lab_b200 import mesh as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
