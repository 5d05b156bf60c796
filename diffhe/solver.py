"""Alias of ``difffe_physics_lab_b200.solver`` under the reference's module path ``diffhe.solver``."""
from difffe_physics_lab_b200.solver import *  # noqa: F401,F403
from difffe_physics_lab_b200 import solver as _m

globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
