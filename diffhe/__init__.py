"""``diffhe`` — the reference's package name, aliased to the B200 implementation.

``from diffhe.mesh import FEMesh`` / ``from diffhe.solver import DifferentiableFESolver`` (the import
paths the reference's tests use, tests/test_fem.py:22-23 upstream) resolve to
``difffe_physics_lab_b200``.
"""
from difffe_physics_lab_b200 import DifferentiableFESolver, FEMesh, NeuralPDE, PhysicsLoss, __version__

__all__ = ["FEMesh", "DifferentiableFESolver", "PhysicsLoss", "NeuralPDE"]
