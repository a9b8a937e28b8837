"""driftdiffusionpoissonsystems.jl_b200 -- B200-native hot path of DriftDiffusionPoissonSystems.jl.

Import as ``import ddps_b200`` (the root-level shim maps that name onto this directory, whose
literal name is not a valid Python identifier).

Public surface (mirrors the reference's exports, ``src:6-7``):
    Mesh, read_mesh, construct_mesh4, construct_mesh8, construct_mesh, structured_mesh,
    solve_ddpe, solve_semlin_poisson, aquire_boundary, SinhRHS, PolyRHS, Session
"""
from .mesh import (Mesh, read_mesh, construct_mesh4, construct_mesh8, construct_mesh, structured_mesh, structured_mesh_part,
                   structured_boundary)
from .functors import SinhRHS, PolyRHS, Const, Affine, StepX, StepY
from . import functors
from .api import (Session, DDPSError, solve_ddpe, solve_semlin_poisson, aquire_boundary, dirichlet_nodes,
                  neumann_edges)
from .postprocess import aquire_boundary_separate_edges, get_gradients, calculate_current, current_from_fields
from . import problems
from . import amg

__all__ = ["Mesh", "read_mesh", "construct_mesh4", "construct_mesh8", "construct_mesh", "structured_mesh", "structured_mesh_part", "structured_boundary",
           "solve_ddpe", "solve_semlin_poisson", "aquire_boundary", "SinhRHS", "PolyRHS", "Const", "Affine", "StepX", "StepY", "functors", "Session", "DDPSError",
           "problems", "amg", "dirichlet_nodes", "neumann_edges", "aquire_boundary_separate_edges", "get_gradients",
           "calculate_current", "current_from_fields"]
