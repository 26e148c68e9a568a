"""B200-native ERT forward + Jacobian path (drop-in for pyGIMLi's DCSRMultiElectrodeModelling
behind ERTModelling.response / createJacobian).  See DESIGN.md."""
from .mesh import MeshArrays, grid_mesh_2d, grid_mesh_3d, graded_axis, create_p2, create_h2, mark_electrode_nodes
from .scheme import SchemeArrays, create_dd, create_slm, create_dd_complete, create_grid_dd, geometric_factors
from .ert_modelling import (ERTModellingB200, CoreB200, JacobianB200, MultLeftRightMatrixB200, coverageDCtrans,
                            createCoverage, managerCoverage)

__all__ = ["MeshArrays", "grid_mesh_2d", "grid_mesh_3d", "graded_axis", "create_p2", "create_h2", "mark_electrode_nodes",
           "SchemeArrays", "create_dd", "create_slm", "create_dd_complete", "create_grid_dd", "geometric_factors",
           "ERTModellingB200", "CoreB200", "JacobianB200", "MultLeftRightMatrixB200", "coverageDCtrans", "createCoverage", "managerCoverage"]
