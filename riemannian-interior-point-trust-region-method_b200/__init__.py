"""riptrm_b200 -- B200-native (sm_100a CUDA) trust-region path of the Riemannian interior-point
trust-region method, behind the reference's `RIPTRM(option).run(problem) -> Output` interface.

The directory name carries the reference's name; import it as `riptrm_b200` (repo-root shim)."""
from . import _lib, datagen, io, options, structure
from ._lib import RiptrmError, load_library
from .solver import RIPTRM, BatchSolver, ColumnsSolver, StiefelSolver, Output, columns_bench, trace_to_log
from .structure import (NonnegPCAStructure, NonnegPCAStiefelStructure, RosenbrockStructure, StableIdStructure,
                        structure_from_problem)

__all__ = ["RIPTRM", "BatchSolver", "ColumnsSolver", "StiefelSolver", "columns_bench", "Output", "trace_to_log", "RiptrmError", "load_library",
           "NonnegPCAStructure", "NonnegPCAStiefelStructure", "RosenbrockStructure", "StableIdStructure", "structure_from_problem"]


def __getattr__(name):
    # `sharding` needs torch.distributed; the library, the drop-in module and the solver classes do not need torch at all
    if name == "sharding":
        import importlib
        return importlib.import_module(".sharding", __name__)
    raise AttributeError(name)
