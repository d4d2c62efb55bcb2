"""djets-b200: B200-native backend for the hot path of ChevronETC/DistributedJets.jl.

The directory name carries a dot, so the package is loaded through ``djets_loader.load()`` (repo
root) under the module name ``djets_b200``; see INTEGRATION.md.
"""
from ._lib import (DJetsError, DimensionMismatch, NotLocal, LIB_PATH, F32, F64, BLOCK_ZERO, BLOCK_DENSE,
                   BLOCK_DIAG)
from .partition import (urange, even_cuts, even_ranges, default_grid, cuts_from_ranges, ranges_from_cuts,
                        grid_of, plan_exchange)
from .core import (Context, init, set_default_context, context, workers, nworkers, nccl_unique_id, torch_allgather,
                   PinnedBuffer, Graph, Scalar, Expr, lazy, bc_abs, bc_sqrt, bc_exp, bc_log, bc_max, bc_min,
                   JetDSpace, DBArray, LinExpr, dotted, zeros, ones, rand, Array, getblock,
                   getblock_into, setblock, collect, dot, norm, blockmap, localblockindices, procs, nprocs,
                   nblocks, JopDense, JopDiagonal, JopZeroBlock, iszero, JopDBlock, JopAdjoint, blockop,
                   mul)

# the reference's exports (src/DistributedJets.jl:863)
__all__ = ["DBArray", "JetDSpace", "blockmap", "localblockindices"]
