"""sparspak.jl_b200 — B200-native numeric factor/solve engine behind Sparspak.jl's API.

Host mirror of the reference interface (Python, no bangs) over a C-ABI CUDA library
(csrc/, include/spk_b200.h).  Only the numeric hot path runs on the GPU; ordering,
etree and symbolic factorisation are host code, as in the reference."""
from .problem import (Problem, Graph, Ordering, ETree, Grid, inaij, inbi, insparse, outsparse, infullrhs,
                      computeresidual, makerhs, makegridproblem, makestructuresymmetric,
                      isstructuresymmetric, mmd, nd_grid_order)
from .sparse_base import _SparseBase, _SparseSpdBase
from .sparse_solver import (SparseSolver, SparseSpdSolver, SequenceError, findorder, findorderperm,
                            symbolicfactor, inmatrix, factor, triangularsolve, solve)
from .csc_interface import sparspaklu, sparspaklu_, ldiv, backslash
from . import matrices

__version__ = "0.1.0"
