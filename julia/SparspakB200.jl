# SparspakB200.jl — the reference-side binding of libspkb200.so (include/spk_b200.h).
#
#   using Sparspak, SparspakB200        # after this, Float64 / Float32 problems factor and solve on the GPU
#
# Sparspak.jl dispatches its dense kernels on the element type (SpkSpdMMOps.jl:186-351: generic Julia for BigFloat /
# Dual / MultiFloat, BLAS ccalls for Float64 / Float32).  This file does the same one level up: it adds the methods
#
#   _factor!(s::_SparseBase{Int64,FT})              SpkSparseBase.jl:378-391       FT in (Float64, Float32)
#   _triangularsolve!(s::_SparseBase{Int64,FT}, x)  SpkSparseBase.jl:400-416
#   _factor!(s::_SparseSpdBase{Int64,FT})           SpkSparseSpdBase.jl:313-332
#   _triangularsolve!(s::_SparseSpdBase{Int64,FT}, x) SpkSparseSpdBase.jl:334-356
#
# which are more specific than the reference's `where {IT, FT}` methods, so every other element type keeps falling
# through to the reference.  Each solver object gets a PLAN (structure + factors resident in HBM) kept in a
# WeakKeyDict; the plan is rebuilt whenever `_symbolicfactor!` re-allocated the factor storage (detected through the
# identity of `s.lnz`, which `_symbolicfactor!` replaces, SpkSparseBase.jl:242-243).  `s.lnz / s.unz / s.ipiv` are
# overwritten in place after `factor!`, exactly as the reference leaves them (tests inspect them).
#
# NOT RUN in the build image of this repository (it has no Julia): the same boundary is exercised there through
# ctypes (sparspak.jl_b200/_cudalib.py).  The argument lists below are the ones include/spk_b200.h declares.
module SparspakB200

using Sparspak
using Sparspak.SpkSparseBase: _SparseBase
using Sparspak.SpkSparseSpdBase: _SparseSpdBase
import Sparspak.SpkSparseBase
import Sparspak.SpkSparseSpdBase

const libspk = get(ENV, "SPARSPAK_B200_LIB", "libspkb200.so")

lasterror() = unsafe_string(ccall((:spk_last_error, libspk), Cstring, ()))

mutable struct Plan
    h::Ptr{Cvoid}
    lnz_id::UInt            # objectid of the lnz array the plan was built for
    permset::Bool
end

const plans = WeakKeyDict{Any, Plan}()

function destroy!(p::Plan)
    if p.h != C_NULL
        ccall((:spk_plan_destroy, libspk), Cvoid, (Ptr{Cvoid},), p.h)
        p.h = C_NULL
    end
    return nothing
end

xunz_ptr(s::_SparseBase) = pointer(s.xunz)
xunz_ptr(s::_SparseSpdBase) = Ptr{Int64}(C_NULL)

# one plan per solver object and symbolic factorisation; device = SPARSPAK_B200_DEVICE (default 0)
function plan!(s)
    p = get(plans, s, nothing)
    if p === nothing || p.h == C_NULL || p.lnz_id != objectid(s.lnz)
        p === nothing || destroy!(p)
        dev = parse(Int32, get(ENV, "SPARSPAK_B200_DEVICE", "0"))
        h = GC.@preserve s ccall((:spk_plan_create, libspk), Ptr{Cvoid},
            (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int32, Int32, Int32),
            s.n, s.nsuper, s.xsuper, s.snode, s.xlindx, s.lindx, s.xlnz, xunz_ptr(s), dev, 0, 1)
        h == C_NULL && error("spk_plan_create: " * lasterror())
        p = Plan(h, objectid(s.lnz), false)
        finalizer(destroy!, p)
        plans[s] = p
    end
    return p
end

check(rc, what) = rc <= -100 ? error(what * ": " * lasterror()) : rc

# ---- Float64 -------------------------------------------------------------------------------------------------
function SpkSparseBase._factor!(s::_SparseBase{Int64, Float64})
    s.n == 0 && error("An empty problem. No matrix.")
    p = plan!(s)
    check(ccall((:spk_plan_set_values, libspk), Int64, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), p.h, s.lnz, s.unz), "spk_plan_set_values")
    s.errflag = check(ccall((:spk_plan_factor, libspk), Int64, (Ptr{Cvoid},), p.h), "spk_plan_factor")
    check(ccall((:spk_plan_get_factors, libspk), Int64, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), p.h, s.lnz, s.unz, s.ipiv), "spk_plan_get_factors")
    s.errflag != 0 && error("An empty problem. No matrix.")          # sic, SpkSparseBase.jl:387
    return true
end

function SpkSparseSpdBase._factor!(s::_SparseSpdBase{Int64, Float64})
    s.n == 0 && error("An empty problem. No matrix.")
    p = plan!(s)
    check(ccall((:spk_plan_set_values, libspk), Int64, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), p.h, s.lnz, C_NULL), "spk_plan_set_values")
    s.errflag = check(ccall((:spk_plan_factor, libspk), Int64, (Ptr{Cvoid},), p.h), "spk_plan_factor")
    check(ccall((:spk_plan_get_factors, libspk), Int64, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), p.h, s.lnz, C_NULL, C_NULL), "spk_plan_get_factors")
    s.errflag != 0 && error("An empty problem. No matrix.")
    return true
end

function trisolve64!(s, solution::AbstractVector{Float64})
    s.n == 0 && error("An empty problem. No solution.")
    p = plan!(s)
    if !p.permset
        check(ccall((:spk_plan_set_perm, libspk), Int64, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), p.h, s.order.rperm, s.order.rinvp), "spk_plan_set_perm")
        p.permset = true
    end
    x = solution isa Vector{Float64} ? solution : collect(solution)
    check(ccall((:spk_plan_triangularsolve, libspk), Int64, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), p.h, x, 1, s.n), "spk_plan_triangularsolve")
    x === solution || (solution .= x)
    return true
end
SpkSparseBase._triangularsolve!(s::_SparseBase{Int64, Float64}, solution::AbstractVector{Float64}) = trisolve64!(s, solution)
SpkSparseSpdBase._triangularsolve!(s::_SparseSpdBase{Int64, Float64}, solution::AbstractVector{Float64}) = trisolve64!(s, solution)

# ---- Float32: values cross the boundary as Float32, the arithmetic (and the pivot sequence) is the FP64 engine's --
function SpkSparseBase._factor!(s::_SparseBase{Int64, Float32})
    s.n == 0 && error("An empty problem. No matrix.")
    s.errflag = check(ccall((:spk_lufactor_f32, libspk), Int64,
        (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float32}, Ptr{Int64}, Ptr{Float32}, Ptr{Int64}),
        s.n, s.nsuper, s.xsuper, s.snode, s.xlindx, s.lindx, s.xlnz, s.lnz, s.xunz, s.unz, s.ipiv), "spk_lufactor_f32")
    s.errflag != 0 && error("An empty problem. No matrix.")
    return true
end
function SpkSparseSpdBase._factor!(s::_SparseSpdBase{Int64, Float32})
    s.n == 0 && error("An empty problem. No matrix.")
    s.errflag = check(ccall((:spk_ldltfactor_f32, libspk), Int64,
        (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float32}),
        s.n, s.nsuper, s.xsuper, s.snode, s.xlindx, s.lindx, s.xlnz, s.lnz), "spk_ldltfactor_f32")
    s.errflag != 0 && error("An empty problem. No matrix.")
    return true
end
# the stateless Float32 entry points cache their plan per structure and keep the factors resident (spk_b200.cu: plan cache)
function SpkSparseBase._triangularsolve!(s::_SparseBase{Int64, Float32}, solution::AbstractVector{Float32})
    s.n == 0 && error("An empty problem. No solution.")
    rhs = solution[s.order.rperm]
    check(ccall((:spk_lulsolve_f32, libspk), Int64, (Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float32}, Ptr{Int64}, Ptr{Float32}),
        s.nsuper, s.xsuper, s.xlindx, s.lindx, s.xlnz, s.lnz, s.ipiv, rhs), "spk_lulsolve_f32")
    check(ccall((:spk_luusolve_f32, libspk), Int64, (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float32}, Ptr{Int64}, Ptr{Float32}, Ptr{Float32}),
        s.n, s.nsuper, s.xsuper, s.xlindx, s.lindx, s.xlnz, s.lnz, s.xunz, s.unz, rhs), "spk_luusolve_f32")
    solution .= rhs[s.order.rinvp]
    return true
end
function SpkSparseSpdBase._triangularsolve!(s::_SparseSpdBase{Int64, Float32}, solution::AbstractVector{Float32})
    s.n == 0 && error("An empty problem. No solution.")
    rhs = solution[s.order.rperm]
    check(ccall((:spk_ldltsolve_f32, libspk), Int64, (Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float32}, Ptr{Float32}),
        s.nsuper, s.xsuper, s.xlindx, s.lindx, s.xlnz, s.lnz, rhs), "spk_ldltsolve_f32")
    solution .= rhs[s.order.rinvp]
    return true
end

# ---- beyond the reference --------------------------------------------------------------------------------------
"""
    condest1(s, A) -> (cond, normA, norminvA, lowerbound)

1-norm condition estimate with the factors resident on the GPU (`spk_plan_condest`; the reference keeps its
estimator only as commented-out Fortran, SpkSparseSpdSolver.jl:267-459).  `A` is the SparseMatrixCSC that was factored.
"""
function condest1(s, A)
    p = plan!(s)
    check(ccall((:spk_plan_set_perm, libspk), Int64, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), p.h, s.order.rperm, s.order.rinvp), "spk_plan_set_perm")
    check(ccall((:spk_plan_set_matrix, libspk), Int64, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        p.h, length(A.nzval), Vector{Int64}(A.colptr), Vector{Int64}(A.rowval), Vector{Float64}(A.nzval)), "spk_plan_set_matrix")
    out = zeros(Float64, 2); info = zeros(Int32, 2)
    c = ccall((:spk_plan_condest, libspk), Float64, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}), p.h, out, info)
    c < 0 && error("spk_plan_condest: " * lasterror())
    return c, out[1], out[2], info[2] != 0
end

"""
    MultiGPU(s, ngpus)

One handle over `ngpus` GPUs of the box (`spk_multi_*`): elimination subtrees dealt to the GPUs, the top separators
distributed by column blocks, NCCL inside the library.  `factor!(m, s)` / `solve!(m, s, b)` mirror the plan calls.
"""
mutable struct MultiGPU
    h::Ptr{Cvoid}
end
function MultiGPU(s, ngpus::Integer)
    h = GC.@preserve s ccall((:spk_multi_create, libspk), Ptr{Cvoid},
        (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int32),
        s.n, s.nsuper, s.xsuper, s.snode, s.xlindx, s.lindx, s.xlnz, xunz_ptr(s), Int32(ngpus))
    h == C_NULL && error("spk_multi_create: " * lasterror())
    m = MultiGPU(h)
    finalizer(m -> (m.h != C_NULL && ccall((:spk_multi_destroy, libspk), Cvoid, (Ptr{Cvoid},), m.h); m.h = C_NULL), m)
    check(ccall((:spk_multi_set_perm, libspk), Int64, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h, s.order.rperm, s.order.rinvp), "spk_multi_set_perm")
    return m
end
function factor!(m::MultiGPU, s)
    unz = s isa _SparseBase ? pointer(s.unz) : Ptr{Float64}(C_NULL)
    ipiv = s isa _SparseBase ? pointer(s.ipiv) : Ptr{Int64}(C_NULL)
    GC.@preserve s begin
        check(ccall((:spk_multi_set_values, libspk), Int64, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m.h, s.lnz, unz), "spk_multi_set_values")
        s.errflag = check(ccall((:spk_multi_factor, libspk), Int64, (Ptr{Cvoid},), m.h), "spk_multi_factor")
        check(ccall((:spk_multi_get_factors, libspk), Int64, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), m.h, s.lnz, unz, ipiv), "spk_multi_get_factors")
    end
    return s.errflag == 0
end
solve!(m::MultiGPU, s, b::Vector{Float64}) =
    (check(ccall((:spk_multi_triangularsolve, libspk), Int64, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), m.h, b, 1, s.n), "spk_multi_triangularsolve"); b)

end # module
