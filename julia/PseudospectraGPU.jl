# PseudospectraGPU.jl -- the reference-side binding for libpsagpu.so (NOT runnable in the build container:
# Julia is not installed there; every ccall below is exercised from plain C by tests/abi_driver.c and from
# Python/ctypes by the test-suite instead, so the shim is kept thin enough that review is enough).
#
# It adds a GPU option to Pseudospectra.jl's dense branch: when `opts[:gpu] == true`, the row-batched loop of
# `psa_compute` (src/compute.jl:199-219) is replaced by ONE call into the C ABI of include/psa_gpu.h; everything
# before (options, grid, projection: compute.jl:94-153) is the reference's own code, the post-processing
# (mirror, clamp, level statistics: compute.jl:222-267) can run on the device, and the returned tuple is unchanged, so
# new_matrix / driver! / spectralportrait and the plotting extensions need no change.
module PseudospectraGPU

using LinearAlgebra
import Pseudospectra

const libpsagpu = get(ENV, "PSAGPU_LIB", "libpsagpu.so")

const PSA_OK = Cint(0)
const PSA_ERR_CANCELLED = Cint(-3)
const PSA_ERR_UNSUPPORTED = Cint(-4)
const PSA_MATRIX_UPPER_WRAPPER = Cint(1)
const ALGO = Dict(Cint(1) => :sq_lanc, Cint(2) => :HessQR, Cint(3) => :rect_qr, Cint(4) => :SVD)

strerror(st) = unsafe_string(ccall((:psa_gpu_strerror, libpsagpu), Cstring, (Cint,), st))

"""
    plan(m, n, maxit=99) -> (algo::Union{Symbol,Nothing}, points_in_flight_per_sm::Int)

Which machine a grid call would use for an m x n matrix (host logic only, no device needed): the `psacore` branch and,
for small square matrices, the number of grid points one SM keeps in flight in the single-launch point-resident kernel
(0: tick engine / HessQR / SVD kernels).
"""
function plan(m::Integer, n::Integer, maxit::Integer=99)
    w = Ref{Cint}(0)
    a = ccall((:psa_gpu_plan, libpsagpu), Cint, (Cint, Cint, Cint, Ref{Cint}), m, n, maxit, w)
    return get(ALGO, a, nothing), Int(w[])
end

mutable struct Context
    ptr::Ptr{Cvoid}
    ngpus::Int
    # identity of the resident matrix: a zoom (src/zooming.jl:24-99 -> driver!) re-enters psa_compute with the SAME
    # Schur factor, so the upload and the packing are skipped when nothing changed
    matrix_id::UInt
    matrix_size::Tuple{Int,Int}
    matrix_probe::ComplexF64
    thresholds::Tuple{Int,Int}
end

function Context(ngpus::Integer=1)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:psa_gpu_init, libpsagpu), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), ngpus, C_NULL, r)
    st == PSA_OK || error("psa_gpu_init: ", strerror(st))
    ctx = Context(r[], ngpus, UInt(0), (0, 0), 0.0im, (0, 0))
    finalizer(c -> ccall((:psa_gpu_free, libpsagpu), Cint, (Ptr{Cvoid},), c.ptr), ctx)
    return ctx
end

const _ctx = Ref{Union{Nothing,Context}}(nothing)
context(ngpus) = (_ctx[] === nothing || _ctx[].ngpus != ngpus) ? (_ctx[] = Context(ngpus)) : _ctx[]

# The dense column-major ComplexF64 array behind the argument, and whether the strict lower part is to be ignored.
# Only the shapes the reference itself hands to psacore are accepted; anything else (Adjoint, Transpose, sparse, ...)
# returns `nothing` so that the caller falls through to the stock code.
_dense(T::UpperTriangular{ComplexF64,Matrix{ComplexF64}}) = (parent(T), true)
_dense(T::UpperTriangular{<:Number,<:Matrix}) = (Matrix{ComplexF64}(parent(T)), true)   # compute.jl:450-454 (complexify)
_dense(T::Matrix{ComplexF64}) = (T, false)
_dense(T::Matrix{<:Number}) = (Matrix{ComplexF64}(T), false)
_dense(T) = nothing

_probe(Td) = sum(diag(Td)) + sum(@view Td[1, :]) + sum(@view Td[:, end])

"""
    set_matrix!(ctx, T, thresholds) -> Bool

Upload + pack `T` unless the very same matrix is already resident.  `false` = shape not covered by the GPU path.
"""
function set_matrix!(ctx::Context, T, thresholds)
    dw = _dense(T)
    dw === nothing && return false
    Td, wrapper = dw
    th = (thresholds.minlancs4psa, thresholds.maxstdqr4hess)
    id, pr = objectid(parent(T) isa Matrix ? parent(T) : T), _probe(Td)
    if ctx.matrix_id == id && ctx.matrix_size == size(Td) && ctx.matrix_probe == pr && ctx.thresholds == th
        return true
    end
    ctx.matrix_id = UInt(0)
    ccall((:psa_gpu_set_thresholds, libpsagpu), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.ptr, th[1], th[2])
    m, n = size(Td)
    st = GC.@preserve Td ccall((:psa_gpu_set_matrix_ex, libpsagpu), Cint,
                               (Ptr{Cvoid}, Ptr{ComplexF64}, Int64, Cint, Cint, Cint),
                               ctx.ptr, Td, stride(Td, 2), m, n, wrapper ? PSA_MATRIX_UPPER_WRAPPER : Cint(0))
    st == PSA_ERR_UNSUPPORTED && return false      # generalised / non-Hessenberg / not upper triangular: host path
    st == PSA_OK || error("psa_gpu_set_matrix_ex: ", strerror(st))
    ctx.matrix_id, ctx.matrix_size, ctx.matrix_probe, ctx.thresholds = id, size(Td), pr, th
    return true
end

"""
    psacore_gpu(T, qs, row_batch, x, y; tol, maxit, ngpus, ctrlflag, thresholds, n_mirror)
        -> (Z, yfull, zstats, algo, counts) or `nothing` (unsupported shape / cancelled)

GPU replacement of the whole row-batched `psacore` loop (src/compute.jl:199-219 + 448-617) followed by the
device-side un-mirror / clamp (222-238).  `qs` is n x nbatch (one start vector per row batch, drawn by the caller in the
reference's order), `row_batch[j]` the 0-based batch of grid row j.
"""
function psacore_gpu(T, qs::Matrix{ComplexF64}, row_batch::Vector{Int32}, x::Vector{Float64}, y::Vector{Float64};
                     tol=1e-5, maxit=99, ngpus=1, ctrlflag=nothing,
                     thresholds=Pseudospectra.psathresholds, n_mirror::Integer=0, resident::Bool=false)
    ctx = context(ngpus)
    # resident=true: the context's current matrix is used as it is (e.g. right after project_gpu!)
    resident || set_matrix!(ctx, T, thresholds) || return nothing
    lx, ly = length(x), length(y)
    counts = zeros(Int64, 4)
    algo = Ref{Cint}(0)
    # The library polls a C int; the reference's flag is any Ref-like with `ctrlflag[] == 1` (compute.jl:202-204).
    # A Ref{Cint}/Vector{Cint} is passed through; anything else is mirrored into a Cint cell by a watcher task.
    cell = ctrlflag isa Union{Base.RefValue{Cint},Vector{Cint}} ? ctrlflag : Ref{Cint}(0)
    watcher = (ctrlflag === nothing || cell === ctrlflag) ? nothing :
              Threads.@spawn while cell[] == 0
                  ctrlflag[] == 1 && (cell[] = Cint(1))
                  sleep(0.05)
              end
    bigσ = 0.1 * floatmax(Float64)                        # compute.jl:475
    st = GC.@preserve qs row_batch x y counts cell algo begin
        cptr = ctrlflag === nothing ? Ptr{Cint}(C_NULL) :
               (cell isa Vector ? pointer(cell) : Base.unsafe_convert(Ptr{Cint}, cell))
        # @threadcall keeps the Julia scheduler (and the watcher) alive during the blocking call
        @threadcall((:psa_gpu_grid_batched, libpsagpu), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{ComplexF64}, Cint, Ptr{Int32}, Float64, Cint,
                     Float64, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Cint}, Ptr{Cint}),
                    ctx.ptr, x, lx, y, ly, qs, size(qs, 2), row_batch, tol, maxit, bigσ, C_NULL, C_NULL, counts,
                    Base.unsafe_convert(Ptr{Cint}, algo), cptr)
    end
    watcher === nothing || (cell[] == 0 && (cell[] = Cint(2)))   # stop the watcher
    st == PSA_ERR_CANCELLED && return nothing
    st == PSA_OK || error("psa_gpu_grid_batched: ", strerror(st))
    nm = abs(n_mirror)
    Z = Matrix{Float64}(undef, ly + nm, lx)
    yfull = Vector{Float64}(undef, ly + nm)
    zstats = zeros(3)
    ps_tiny = 10 * sqrt(floatmin(Float64))                # compute.jl:236
    st = GC.@preserve y Z yfull zstats ccall((:psa_gpu_postprocess, libpsagpu), Cint,
              (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Cint, Float64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
              ctx.ptr, y, ly, lx, n_mirror, ps_tiny, Pseudospectra.smallσ, Z, ly + nm, yfull, zstats)
    st == PSA_OK || error("psa_gpu_postprocess: ", strerror(st))
    return Z, yfull, zstats, ALGO[algo[]], counts
end

"""
    project_gpu!(ctx, T, selection, thresholds) -> Tproj or `nothing`

The Givens reordering + truncation of `_maybe_project` (src/compute.jl:339-357) on the device, as a wavefront of the
reference's own swaps.  `selection` is what the host logic of compute.jl:295-316 produced (1-based, ascending).  The
projected matrix becomes the current matrix of the context: call the following `psacore_gpu` with `resident=true`.
The full matrix stays resident too, so the next zoom can project again (or `selection = 1:n` restores it).
"""
function project_gpu!(ctx::Context, T, selection::Vector{Int}, thresholds=Pseudospectra.psathresholds)
    set_matrix!(ctx, T, thresholds) || return nothing
    np = length(selection)
    sel0 = Int32.(selection .- 1)
    Tproj = Matrix{ComplexF64}(undef, np, np)
    st = GC.@preserve sel0 Tproj ccall((:psa_gpu_project, libpsagpu), Cint,
              (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{ComplexF64}, Int64), ctx.ptr, sel0, np, Tproj, np)
    st == PSA_ERR_UNSUPPORTED && return nothing
    st == PSA_OK || error("psa_gpu_project: ", strerror(st))
    return UpperTriangular(Tproj)
end

"""
    psmode_gpu(T, q0, z, tol, num_its) -> (Z, q) or `nothing`

`psmode_inv_lanczos` (src/modes.jl:185-274) for the dense n >= 200 branch on the resident Schur factor.
"""
function psmode_gpu(T, q0::Vector{ComplexF64}, z::Complex, tol, num_its; ngpus=1, thresholds=Pseudospectra.psathresholds)
    ctx = context(ngpus)
    set_matrix!(ctx, T, thresholds) || return nothing
    n = length(q0)
    sig, its, info = Ref{Float64}(0), Ref{Cint}(0), Ref{Cint}(0)
    mode = Vector{ComplexF64}(undef, n)
    st = GC.@preserve q0 mode ccall((:psa_gpu_pseudomode, libpsagpu), Cint,
              (Ptr{Cvoid}, Float64, Float64, Ptr{ComplexF64}, Float64, Cint, Ref{Float64}, Ptr{ComplexF64}, Ref{Cint}, Ref{Cint}),
              ctx.ptr, real(z), imag(z), q0, tol, num_its, sig, mode, its, info)
    st == PSA_ERR_UNSUPPORTED && return nothing           # n < 200: the reference's plain-SVD branch (modes.jl:193-200)
    st == PSA_OK || error("psa_gpu_pseudomode: ", strerror(st))
    info[] == 0 || return (NaN, NaN)                      # modes.jl:262-272 (caller may fall back to the dense SVD)
    return sig[], mode
end

# The patch a maintainer applies inside psa_compute, replacing src/compute.jl:199-238 when the option is set
# (`step`, `x`, `y`, `n`, `m`, `Tproj`, `n_mirror_pts`, `psatol`, `thresholds`, `ctrlflag` are the function's locals):
#
#     if get(all_opts, :gpu, false) && isa(S, UniformScaling)
#         nb = cld(ly, step)
#         qs = Matrix{ComplexF64}(undef, size(Tproj, 2), nb)
#         for b in 1:nb                                                      # same RNG draws, same order as 207-208
#             q = randn(n) + randn(n)*im; q = q / norm(q)
#             qs[:, b] = q[1:size(Tproj, 2)]                                  # `q = q0[1:n]`, compute.jl:622
#         end
#         row_batch = Int32[(ly - j) ÷ step for j in 1:ly]                    # batch of row j in `for j=ly:-step:1`
#         r = PseudospectraGPU.psacore_gpu(Tproj, qs, row_batch, x, y; tol=psatol, maxit=thresholds.maxit_lancs,
#                                          ngpus=get(all_opts, :ngpus, 1), ctrlflag=ctrlflag,
#                                          n_mirror=n_mirror_pts)
#         if r !== nothing
#             Z, y, zstats, algo, counts = r
#             counts[1] > 0 && @warn "Lanczos convergence failure(s) while computing resolvent norms"
#             counts[2] > 0 && @warn "Eigenvalue convergence failure(s) while computing resolvent norms"
#             @goto levels                                                  # -> compute.jl:240 (mirror + clamp done on device)
#         elseif (ctrlflag !== nothing) && (ctrlflag[] == 1)
#             return nothing                                                # compute.jl:202-204
#         end
#     end
end # module
