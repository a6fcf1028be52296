# PseudospectraGPU.jl -- the reference-side binding for libpsagpu.so (NOT runnable in the build container:
# Julia is not installed there; kept deliberately thin so that review is enough).
#
# It adds a GPU option to Pseudospectra.jl's dense branch: when `opts[:gpu] == true`, the row-batched loop of
# `psa_compute` (src/compute.jl:199-219) is replaced by ONE call into the C ABI of include/psa_gpu.h; everything
# before (options, grid, projection: compute.jl:94-153) and after (mirror, clamp, levels: compute.jl:222-267)
# is the reference's own code, and the returned tuple is unchanged, so new_matrix / driver! / spectralportrait
# and the plotting extensions need no change.
module PseudospectraGPU

using LinearAlgebra
import Pseudospectra

const libpsagpu = get(ENV, "PSAGPU_LIB", "libpsagpu.so")

const PSA_OK = Cint(0)
const PSA_ERR_CANCELLED = Cint(-3)
const PSA_ERR_UNSUPPORTED = Cint(-4)
const ALGO = Dict(Cint(1) => :sq_lanc, Cint(2) => :HessQR, Cint(3) => :rect_qr, Cint(4) => :SVD)

mutable struct Context
    ptr::Ptr{Cvoid}
    ngpus::Int
    shape::Tuple{Int,Int}
end

function Context(ngpus::Integer=1)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:psa_gpu_init, libpsagpu), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), ngpus, C_NULL, r)
    st == PSA_OK || error("psa_gpu_init: ", unsafe_string(ccall((:psa_gpu_strerror, libpsagpu), Cstring, (Cint,), st)))
    ctx = Context(r[], ngpus, (0, 0))
    finalizer(c -> ccall((:psa_gpu_free, libpsagpu), Cint, (Ptr{Cvoid},), c.ptr), ctx)
    return ctx
end

const _ctx = Ref{Union{Nothing,Context}}(nothing)
context(ngpus) = (_ctx[] === nothing || _ctx[].ngpus != ngpus) ? (_ctx[] = Context(ngpus)) : _ctx[]

"""
    psacore_gpu(T, q0, x, y; tol, maxit, ngpus, ctrlflag) -> (Z, algo, counts) or `nothing` (unsupported / cancelled)

GPU replacement of `psacore(T, I, q0, x, y, bw)` (src/compute.jl:448-617) for the whole grid at once.
"""
function psacore_gpu(T::AbstractMatrix, q0::Vector{ComplexF64}, x::Vector{Float64}, y::Vector{Float64};
                     tol=1e-5, maxit=99, ngpus=1, ctrlflag=nothing)
    Td = Matrix{ComplexF64}(parent(T) isa Matrix ? parent(T) : Matrix(T))   # compute.jl:450-454 (complexify)
    m, n = size(Td)
    ctx = context(ngpus)
    st = ccall((:psa_gpu_set_matrix, libpsagpu), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64, Cint, Cint),
               ctx.ptr, Td, stride(Td, 2), m, n)
    st == PSA_ERR_UNSUPPORTED && return nothing           # n < 55, rect_qr, ...: fall through to the stock code
    st == PSA_OK || error("psa_gpu_set_matrix: status $st")
    lx, ly = length(x), length(y)
    Z = Matrix{Float64}(undef, ly, lx)
    counts = zeros(Int64, 4)
    algo = Ref{Cint}(0)
    cancel = ctrlflag === nothing ? Ptr{Cint}(C_NULL) : Base.unsafe_convert(Ptr{Cint}, ctrlflag)
    bigσ = 0.1 * floatmax(Float64)                        # compute.jl:475
    GC.@preserve Td q0 x y Z counts ctrlflag begin
        st = ccall((:psa_gpu_grid, libpsagpu), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{ComplexF64}, Float64, Cint, Float64,
                    Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ref{Cint}, Ptr{Cint}),
                   ctx.ptr, x, lx, y, ly, q0, tol, maxit, bigσ, Z, C_NULL, counts, algo, cancel)
    end
    st == PSA_ERR_CANCELLED && return nothing
    st == PSA_OK || error("psa_gpu_grid: status $st")
    return Z, ALGO[algo[]], counts
end

# The patch a maintainer applies inside psa_compute, replacing src/compute.jl:199-219 when the option is set:
#
#     if get(all_opts, :gpu, false) && isa(S, UniformScaling)
#         q = randn(n) + randn(n)*im; q = q / norm(q)                      # as compute.jl:207-208, once
#         r = PseudospectraGPU.psacore_gpu(Tproj, q, x, y; tol=psatol, maxit=thresholds.maxit_lancs,
#                                          ngpus=get(all_opts, :ngpus, 1), ctrlflag=ctrlflag)
#         if r !== nothing
#             Z, algo, counts = r
#             counts[1] > 0 && @warn "Lanczos convergence failure(s) while computing resolvent norms"
#             counts[2] > 0 && @warn "Eigenvalue convergence failure(s) while computing resolvent norms"
#             @goto postprocess                                             # -> compute.jl:222
#         elseif (ctrlflag !== nothing) && (ctrlflag[] == 1)
#             return nothing                                                # compute.jl:202-204
#         end
#     end
end # module
