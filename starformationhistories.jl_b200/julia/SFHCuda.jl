# SFHCuda.jl -- the reference-side binding of libsfhcuda.so (include/sfhcuda.h).
#
# NOT EXECUTABLE IN THIS IMAGE (no julia binary); shipped as the thin `ccall` layer a maintainer of
# StarFormationHistories.jl would add.  Every method below keeps the exact signature and return
# convention of the reference method it re-bodies (file:line in comments); dispatch on `DeviceStack`
# selects the GPU path, plain `Matrix` arguments keep the reference's CPU behaviour.
# The Python ctypes binding (starformationhistories.jl_b200/_lib.py) exercises the identical symbols.
module SFHCuda

import StarFormationHistories as SFH
using StarFormationHistories: AbstractMZR, AbstractAMR, PowerLawMZR, LinearAMR, LogarithmicAMR,
                              GaussianDispersion, fittable_params, free_params
using TaskLocalValues: TaskLocalValue

const libsfh = get(ENV, "LIBSFHCUDA", "libsfhcuda.so")
const SFH_F32, SFH_F64, SFH_I64 = Cint(0), Cint(1), Cint(2)

struct SFHError <: Exception
    status::Cint
    msg::String
end
function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:sfh_last_error, libsfh), Cstring, ()))
    # SFH_ERR_INVALID_ARG / SFH_ERR_SHAPE mirror the reference's @argcheck ArgumentErrors
    (status == 1 || status == 2) ? throw(ArgumentError(msg)) : throw(SFHError(status, msg))
end

dtype_code(::Type{Float32}) = SFH_F32
dtype_code(::Type{Float64}) = SFH_F64
dtype_code(::Type{Int64}) = SFH_I64

# ---- the device mirror of stack_models (src/fitting/utilities.jl:12-13) -------------------------
mutable struct DeviceStack{S} <: AbstractMatrix{S}
    host::Matrix{S}             # kept so that size/getindex and CPU-only helpers keep working
    handle::Ptr{Cvoid}
    ctx::TaskLocalValue{Ptr{Cvoid}}   # one sfh_ctx per task, like HMCModel's TaskLocalValue (hmc_sample.jl:127)
    function DeviceStack(models::Matrix{S}, data::AbstractVector{D}) where {S <: Union{Float32, Float64}, D}
        size(models, 1) == length(data) || throw(ArgumentError("axes(models,1) != axes(data,1)"))
        d = D <: Union{Float32, Float64, Int64} ? collect(data) : Float64.(data)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve models d check(ccall((:sfh_stack_create, libsfh), Cint,
            (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
            h, models, size(models, 1), size(models, 2), dtype_code(S), d, dtype_code(eltype(d)), C_NULL))
        handle = h[]
        ctx = TaskLocalValue{Ptr{Cvoid}}() do
            c = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:sfh_ctx_create, libsfh), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), handle, C_NULL, c))
            c[]
        end
        obj = new{S}(models, handle, ctx)
        finalizer(o -> ccall((:sfh_stack_destroy, libsfh), Cint, (Ptr{Cvoid},), o.handle), obj)
        return obj
    end
end
Base.size(s::DeviceStack) = size(s.host)
Base.getindex(s::DeviceStack, i...) = getindex(s.host, i...)
DeviceStack(models::AbstractVector{<:AbstractMatrix}, data::AbstractMatrix) = DeviceStack(SFH.stack_models(models), vec(data))

# ---- composite!(composite, coeffs, models)   src/fitting/fitting_base.jl:55-65 --------------------
function SFH.composite!(composite::AbstractVector{<:Number}, coeffs::AbstractVector{<:Number}, models::DeviceStack)
    axes(composite, 1) == axes(models, 1) || throw(ArgumentError("axes(composite,1) != axes(models,1)"))
    axes(coeffs, 1) == axes(models, 2) || throw(ArgumentError("axes(coeffs,1) != axes(models,2)"))
    x = convert(Vector{Float64}, coeffs); out = Vector{Float64}(undef, length(composite))
    check(ccall((:sfh_composite, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), models.ctx[], x, out))
    composite .= out
    return
end

# ---- loglikelihood(coeffs, models, data)      src/fitting/fitting_base.jl:117-125 -----------------
function SFH.loglikelihood(coeffs::AbstractVector{<:Number}, models::DeviceStack{S}, data::AbstractVector{<:Number}) where S
    x = convert(Vector{Float64}, coeffs); r = Ref{Float64}()
    check(ccall((:sfh_loglikelihood_coeffs, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), models.ctx[], x, r))
    return convert(S, r[])      # reference returns the promoted eltype (fitting_core_test.jl:40)
end

# ---- ∇loglikelihood!(G, composite, models, data)   src/fitting/fitting_base.jl:265-285 ------------
function SFH.∇loglikelihood!(G::AbstractVector, composite::AbstractVector{<:Number}, models::DeviceStack, data::AbstractVector{<:Number})
    c = convert(Vector{Float64}, composite); g = Vector{Float64}(undef, length(G))
    check(ccall((:sfh_grad_loglikelihood, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), models.ctx[], c, g))
    composite .= c      # the documented side effect: composite now holds 1 - n/m (:219)
    G .= g
    return G
end

# ---- fg!(F, G, coeffs, models, data, composite)    src/fitting/solvers.jl:20-38 -------------------
function SFH.fg!(F, G, coeffs::AbstractVector{<:Number}, models::DeviceStack{S}, data::AbstractVector{<:Number},
                 composite::AbstractVector{<:Number}) where S
    x = convert(Vector{Float64}, coeffs)
    nl = Ref{Float64}()
    g = G === nothing ? C_NULL : Vector{Float64}(undef, length(x))
    check(ccall((:sfh_eval_fg, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                models.ctx[], x, F === nothing ? C_NULL : nl, g, C_NULL))
    G === nothing || (G .= g)
    return F === nothing ? nothing : convert(S, nl[])
end

# ---- hierarchical fg!   mzr.jl:84-215 ("fg_mzr!") / amr.jl:78-173 ("fg_amr!") ---------------------
mh_kind(::PowerLawMZR) = Cint(0); mh_fixed(m::PowerLawMZR) = Float64[m.logMstar0, 0, 0, 0]
mh_kind(::LinearAMR) = Cint(1);   mh_fixed(m::LinearAMR) = Float64[m.T_max, 0, 0, 0]
mh_kind(::LogarithmicAMR) = Cint(2); mh_fixed(m::LogarithmicAMR) = Float64[m.T_max, 0.01524, 0.2485, 1.78]

const _bound = IdDict{Any, Tuple{Vector{Float64}, Vector{Float64}}}()   # ctx -> (logAge, MH) already bound
function bind!(models::DeviceStack, logAge, MH)
    c = models.ctx[]
    la, mh = convert(Vector{Float64}, logAge), convert(Vector{Float64}, MH)
    get(_bound, c, nothing) == (la, mh) && return c
    n = Ref{Int64}()
    check(ccall((:sfh_hier_bind, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}), c, la, mh, n))
    _bound[c] = (la, mh)
    return c
end

function SFH.fg!(F, G, MHmodel0::Union{PowerLawMZR, LinearAMR, LogarithmicAMR}, dispmodel0::GaussianDispersion,
                 variables::AbstractVector{<:Number}, models::DeviceStack, data, composite,
                 logAge::AbstractVector{<:Number}, metallicities::AbstractVector{<:Number})
    c = bind!(models, logAge, metallicities)
    v = convert(Vector{Float64}, variables)
    free = UInt8[free_params(MHmodel0)..., free_params(dispmodel0)..., 0]
    nl = Ref{Float64}()
    g = G === nothing ? C_NULL : Vector{Float64}(undef, length(v))
    check(ccall((:sfh_eval_fg_hier, libsfh), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{UInt8}, Ref{Float64}, Ptr{Float64}),
                c, mh_kind(MHmodel0), mh_fixed(MHmodel0), Cint(0), v, free, nl, g))
    G === nothing || (G .= g)
    return F === nothing ? nothing : nl[]
end

# ---- MCMCModel: W walkers per call   src/fitting/mcmc_sample.jl:12-23 ------------------------------
function batched_loglikelihood(models::DeviceStack, X::Matrix{Float64})
    out = Vector{Float64}(undef, size(X, 2))
    check(ccall((:sfh_eval_logl_batched, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}),
                models.ctx[], X, size(X, 2), out))
    return out
end
(problem::SFH.MCMCModel{<:DeviceStack})(θ) = batched_loglikelihood(problem.models, reshape(convert(Vector{Float64}, θ), :, 1))[1]

end # module
