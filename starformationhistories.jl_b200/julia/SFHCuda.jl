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
    host::Union{Nothing, Matrix{S}}   # kept (when the stack was uploaded) so that getindex and CPU-only helpers keep working
    dims::Tuple{Int, Int}
    handle::Ptr{Cvoid}
    # Contexts (sfh_ctx: stream + device scratch + pinned buffers) are POOLED per stack: a caller checks one out for the duration
    # of a call (`with_ctx`), so the pool grows to the number of CONCURRENT callers and no further -- tsample_sfh spawns one
    # task per short chain (generic_fitting.jl:617-626), and a context per task (the TaskLocalValue of hmc_sample.jl:127) would
    # leave thousands of them alive.  The library does not track contexts: the finalizer destroys them, then the stack.
    lock::ReentrantLock
    idle::Vector{Ptr{Cvoid}}
    all::Vector{Ptr{Cvoid}}
    bound::Dict{Ptr{Cvoid}, Tuple{Vector{Float64}, Vector{Float64}}}   # ctx -> the (logAge, MH) grid sfh_hier_bind gave it
    # Multi-GPU from this ONE process (sfh_group_*): `group` owns one shard per GPU and a single PRIMARY context whose evaluations
    # fan out to every GPU and come back all-reduced; callers take turns on it (`grouplock`).  handle == C_NULL then.
    group::Ptr{Cvoid}
    primary::Ptr{Cvoid}
    grouplock::ReentrantLock
    function DeviceStack{S}(host, dims, handle::Ptr{Cvoid}, group::Ptr{Cvoid}=C_NULL, primary::Ptr{Cvoid}=C_NULL) where S
        obj = new{S}(host, dims, handle, ReentrantLock(), Ptr{Cvoid}[], Ptr{Cvoid}[],
                     Dict{Ptr{Cvoid}, Tuple{Vector{Float64}, Vector{Float64}}}(), group, primary, ReentrantLock())
        finalizer(obj) do o   # o is unreachable: nobody holds its lock or one of its contexts any more
            if o.group != C_NULL
                ccall((:sfh_group_destroy, libsfh), Cint, (Ptr{Cvoid},), o.group)   # releases the primary context and every shard
            else
                foreach(c -> ccall((:sfh_ctx_destroy, libsfh), Cint, (Ptr{Cvoid},), c), o.all)
                ccall((:sfh_stack_destroy, libsfh), Cint, (Ptr{Cvoid},), o.handle)
            end
        end
        return obj
    end
end
function new_ctx!(s::DeviceStack)   # caller holds s.lock
    c = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sfh_ctx_create, libsfh), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), s.handle, C_NULL, c))
    push!(s.all, c[])
    return c[]
end
# f(ctx) with a context nobody else is using; safe from any number of tasks / threads (one ctx <=> one concurrent caller)
function with_ctx(f, s::DeviceStack)
    s.group != C_NULL && return lock(() -> f(s.primary), s.grouplock)   # one evaluation at a time spans all the group's GPUs
    c = lock(() -> isempty(s.idle) ? new_ctx!(s) : pop!(s.idle), s.lock)
    try
        return f(c)
    finally
        lock(() -> push!(s.idle, c), s.lock)
    end
end
function DeviceStack(models::Matrix{S}, data::AbstractVector{D}) where {S <: Union{Float32, Float64}, D}
    size(models, 1) == length(data) || throw(ArgumentError("axes(models,1) != axes(data,1)"))
    d = D <: Union{Float32, Float64, Int64} ? collect(data) : Float64.(data)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve models d check(ccall((:sfh_stack_create, libsfh), Cint,
        (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        h, models, size(models, 1), size(models, 2), dtype_code(S), d, dtype_code(eltype(d)), C_NULL))
    return DeviceStack{S}(models, size(models), h[])
end
# The same stack sharded by bin rows over several GPUs of THIS process -- what fit_sfh / fit_templates need for a stack that does
# not fit one GPU (every reference caller is a single process: generic_fitting.jl:242-411, solvers.jl:82-90).  No launcher, no
# NCCL: the library splits the rows, enables peer access and all-reduces [logL, G] inside its finalize kernel.
function DeviceStack(models::Matrix{S}, data::AbstractVector{D}, devices::AbstractVector{<:Integer}) where {S <: Union{Float32, Float64}, D}
    size(models, 1) == length(data) || throw(ArgumentError("axes(models,1) != axes(data,1)"))
    d = D <: Union{Float32, Float64, Int64} ? collect(data) : Float64.(data)
    devs = convert(Vector{Cint}, devices)
    g = Ref{Ptr{Cvoid}}(C_NULL); c = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve models d devs check(ccall((:sfh_group_create, libsfh), Cint,
        (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Cvoid}, Cint, Ptr{Cint}, Cint, Ptr{Cvoid}),
        g, models, size(models, 1), size(models, 2), dtype_code(S), d, dtype_code(eltype(d)), devs, Cint(length(devs)), C_NULL))
    check(ccall((:sfh_group_ctx, libsfh), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), g[], c))
    return DeviceStack{S}(models, size(models), C_NULL, g[], c[])
end
Base.size(s::DeviceStack) = s.dims
Base.getindex(s::DeviceStack, i...) = getindex(s.host, i...)
DeviceStack(models::AbstractVector{<:AbstractMatrix}, data::AbstractMatrix) = DeviceStack(SFH.stack_models(models), vec(data))

# ---- composite!(composite, coeffs, models)   src/fitting/fitting_base.jl:55-65 --------------------
function SFH.composite!(composite::AbstractVector{<:Number}, coeffs::AbstractVector{<:Number}, models::DeviceStack)
    axes(composite, 1) == axes(models, 1) || throw(ArgumentError("axes(composite,1) != axes(models,1)"))
    axes(coeffs, 1) == axes(models, 2) || throw(ArgumentError("axes(coeffs,1) != axes(models,2)"))
    x = convert(Vector{Float64}, coeffs); out = Vector{Float64}(undef, length(composite))
    with_ctx(c -> check(ccall((:sfh_composite, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), c, x, out)), models)
    composite .= out
    return
end

# ---- loglikelihood(coeffs, models, data)      src/fitting/fitting_base.jl:117-125 -----------------
function SFH.loglikelihood(coeffs::AbstractVector{<:Number}, models::DeviceStack{S}, data::AbstractVector{<:Number}) where S
    x = convert(Vector{Float64}, coeffs); r = Ref{Float64}()
    with_ctx(c -> check(ccall((:sfh_loglikelihood_coeffs, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), c, x, r)), models)
    return convert(S, r[])      # reference returns the promoted eltype (fitting_core_test.jl:40)
end

# ---- ∇loglikelihood!(G, composite, models, data)   src/fitting/fitting_base.jl:265-285 ------------
function SFH.∇loglikelihood!(G::AbstractVector, composite::AbstractVector{<:Number}, models::DeviceStack, data::AbstractVector{<:Number})
    c = convert(Vector{Float64}, composite); g = Vector{Float64}(undef, length(G))
    with_ctx(ctx -> check(ccall((:sfh_grad_loglikelihood, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx, c, g)), models)
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
    with_ctx(models) do c
        check(ccall((:sfh_eval_fg, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    c, x, F === nothing ? C_NULL : nl, g, C_NULL))
    end
    G === nothing || (G .= g)
    return F === nothing ? nothing : convert(S, nl[])
end

# ---- hierarchical fg!   mzr.jl:84-215 ("fg_mzr!") / amr.jl:78-173 ("fg_amr!") ---------------------
# Metallicity models whose formulae the library's prologue / epilogue kernels hold.  LogarithmicAMR carries its Z -> [M/H]
# conversion as two callables (amr.jl:250-256): only the default pair (MH_from_Z, dMH_dZ with solZ = 0.01524, Y_p = 0.2485,
# gamma = 1.78; src/utilities.jl:138-156) is what the device evaluates, so the device methods dispatch on exactly that
# instantiation (functions are singleton types).  Any other AbstractMZR / AbstractAMR / dispersion model -- a LogarithmicAMR with
# user-supplied conversions included -- falls through to the reference's own generic fg! (mzr.jl:84 / amr.jl:78), which reaches the
# device through composite! / loglikelihood / ∇loglikelihood! on the DeviceStack and does the chain rule on the host.
const DeviceLogAMR = LogarithmicAMR{<:Real, typeof(SFH.MH_from_Z), typeof(SFH.dMH_dZ)}
const DeviceMH = Union{PowerLawMZR, LinearAMR, DeviceLogAMR}
mh_kind(::PowerLawMZR) = Cint(0); mh_fixed(m::PowerLawMZR) = Float64[m.logMstar0, 0, 0, 0]
mh_kind(::LinearAMR) = Cint(1);   mh_fixed(m::LinearAMR) = Float64[m.T_max, 0, 0, 0]
mh_kind(::LogarithmicAMR) = Cint(2); mh_fixed(m::LogarithmicAMR) = Float64[m.T_max, 0.01524, 0.2485, 1.78]

# sfh_hier_bind once per (context, grid): the grid a context is bound to is remembered in the stack (under its lock)
function bind!(s::DeviceStack, c::Ptr{Cvoid}, logAge, MH)
    la, mh = Vector{Float64}(logAge), Vector{Float64}(MH)   # copies: the cache must not alias arrays the caller may change
    lock(() -> get(s.bound, c, nothing), s.lock) == (la, mh) && return c
    n = Ref{Int64}()
    check(ccall((:sfh_hier_bind, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}), c, la, mh, n))
    lock(() -> (s.bound[c] = (la, mh)), s.lock)
    return c
end

function SFH.fg!(F, G, MHmodel0::DeviceMH, dispmodel0::GaussianDispersion,
                 variables::AbstractVector{<:Number}, models::DeviceStack,
                 data::Union{AbstractVector{<:Number},AbstractMatrix{<:Number}},
                 composite::Union{AbstractVector{<:Number},AbstractMatrix{<:Number}},
                 logAge::AbstractVector{<:Number}, metallicities::AbstractVector{<:Number})
    v = convert(Vector{Float64}, variables)
    free = UInt8[free_params(MHmodel0)..., free_params(dispmodel0)..., 0]
    nl = Ref{Float64}()
    g = G === nothing ? C_NULL : Vector{Float64}(undef, length(v))
    with_ctx(models) do c
        bind!(models, c, logAge, metallicities)
        check(ccall((:sfh_eval_fg_hier, libsfh), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{UInt8}, Ref{Float64}, Ptr{Float64}),
                    c, mh_kind(MHmodel0), mh_fixed(MHmodel0), Cint(0), v, free, nl, g))
    end
    G === nothing || (G .= g)
    return F === nothing ? nothing : nl[]
end

# ---- MCMCModel: W walkers per call   src/fitting/mcmc_sample.jl:12-23 ------------------------------
function batched_loglikelihood(models::DeviceStack, X::Matrix{Float64})
    out = Vector{Float64}(undef, size(X, 2))
    with_ctx(models) do c
        check(ccall((:sfh_eval_logl_batched, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), c, X, size(X, 2), out))
    end
    return out
end
(problem::SFH.MCMCModel{<:DeviceStack})(θ) = batched_loglikelihood(problem.models, reshape(convert(Vector{Float64}, θ), :, 1))[1]

# fg! for C coefficient vectors at once: what the chain threads of hmc_sample / sample_sfh evaluate one by one
# (hmc_sample.jl:123-141, generic_fitting.jl:617-626).  X is ntemplates x C; returns (-logL[C], G[ntemplates, C]).
function batched_fg(models::DeviceStack, X::Matrix{Float64}; want_G::Bool=true)
    nl = Vector{Float64}(undef, size(X, 2))
    G = want_G ? Matrix{Float64}(undef, size(X)) : nothing
    with_ctx(models) do c
        check(ccall((:sfh_eval_fg_batched, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                    c, X, size(X, 2), nl, want_G ? G : C_NULL))
    end
    return nl, G
end

# Hierarchical fg! for C variable vectors at once (every chain task of tsample_sfh, generic_fitting.jl:617-626).
# V is (Nj + nparams) x C in natural units; returns (-logL[C], G[(Nj + nparams), C]).
function batched_fg(MHmodel0::DeviceMH, dispmodel0::GaussianDispersion, V::Matrix{Float64},
                    models::DeviceStack, logAge, MH)
    nl = Vector{Float64}(undef, size(V, 2)); G = similar(V)
    free = UInt8[free_params(MHmodel0)..., free_params(dispmodel0)..., false]
    with_ctx(models) do c
        bind!(models, c, logAge, MH)
        check(ccall((:sfh_eval_fg_hier_batched, libsfh), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Int64, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}),
                    c, mh_kind(MHmodel0), mh_fixed(MHmodel0), Cint(0), V, size(V, 2), free, nl, G))
    end
    return nl, G
end

# The whole stretch-move ensemble sampler on the device: replaces KissMCMC.emcee(MCMCModel(...), x0; ...) inside
# mcmc_sample (mcmc_sample.jl:104).  x0 is ntemplates x nwalkers (nwalkers even); returns samples shaped like
# convert_kissmcmc (:30-44) -- (nsteps / nthin, ntemplates, nwalkers) -- their log-likelihoods and the acceptance fraction.
function device_emcee(models::DeviceStack, x0::Matrix{Float64}, nsteps::Integer; nthin::Integer=1, a_scale::Real=2.0,
                      seed::UInt64=rand(UInt64))
    T, W = size(x0); nstore = nsteps ÷ nthin
    X = copy(x0); chain = Array{Float64,3}(undef, T, W, nstore); lps = Matrix{Float64}(undef, W, nstore)
    lfin = Vector{Float64}(undef, W); acc = Ref{Float64}(0.0)
    with_ctx(models) do c
        check(ccall((:sfh_mcmc_run, libsfh), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Float64, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                    c, X, W, nsteps, nthin, a_scale, seed, chain, lps, lfin, acc))
    end
    return permutedims(chain, (3, 1, 2)), permutedims(lps), acc[]
end

# The template stack built on the device from the per-point arguments partial_cmd_smooth hands to bin_cmd_smooth
# (src/StarFormationHistories.jl:886-888) for every template: no host Hess diagrams, no upload.  `points[t]` is a
# NamedTuple (colors, mags, color_err, mag_err, weights, cov_mult); edges are the ranges calculate_edges returns.
function DeviceStack(edges::Tuple{<:AbstractRange,<:AbstractRange}, points::AbstractVector, data; S::Type=Float64)
    nx, ny = length(edges[1]) - 1, length(edges[2]) - 1
    offs = Int64[0; cumsum(length(p.colors) for p in points)]
    cat(f) = convert(Vector{Float64}, reduce(vcat, (f(p) for p in points)))
    cov = Int32[p.cov_mult for p in points]
    d = convert(Vector{Float64}, vec(data)); h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sfh_stack_create_from_points, libsfh), Cint,
                (Ref{Ptr{Cvoid}}, Int64, Int64, Float64, Float64, Float64, Float64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                h, nx, ny, first(edges[1]), step(edges[1]), first(edges[2]), step(edges[2]), length(points), offs,
                cat(p -> p.colors), cat(p -> p.mags), cat(p -> p.color_err), cat(p -> p.mag_err), cat(p -> p.weights), cov,
                dtype_code(S), d, dtype_code(Float64), C_NULL))
    return DeviceStack{S}(nothing, (nx * ny, length(points)), h[])   # same finalizer / context pool as the uploading constructor
end

# ---- on-disk container (include/sfhcuda.h: sfh_stack_save / sfh_stack_create_from_file) -------------------------
# The reference has no file format (examples/fitting1.ipynb cell 96 uses Serialization by hand).  `save` streams the device
# copy straight into the memory-mapped file; `DeviceStack(path)` uploads it again, reading only the bin rows this process
# holds (`rows = (first, last)` in Julia's 1-based inclusive convention).
function save(path::AbstractString, models::DeviceStack; logAge=nothing, MH=nothing, hess_size::Tuple{Int,Int}=(0, 0))
    la = logAge === nothing ? C_NULL : convert(Vector{Float64}, logAge)
    mh = MH === nothing ? C_NULL : convert(Vector{Float64}, MH)
    check(ccall((:sfh_stack_save, libsfh), Cint, (Ptr{Cvoid}, Cstring, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                models.handle, path, hess_size[1], hess_size[2], la, mh))
end
struct SfhOpts   # sfh_opts, include/sfhcuda.h
    struct_size::Int32; device::Int32; row_begin::Int64; row_end::Int64; clamp_eps::Float64
    tile_bins::Int32; cluster::Int32; force_unfused::Int32; consumer_warps::Int32; variant::Int32; reserved::Int32
end
function DeviceStack(path::AbstractString; S::Type=Float64, rows::Union{Nothing,Tuple{Int,Int}}=nothing, verify::Bool=false, device::Integer=0)
    o = SfhOpts(sizeof(SfhOpts), device, rows === nothing ? 0 : rows[1] - 1, rows === nothing ? 0 : rows[2], 0.0, 0, 0, 0, 0, 0, 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sfh_stack_create_from_file, libsfh), Cint, (Ref{Ptr{Cvoid}}, Cstring, Cint, Ref{SfhOpts}), h, path, verify, o))
    info = zeros(UInt8, 128)   # sfh_info: nbins_total and ntemplates are its first two int64 fields
    check(ccall((:sfh_stack_info, libsfh), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h[], info))
    nb, nt = reinterpret(Int64, info[1:16])
    return DeviceStack{S}(nothing, (Int(nb), Int(nt)), h[])
end

# ---- native BFGS loops: one ccall per optimisation (include/sfhcuda.h: sfh_fit_*_bfgs) ----------------------------
# What fit_templates / fit_templates_fast (solvers.jl:163-275) and fit_sfh (generic_fitting.jl:296-327) hand to
# Optim.optimize(only_fg!(...), x0, BFGS(...)): here the loop runs inside the library around the device evaluations.
struct BfgsOpts; struct_size::Int32; alphaguess::Int32; g_abstol::Float64; maxiter::Int64; device_hessian::Int32; reserved::Int32; end
mutable struct BfgsReport; f::Float64; g_norm::Float64; iterations::Int64; f_calls::Int64; converged::Int32; status::Int32; BfgsReport() = new(); end
bfgs_opts(g_abstol, iterations; device_hessian=false) = BfgsOpts(sizeof(BfgsOpts), 0, g_abstol, iterations, device_hessian, 0)

# transform: 0 = log-space MAP (solvers.jl:178-186), 1 = log-space MLE (:187-195), 2 = sqrt-space MLE (:254-261).
# Returns (minimiser in the fitting space, inverse Hessian, report): fit_templates builds its LogTransformFTResult from them.
function fit_templates_bfgs(models::DeviceStack, theta0::Vector{Float64}, transform::Integer; g_abstol=1e-8, iterations=5000)
    theta = copy(theta0); invH = Matrix{Float64}(undef, length(theta), length(theta)); rep = BfgsReport()
    with_ctx(models) do c
        check(ccall((:sfh_fit_templates_bfgs, libsfh), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ref{BfgsOpts}, Ref{BfgsReport}, Ptr{Float64}),
                    c, transform, theta, bfgs_opts(g_abstol, iterations), rep, invH))
    end
    return theta, invH, rep
end

# fit_sfh's fg_map! (jacobian_corrections = true) / fg_mle! (false) optimisation, generic_fitting.jl:306-327.
# x0 = [log.(R); transformed free parameters] exactly as fit_sfh assembles it (:285-294).
function fit_sfh_bfgs(MHmodel0::DeviceMH, dispmodel0::GaussianDispersion, x0::Vector{Float64},
                      models::DeviceStack, logAge, MH, jacobian_corrections::Bool; g_abstol=1e-8, iterations=5000)
    par = Float64[fittable_params(MHmodel0)..., fittable_params(dispmodel0)...]
    tf = Int32[SFH.transforms(MHmodel0)..., SFH.transforms(dispmodel0)...]
    free = UInt8[free_params(MHmodel0)..., free_params(dispmodel0)...]
    x = copy(x0); invH = Matrix{Float64}(undef, length(x), length(x)); rep = BfgsReport()
    with_ctx(models) do c
        bind!(models, c, logAge, MH)
        check(ccall((:sfh_fit_sfh_bfgs, libsfh), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{UInt8}, Cint, Ptr{Float64}, Ref{BfgsOpts}, Ref{BfgsReport}, Ptr{Float64}),
                    c, mh_kind(MHmodel0), mh_fixed(MHmodel0), Cint(0), par, tf, free, jacobian_corrections, x, bfgs_opts(g_abstol, iterations), rep, invH))
    end
    return x, invH, rep
end

# ---- native multi-chain NUTS: one ccall per sampling run (include/sfhcuda.h: sfh_hmc_sample_nuts / sfh_sample_sfh_nuts) -----------
# Replaces the Threads.@threads loop of DynamicHMC chains in hmc_sample (hmc_sample.jl:123-141) and the task-per-chain loop of
# tsample_sfh (generic_fitting.jl:617-626): the chains run as threads inside the library and share one batched device pass per round.
struct NutsOpts; struct_size::Int32; max_depth::Int32; nwarmup::Int64; delta::Float64; eps0::Float64; seed::UInt64; mass_kind::Int32; reserved::Int32; end

# theta0: ntemplates x nchains (log coefficients).  Returns samples in natural units shaped (nsteps, ntemplates, nchains) like hmc_sample.
function hmc_sample_nuts(models::DeviceStack, theta0::Matrix{Float64}, nsteps::Integer; nwarmup::Integer=200, max_depth::Integer=8,
                         seed::UInt64=rand(UInt64))
    T, nch = size(theta0); lens = fill(Int64(nsteps), nch)
    samples = Matrix{Float64}(undef, T, nsteps * nch); lps = Vector{Float64}(undef, nsteps * nch); steps = Vector{Float64}(undef, nch)
    o = NutsOpts(sizeof(NutsOpts), max_depth, nwarmup, 0.8, 0.0, seed, 0, 0)
    with_ctx(models) do c
        check(ccall((:sfh_hmc_sample_nuts, libsfh), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ref{NutsOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
                    c, nch, theta0, lens, C_NULL, o, samples, lps, steps, C_NULL, C_NULL))
    end
    return permutedims(reshape(exp.(samples), T, nsteps, nch), (2, 1, 3))
end

# x0s: (Nj + nfree) x nchains starting points in the transformed space (tsample_sfh draws them from MvNormal(MLE, MAP.invH), :586);
# invH = MAP.invH is the dense M^-1 of the kinetic energy (:479-482), ϵ the initial step size.  Returns the transformed-space
# samples (columns, chains concatenated) for exptransform_samples! (:640-658).
function sample_sfh_nuts(MHmodel0::DeviceMH, dispmodel0::GaussianDispersion, x0s::Matrix{Float64},
                         lens::Vector{Int64}, invH::Matrix{Float64}, models::DeviceStack, logAge, MH; ϵ::Real=0.05, max_depth::Integer=8,
                         seed::UInt64=rand(UInt64))
    par = Float64[fittable_params(MHmodel0)..., fittable_params(dispmodel0)...]
    tf = Int32[SFH.transforms(MHmodel0)..., SFH.transforms(dispmodel0)...]
    free = UInt8[free_params(MHmodel0)..., free_params(dispmodel0)...]
    n, nch = size(x0s); tot = sum(lens)
    samples = Matrix{Float64}(undef, n, tot); lps = Vector{Float64}(undef, tot); steps = Vector{Float64}(undef, nch)
    o = NutsOpts(sizeof(NutsOpts), max_depth, 0, 0.8, ϵ, seed, 2, 0)
    with_ctx(models) do c
        bind!(models, c, logAge, MH)
        check(ccall((:sfh_sample_sfh_nuts, libsfh), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{UInt8}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64},
                     Ref{NutsOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
                    c, mh_kind(MHmodel0), mh_fixed(MHmodel0), Cint(0), par, tf, free, nch, x0s, lens, invH, o, samples, lps, steps, C_NULL, C_NULL))
    end
    return samples, lps, steps
end

# fit_templates_lbfgsb's LBFGSB.lbfgsb call (solvers.jl:82-90) as one ccall: x0 is the renormalised start (:86); returns (-logL, coeffs).
struct LbfgsbOpts; struct_size::Int32; m::Int32; factr::Float64; pgtol::Float64; maxiter::Int64; maxfun::Int64; end
mutable struct LbfgsbReport; f::Float64; pg_norm::Float64; iterations::Int64; f_calls::Int64; status::Int32; reserved::Int32; LbfgsbReport() = new(); end
function fit_templates_lbfgsb_native(models::DeviceStack, x0::Vector{Float64}; m::Integer=10, factr::Real=1e-12, pgtol::Real=1e-5)
    x = copy(x0); rep = LbfgsbReport()
    with_ctx(models) do c
        check(ccall((:sfh_fit_templates_lbfgsb, libsfh), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{LbfgsbOpts}, Ref{LbfgsbReport}),
                    c, x, LbfgsbOpts(sizeof(LbfgsbOpts), m, factr, pgtol, 0, 0), rep))
    end
    return rep.f, x
end

end # module
