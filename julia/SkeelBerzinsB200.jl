# SkeelBerzinsB200.jl -- Julia host shim over libskeelberzins_b200.so (C ABI: include/skeelberzins_b200.h).
#
# Drop-in for the internal implicit-Euler / stationary-Newton path of SkeelBerzins.jl, batched over B independent
# PDE instances.  The Julia side owns the time loop (reference src/pdepe.jl:90-108) and the Newton loop
# (src/utils.jl:311-336) exactly as the reference does and calls one C entry point per Newton iteration; all the
# arithmetic runs in hand-written sm_100a CUDA kernels.  No CUDA.jl, no CPU fallback.
#
# STATUS: EXPERIMENTAL, NEVER EXECUTED.  `julia` is not installed in the build image or on the GPU box, so nothing in
# this file has run.  What is checked (tests/test_c_abi.py): every ccall names an exported symbol and passes the number
# and kind of arguments the header declares.  What is NOT checked: array layouts as Julia sees them (column-major views
# of the C arrays), finalizers calling CUDA from the GC thread, keyword merging.  The Python host
# (skeelberzins.jl_b200/api.py) drives the identical sequence of C calls and is what the tests and the bench run.
#
# pdefun / bdfun must be registry functors (device-compiled restatements of the example callbacks, see
# `Functor`): Julia closures cannot run on the GPU.  icfun stays an ordinary Julia function evaluated on the host,
# as in the reference (src/utils.jl:629-632).
module SkeelBerzinsB200

export pdepe, Params, ProblemDefinition, SkeelBerzinsDiscretization, Functor, assemble!, mass_matrix, DiffEqArray,
       interpolate_sol_time, interpolate_sol_space, pdeval, ode_operator

const libsb = get(ENV, "SKEELBERZINS_B200_LIB", "libskeelberzins_b200.so")

# registry ids (enum sb_functor_id)
const LIN_DIFF, NONLIN_DIFF, LIN_DIFF_CYL, POISSON, STAT_NONLIN, REACT_DIFF = 0, 1, 2, 3, 4, 5
const LIN_DIFF_SPH, CONV_DIFF, LIN_REACT, GENERIC_SCALAR = 6, 7, 8, 9
const DESIGN_SPLIT, DESIGN_FUSED, DESIGN_FUSED_TMA, DESIGN_RESIDENT = 0, 1, 2, 3

struct SBError <: Exception
    code::Cint
    msg::String
end

last_error() = unsafe_string(ccall((:sb_last_error, libsb), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : throw(SBError(rc, last_error()))

"""Registry coefficient functor (pdefun + bdfun) with per-instance parameters, `params[nparams, B]`."""
struct Functor
    id::Int
    params::Matrix{Float64}     # column k = parameters of instance k  (C side: instance-major, contiguous)
end
Functor(id::Integer, p::AbstractVector) = Functor(Int(id), reshape(collect(Float64, p), :, 1))
batch(f::Functor) = size(f.params, 2)

function functor_info(id::Integer)
    npde = Ref{Cint}(0); nprm = Ref{Cint}(0); name = Ref{Cstring}(C_NULL)
    check(ccall((:sb_functor_info, libsb), Cint, (Cint, Ref{Cint}, Ref{Cint}, Ref{Cstring}), id, npde, nprm, name))
    Int(npde[]), Int(nprm[]), unsafe_string(name[])
end

"""Same fields as SkeelBerzins.Params (src/utils.jl:191-238)."""
Base.@kwdef struct Params
    solver::Symbol = :euler
    tstep::Union{Float64, Vector{Float64}} = 1e-2
    hist::Bool = false
    sparsity::Symbol = :sparseArrays      # accepted; the Jacobian lives on the device in block-tridiagonal form
    linsolve::Any = nothing               # accepted; the GPU solver is fixed (partitioned block-Thomas)
    maxit::Int = 100
    tol::Float64 = 1e-10
    data::Bool = false
    nb_design_var::Int = 0
end

mutable struct Context
    h::Ptr{Cvoid}
    device::Int
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:sb_ctx_create, libsb), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, r))
        c = new(r[], device)
        finalizer(x -> ccall((:sb_ctx_destroy, libsb), Cint, (Ptr{Cvoid},), x.h), c)
        c
    end
end

"""Problem record, same field names as SkeelBerzins.ProblemDefinition (src/utils.jl:113-182); `handle` is the
sb_problem*; the Jacobian (`jac` in the reference) is device scratch rewritten every Newton iteration."""
mutable struct ProblemDefinition
    npde::Int
    Nx::Int
    xmesh::Vector{Float64}
    tspan::Tuple{Float64, Float64}
    singular::Bool
    m::Int
    nb_design_var::Int
    inival::Matrix{Float64}      # [npde*Nx, B], component fastest (src/utils.jl:632)
    ξ::Vector{Float64}
    ζ::Vector{Float64}
    pdefunction::Functor
    icfunction::Any
    bdfunction::Functor
    ctx::Context
    handle::Ptr{Cvoid}
end
const SkeelBerzinsDiscretization = ProblemDefinition
Base.length(pb::ProblemDefinition) = pb.nb_design_var          # src/utils.jl:665

function problem_init(m, xmesh, tspan, pdefun::Functor, icfun, bdfun, params; ctx = Context(0), design = DESIGN_RESIDENT)
    x = collect(Float64, xmesh)
    Nx = length(x); B = batch(pdefun)
    npde, nprm, _ = functor_info(pdefun.id)
    size(pdefun.params, 1) == nprm || error("functor takes $nprm parameters per instance")
    params.sparsity in (:sparseArrays, :banded) ||
        throw("Error: Invalid sparsity pattern selected. Please choose from the available options: :sparseArrays, :banded")
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sb_problem_create, libsb), Cint,
                (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ref{Ptr{Cvoid}}),
                ctx.h, m, Nx, B, x, pdefun.id, pdefun.params, nprm, h))
    check(ccall((:sb_problem_set_design, libsb), Cint, (Ptr{Cvoid}, Cint), h[], design))
    u0 = npde == 1 ? icfun.(x) : vec(reduce(hcat, icfun.(x)))   # src/utils.jl:632
    inival = repeat(reshape(collect(Float64, u0), :, 1), 1, B)
    ξ = zeros(Nx - 1); ζ = zeros(Nx - 1)
    check(ccall((:sb_problem_quadrature, libsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], ξ, ζ))
    sing = Ref{Cint}(0)
    check(ccall((:sb_problem_info, libsb), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Int64}, Ref{Cint}),
                h[], C_NULL, C_NULL, C_NULL, sing))
    pb = ProblemDefinition(npde, Nx, x, (Float64(tspan[1]), Float64(tspan[2])), sing[] != 0, m, params.nb_design_var,
                           inival, ξ, ζ, pdefun, icfun, pdefun, ctx, h[])
    finalizer(p -> ccall((:sb_problem_destroy, libsb), Cint, (Ptr{Cvoid},), p.handle), pb)
    Nx, npde, inival, pb
end

upload!(pb, u::AbstractArray{Float64}) = check(ccall((:sb_state_upload, libsb), Cint, (Ptr{Cvoid}, Ptr{Float64}), pb.handle, u))
function download(pb)
    u = Matrix{Float64}(undef, pb.npde * pb.Nx, batch(pb.pdefunction))
    check(ccall((:sb_state_download, libsb), Cint, (Ptr{Cvoid}, Ptr{Float64}), pb.handle, u))
    u
end

"""mass_matrix(pb) (src/assembler.jl:138-174): diagonal per instance and the DAE flag."""
function mass_matrix(pb::ProblemDefinition)
    upload!(pb, pb.inival)
    check(ccall((:sb_mass_flags, libsb), Cint, (Ptr{Cvoid}, Cdouble, Cint), pb.handle, pb.tspan[1], 0))
    M = similar(pb.inival)
    check(ccall((:sb_mass_download, libsb), Cint, (Ptr{Cvoid}, Ptr{Float64}), pb.handle, M))
    M, any(==(0.0), M)
end

"""assemble!(du, u, pb, t) (src/assembler.jl:19-125), on `[npde*Nx, B]` arrays."""
function assemble!(du, u, pb::ProblemDefinition, t)
    check(ccall((:sb_assemble, libsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cdouble, Ptr{Float64}), pb.handle, u, t, du))
    du
end

"""newton (src/utils.jl:303-337) on the device-resident state: b = M .* u_n (src/pdepe.jl:94), unP1 = copy(b),
then the Newton loop, host-owned; `depth` iterations stay enqueued behind the one being waited for."""
function newton(pb::ProblemDefinition, tau, timeStep; tol = 1.0e-10, maxit = 100, depth = 2)
    check(ccall((:sb_begin_step, libsb), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), pb.handle, timeStep, isinf(tau) ? 0.0 : 1 / tau))
    launched = 0; done = 0
    nact = Ref{Int64}(0)
    while true
        while launched < maxit && launched - done < depth
            check(ccall((:sb_newton_iteration, libsb), Cint, (Ptr{Cvoid}, Cdouble), pb.handle, tol))
            launched += 1
        end
        check(ccall((:sb_wait_iteration, libsb), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}), pb.handle, done, nact))
        done += 1
        nact[] == 0 && return launched
        done >= maxit && throw("convergence failed")            # src/utils.jl:336
    end
end

function history(pb, maxit)
    B = batch(pb.pdefunction)
    h = Matrix{Float64}(undef, maxit, B)
    check(ccall((:sb_get_history, libsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), pb.handle, h, maxit))
    it = Vector{Int32}(undef, B)
    check(ccall((:sb_get_step_iters, libsb), Cint, (Ptr{Cvoid}, Ptr{Int32}), pb.handle, it))
    [h[1:it[k], k] for k in 1:B]
end

"""pdeval(m, xmesh, u_approx, x_eval, pb) (src/utils.jl:396-438) on the GPU: `(u, dudx)` as `[npde, nq, B]` arrays."""
function pdeval(m, xmesh, u_approx, x_eval, pb::ProblemDefinition)
    m == pb.m || error("The symmetry argument is different from the one used to solve the PDE problem.")
    xq = collect(Float64, x_eval isa Number ? [x_eval] : x_eval)
    B = batch(pb.pdefunction)
    u = Array{Float64}(undef, pb.npde, length(xq), B); du = similar(u)
    check(ccall((:sb_pdeval, libsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                pb.handle, u_approx, length(xq), xq, u, du))
    u, du
end

"""Plugin slot: CUDA C++ source of a functor struct (see include/skeelberzins_b200.h) -> functor id for `Functor`."""
function register_user_functor(cuda_source::AbstractString, struct_name::AbstractString, npde::Integer, nparams::Integer)
    id = Ref{Cint}(0)
    check(ccall((:sb_register_user_functor, libsb), Cint, (Ptr{Cvoid}, Cstring, Cstring, Cint, Cint, Ref{Cint}),
                C_NULL, cuda_source, struct_name, npde, nparams, id))
    Int(id[])
end

"""Result of the transient solve with the two fields the reference's `RecursiveArrayTools.DiffEqArray` result is used
through (src/pdepe.jl:110): `sol.u[k]` = states `[npde*Nx, B]` at `sol.t[k]`; `sol(t)` interpolates in time
(src/utils.jl:597), `sol(x_eval, t, pb)` in time and space (src/utils.jl:561)."""
struct DiffEqArray
    u::Vector{Matrix{Float64}}
    t::Vector{Float64}
end
Base.length(s::DiffEqArray) = length(s.u)
Base.getindex(s::DiffEqArray, i) = s.u[i]
Base.lastindex(s::DiffEqArray) = lastindex(s.u)
(s::DiffEqArray)(t::Real) = interpolate_sol_time(s, t)
(s::DiffEqArray)(x_eval, t, pb::ProblemDefinition; val = true, deriv = true) = interpolate_sol_space(s, x_eval, t, pb; val = val, deriv = deriv)

"""interpolate_sol_time(u_approx, t) (src/utils.jl:575-595): linear interpolation between stored time steps."""
function interpolate_sol_time(u_approx::DiffEqArray, t)
    ts = u_approx.t
    (t == ts[1] || abs(t - ts[1]) ≤ 1.0e-10 * abs(ts[2] - ts[1])) && return u_approx.u[1]
    idx = searchsortedfirst(ts, t)
    (idx == 1 || idx > length(u_approx)) && error("The chosen time is not within the interval.")
    dt = ts[idx] - ts[idx - 1]
    ((ts[idx] - t) / dt) .* u_approx.u[idx - 1] .+ ((t - ts[idx - 1]) / dt) .* u_approx.u[idx]
end

"""interpolate_sol_space(u_approx, x_eval, times, pb; val, deriv) (src/utils.jl:460-559), evaluated on the GPU."""
function interpolate_sol_space(u_approx::DiffEqArray, x_eval, times, pb::ProblemDefinition; val = true, deriv = true)
    res = [pdeval(pb.m, pb.xmesh, interpolate_sol_time(u_approx, t), x_eval, pb) for t in (times isa Number ? [times] : times)]
    us = [r[1] for r in res]; dus = [r[2] for r in res]
    times isa Number && ((us, dus) = (us[1], dus[1]))
    val && !deriv && return us
    deriv && !val && return dus
    us, dus
end

"""assemble! on DEVICE memory (sb_assemble_device): `d_u`, `d_du` are device pointers (e.g. `pointer(::CuArray)` reinterpreted
to `Ptr{Float64}`) to `[npde*Nx, B]` arrays; asynchronous on the context's stream."""
function assemble_device!(d_du::Ptr{Float64}, d_u::Ptr{Float64}, pb::ProblemDefinition, t)
    check(ccall((:sb_assemble_device, libsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cdouble, Ptr{Float64}), pb.handle, d_u, t, d_du))
    d_du
end

"""DiffEq-facing operator (src/diffEq.jl:18-43): `(f!, M, tspan)` with `f!(du, u, p, t) = assemble!(du, u, pb, t)` on the
GPU and `M` the diagonal of the mass matrix (`nothing` for a pure ODE system); pass them on as
`ODEFunction(f!; mass_matrix = Diagonal(vec(M)))` / `ODEProblem(..., vec(pb.inival), tspan)`."""
function ode_operator(pb::ProblemDefinition)
    M, flag_dae = mass_matrix(pb)
    f!(du, u, p, t) = assemble!(du, u, pb, t)
    f!, (flag_dae ? M : nothing), pb.tspan
end

_merge(params, kwargs) = Params(; (f => get(kwargs, f, getfield(params, f)) for f in fieldnames(Params))...)

"""Transient pdepe (src/pdepe.jl:42-128) for the batch.  Returns a `DiffEqArray` (`sol.u[k]` = `[npde*Nx, B]` states at
`sol.t[k]`, `sol(t)`, `sol(x_eval, t, pb)`) and, as in the reference, `pb` / `storage` when `data` / `hist` are set.
`c_loop = true` hands the time loop and the Newton loop to C (one kernel launch for the whole solve); with
`save = :last` only u(t_end) is kept (`sb_solve_steps_final`: the download travels while later instances compute)."""
function pdepe(m, pdefun::Functor, icfun, bdfun, xmesh, tspan; params = Params(), ctx = Context(0), c_loop = false, save = :all, kwargs...)
    params = _merge(params, kwargs)
    @assert m == 0 || m == 1 || m == 2 "Parameter m invalid"
    @assert m == 0 || m > 0 && xmesh[1] ≥ 0 "Non-conforming mesh"
    @assert params.tstep ≠ Inf "Adjust time step or try the alternative method for solving elliptic problems"
    Nx, npde, inival, pb = problem_init(m, xmesh, tspan, pdefun, icfun, bdfun, params; ctx = ctx)
    params.solver == :DiffEq && return pb
    params.hist && check(ccall((:sb_enable_history, libsb), Cint, (Ptr{Cvoid}, Cint), pb.handle, params.maxit))
    flag_tstep = params.tstep isa Real
    timeSteps, tstep = flag_tstep ? (collect(tspan[1]:(params.tstep):tspan[2]), params.tstep) :
                                    (params.tstep, diff(params.tstep))          # src/pdepe.jl:82-85
    upload!(pb, inival)
    check(ccall((:sb_mass_flags, libsb), Cint, (Ptr{Cvoid}, Cdouble, Cint), pb.handle, timeSteps[1], 0))
    if c_loop && !params.hist
        # the time loop and the Newton loop in C: with the resident design the whole solve is one kernel launch
        # (sb_solve_steps_opt) that also writes every state, streamed out while later instances compute
        nsteps = length(timeSteps) - 1
        done = Ref{Cint}(0)
        if save == :last
            final = Array{Float64, 2}(undef, npde * Nx, batch(pdefun))                   # C layout [B][Nx][npde]
            rc = ccall((:sb_solve_steps_final, libsb), Cint,
                       (Ptr{Cvoid}, Cint, Ptr{Float64}, Cdouble, Cdouble, Cint, Ptr{Float64}, Ref{Cint}),
                       pb.handle, nsteps, collect(Float64, timeSteps), flag_tstep ? Float64(tstep) : -1.0, params.tol,
                       params.maxit, final, done)
            rc == -4 && throw("convergence failed")                                   # src/utils.jl:336
            check(rc)
            sol = DiffEqArray([final], [Float64(timeSteps[end])])
            return params.data ? (sol, pb) : sol
        end
        snaps = Array{Float64, 3}(undef, npde * Nx, batch(pdefun), nsteps + 1)      # C layout [nsteps+1][B][Nx][npde]
        rc = ccall((:sb_solve_steps_opt, libsb), Cint,
                   (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Cdouble, Cdouble, Cint, Ptr{Int32}, Ptr{Float64}, Ref{Cint}),
                   pb.handle, nsteps, collect(Float64, timeSteps), C_NULL, flag_tstep ? Float64(tstep) : -1.0, params.tol,
                   params.maxit, C_NULL, snaps, done)
        rc == -4 && throw("convergence failed")                                       # src/utils.jl:336
        check(rc)
        sol = DiffEqArray([snaps[:, :, k] for k in 1:(nsteps + 1)], collect(Float64, timeSteps))
        return params.data ? (sol, pb) : sol
    end
    results = [copy(inival)]
    storage = []
    for i in eachindex(timeSteps[1:(end - 1)])                                     # src/pdepe.jl:90
        newton(pb, flag_tstep ? tstep : tstep[i], timeSteps[i + 1]; tol = params.tol, maxit = params.maxit)
        params.hist && push!(storage, history(pb, params.maxit))
        push!(results, download(pb))
    end
    sol = DiffEqArray(results, collect(Float64, timeSteps))                        # src/pdepe.jl:110
    params.hist && params.data && return sol, pb, storage
    params.hist && return sol, storage
    params.data && return sol, pb
    sol
end

"""Stationary pdepe (src/pdepe.jl:156-192): one Newton solve with tau = Inf at t = 1."""
function pdepe(m, pdefun::Functor, icfun, bdfun, xmesh; params = Params(; tstep = Inf), ctx = Context(0), kwargs...)
    params = _merge(params, merge(Dict{Symbol, Any}(:tstep => Inf), Dict{Symbol, Any}(kwargs)))
    @assert params.solver == :euler "Elliptic problems can only be solved using the Euler method"
    @assert params.tstep == Inf "Time step should be set to Inf to obtain the stationary solution"
    Nx, npde, inival, pb = problem_init(m, xmesh, (0, 1), pdefun, icfun, bdfun, params; ctx = ctx)
    params.hist && check(ccall((:sb_enable_history, libsb), Cint, (Ptr{Cvoid}, Cint), pb.handle, params.maxit))
    upload!(pb, inival)
    check(ccall((:sb_mass_flags, libsb), Cint, (Ptr{Cvoid}, Cdouble, Cint), pb.handle, 1.0, 1))
    newton(pb, Inf, 1.0; tol = params.tol, maxit = params.maxit)                   # src/pdepe.jl:180
    u = download(pb)
    hist = params.hist ? history(pb, params.maxit) : nothing
    params.hist && params.data && return u, pb, hist
    params.hist && return u, hist
    params.data && return u, pb
    u
end

end # module
