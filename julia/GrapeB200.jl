# GrapeB200.jl -- `ccall` shim between GRAPE.jl and libgrape_b200.so (include/grape_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The same C
# symbols are exercised through ctypes (grape.jl_b200/_lib.py, tests/), and the
# struct below is checked field-by-field against the header by
# tests/test_abi_and_host.py::test_julia_shim_matches_header.
#
# What it replaces in the reference (paths relative to the GRAPE.jl tree):
#   GrapeWrk(trajectories, tlist, kwargs)     src/workspace.jl:147-362  -> B200Workspace(...)
#   evaluate_functional(pulsevals, wrk)       src/optimize.jl:696-768   -> evaluate_functional(pulsevals, b)
#   evaluate_gradient!(G, pulsevals, wrk)     src/optimize.jl:824-1014  -> evaluate_gradient!(G, pulsevals, b)
# Everything else (optimizer loop, result bookkeeping, callbacks, printing)
# stays the reference's own Julia code; see INTEGRATION.md for the 12-line patch
# that routes `prop_method = GrapeB200.B200ExpProp` through this module.
module GrapeB200

using LinearAlgebra

export B200ExpProp, B200Workspace, B200MultiWorkspace, evaluate_functional, evaluate_gradient!,
    evaluate_gradient_host_chi!, evaluate_functional_amplitudes, evaluate_gradient_amplitudes!

"Selector passed as `prop_method` (reference src/docstring.jl:177, 201-225)."
struct B200ExpProp end

const LIB = get(ENV, "GRAPE_B200_LIB", "libgrape_b200.so")
const ABI_VERSION = Int32(1)

# error codes (include/grape_b200.h)
const OK, EINVAL, ECUDA, ECHINORM, ETAYLOR, ENOCONTROLS, ESTATE, ENCCL = 0:7
# functional / method / running-cost kinds
const JT_SM, JT_RE, JT_SS, JT_HOST = Int32(0), Int32(1), Int32(2), Int32(3)
const GRADGEN, TAYLOR = Int32(0), Int32(1)
const JA_NONE, JA_FLUENCE = Int32(0), Int32(1)
const GB_NONE, GB_QUADFORM = Int32(0), Int32(1)

"Mirror of `struct grape_b200_problem` (include/grape_b200.h:83-113); field order and types must match."
struct Problem
    abi_version::Int32
    K::Int32
    N::Int32
    L::Int32
    NT::Int32
    G::Int32
    K_global::Int32
    device::Int32
    tlist::Ptr{Float64}
    gen_of_traj::Ptr{Int32}
    H0::Ptr{Float64}
    Hc::Ptr{Float64}
    shape::Ptr{Float64}
    psi0::Ptr{Float64}
    tgt::Ptr{Float64}
    weights::Ptr{Float64}
    functional::Int32
    gradient_method::Int32
    ja_kind::Int32
    gb_kind::Int32
    lambda_a::Float64
    lambda_b::Float64
    gb_D::Ptr{Float64}
    gb_nD::Int32
    taylor_max_order::Int32
    taylor_tolerance::Float64
    taylor_check_convergence::Int32
    path::Int32
    chi_min_norm::Float64
end

"""
Device-resident replacement of the hot fields of `GrapeWrk`
(src/workspace.jl:78-144).  The small host vectors that callbacks and
`make_grape_print_iters` read (src/optimize.jl:402-478) are kept here and are
refreshed by every call.
"""
mutable struct B200Workspace
    handle::Ptr{Cvoid}
    K::Int
    N::Int
    L::Int
    NT::Int
    J_parts::Vector{Float64}          # wrk.J_parts
    tau_vals::Vector{ComplexF64}      # wrk.result.tau_vals
    grad_J_Tb::Vector{Float64}        # wrk.grad_J_Tb
    grad_J_a::Vector{Float64}         # wrk.grad_J_a
    functional::Int32
    lambda_a::Float64                 # lambda_a / lambda_b and the running-cost kinds of this workspace: the host-chi path
    lambda_b::Float64                 # assembles J_parts and G itself (src/optimize.jl:755-766, 1002-1011)
    ja_fluence::Bool
    gb_on::Bool
    tlist::Vector{Float64}
end

function _last_error(h::Ptr{Cvoid})
    p = ccall((:grape_b200_last_error, LIB), Cstring, (Ptr{Cvoid},), h)
    return p == C_NULL ? "" : unsafe_string(p)
end

# Library failure -> Julia exception, so that the try/catch of
# src/optimize.jl:125-135 (result.message = "Exception: ...") keeps working.
function _check(rc::Integer, h::Ptr{Cvoid})
    rc == OK && return nothing
    error(_last_error(h))
end

"""
    B200Workspace(tlist, H0, Hc, psi0, tgt; kwargs...)

`H0::Vector{Matrix{ComplexF64}}` (one drift per distinct generator),
`Hc::Vector{Vector{Matrix{ComplexF64}}}` (`Hc[g][l]`), `psi0`/`tgt` vectors of
state vectors, `gen_of_traj` 1-based generator index per trajectory.  Julia
matrices are column-major, which is the ABI's layout: no transposition.
"""
function B200Workspace(tlist::Vector{Float64}, H0, Hc, psi0, tgt;
        gen_of_traj = nothing, shape = nothing, weights = nothing,
        functional = JT_SM, gradient_method = :gradgen,
        J_a_fluence::Bool = false, lambda_a = 1.0,
        g_b_D = nothing, lambda_b = 1.0, chi_min_norm = 1e-100,
        taylor_grad_max_order = 100, taylor_grad_tolerance = 1e-16,
        taylor_grad_check_convergence = true, device = 0, K_global = 0)
    K = length(psi0); N = length(psi0[1]); G = length(H0); L = length(Hc[1])
    L == 0 && error("no controls in trajectories: cannot optimize")   # src/workspace.jl:155-157
    NT = length(tlist) - 1
    h0 = ComplexF64[x for g in 1:G for x in vec(Matrix{ComplexF64}(H0[g]))]
    hc = ComplexF64[x for g in 1:G for l in 1:L for x in vec(Matrix{ComplexF64}(Hc[g][l]))]
    p0 = ComplexF64[x for k in 1:K for x in psi0[k]]
    tg = ComplexF64[x for k in 1:K for x in tgt[k]]
    gen = isnothing(gen_of_traj) ? (G == 1 ? zeros(Int32, K) : Int32.(0:K-1)) : Int32.(gen_of_traj .- 1)
    shp = isnothing(shape) ? Float64[] : Float64[shape[l][n] for l in 1:L for n in 1:NT]
    w = isnothing(weights) ? Float64[] : Float64.(weights)
    Ds = isnothing(g_b_D) ? ComplexF64[] :
         (g_b_D isa AbstractMatrix ? ComplexF64[x for x in vec(Matrix{ComplexF64}(g_b_D))] :
          ComplexF64[x for D in g_b_D for x in vec(Matrix{ComplexF64}(D))])
    nD = isnothing(g_b_D) ? 0 : (g_b_D isa AbstractMatrix ? 1 : length(g_b_D))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = GC.@preserve tlist h0 hc p0 tg gen shp w Ds begin
        prob = Problem(ABI_VERSION, K, N, L, NT, G, K_global, device,
            pointer(tlist), pointer(gen),
            Ptr{Float64}(pointer(h0)), Ptr{Float64}(pointer(hc)),
            isempty(shp) ? Ptr{Float64}(C_NULL) : pointer(shp),
            Ptr{Float64}(pointer(p0)), Ptr{Float64}(pointer(tg)),
            isempty(w) ? Ptr{Float64}(C_NULL) : pointer(w),
            Int32(functional), gradient_method == :taylor ? TAYLOR : GRADGEN,
            J_a_fluence ? JA_FLUENCE : JA_NONE, nD > 0 ? GB_QUADFORM : GB_NONE,
            Float64(lambda_a), Float64(lambda_b),
            nD > 0 ? Ptr{Float64}(pointer(Ds)) : Ptr{Float64}(C_NULL), Int32(nD),
            Int32(taylor_grad_max_order), Float64(taylor_grad_tolerance),
            Int32(taylor_grad_check_convergence), Int32(0), Float64(chi_min_norm))
        ccall((:grape_b200_create, LIB), Cint, (Ref{Problem}, Ref{Ptr{Cvoid}}), prob, out)
    end
    rc == OK || error(_last_error(Ptr{Cvoid}(C_NULL)))
    b = B200Workspace(out[], K, N, L, NT, zeros(3), zeros(ComplexF64, K),
                      zeros(L * NT), zeros(L * NT), Int32(functional),
                      Float64(lambda_a), Float64(lambda_b), J_a_fluence, nD > 0, copy(tlist))
    finalizer(b) do x
        x.handle == C_NULL || ccall((:grape_b200_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle)
        x.handle = C_NULL
    end
    return b
end

"`evaluate_functional(pulsevals, wrk)` (src/optimize.jl:696-768): returns `sum(J_parts)`."
function evaluate_functional(pulsevals::Vector{Float64}, b::B200Workspace)
    rc = GC.@preserve pulsevals b ccall((:grape_b200_eval_f, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{ComplexF64}),
        b.handle, pulsevals, b.J_parts, b.tau_vals)
    _check(rc, b.handle)
    return sum(b.J_parts)
end

"`evaluate_gradient!(G, pulsevals, wrk)` (src/optimize.jl:824-1014): fills `G` in place, returns `J`."
function evaluate_gradient!(G::Vector{Float64}, pulsevals::Vector{Float64}, b::B200Workspace)
    rc = GC.@preserve G pulsevals b ccall((:grape_b200_eval_fg, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Float64}),
        b.handle, pulsevals, G, b.J_parts, b.tau_vals, b.grad_J_Tb, b.grad_J_a)
    _check(rc, b.handle)
    return sum(b.J_parts)
end

"""
Arbitrary `J_T` / `chi` closures (functional = JT_HOST): forward sweep on the
device, `J_T`/`chi` evaluated in Julia on the K final states, backward sweep
on the device (src/optimize.jl:845-855 stays Julia code).
"""
function evaluate_gradient_host_chi!(G, pulsevals, b::B200Workspace, J_T, chi, trajectories)
    sums = zeros(4)
    rc = GC.@preserve pulsevals ccall((:grape_b200_forward, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{ComplexF64}, Ptr{Float64}), b.handle, pulsevals, b.tau_vals, sums)
    _check(rc, b.handle)
    Ψ = final_states(b)
    b.J_parts[1] = J_T(Ψ, trajectories; tau = b.tau_vals)                   # src/optimize.jl:755-760
    χ = chi(Ψ, trajectories; tau = b.tau_vals)                              # src/optimize.jl:845-855
    chiT = ComplexF64[x for k in 1:b.K for x in χ[k]]
    jb = Ref{Float64}(0.0)
    rc = GC.@preserve chiT G ccall((:grape_b200_backward_chi, LIB), Cint,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
        b.handle, chiT, b.grad_J_Tb, jb, b.grad_J_a)
    _check(rc, b.handle)
    copyto!(G, b.grad_J_Tb)                                                 # src/optimize.jl:1003
    # running costs: the library returns sum_k J_b_trajectory[k] and the fluence gradient 2 eps dt; the lambdas are
    # applied here exactly as the reference does (src/optimize.jl:761-766, 1004-1011)
    b.J_parts[2] = 0.0
    if b.ja_fluence
        dt = diff(b.tlist)
        J_a = 0.0
        for l in 1:b.L, n in 1:b.NT
            J_a += pulsevals[(l - 1) * b.NT + n]^2 * dt[n]
        end
        b.J_parts[2] = b.lambda_a * J_a
        axpy!(b.lambda_a, b.grad_J_a, G)
    end
    b.J_parts[3] = b.gb_on ? b.lambda_b * jb[] : 0.0
    return sum(b.J_parts)
end

"""
Amplitude mode (include/grape_b200.h): non-linear controls and per-term amplitudes.  The workspace's `L` slots are
(control, amplitude) pairs; the caller evaluates `ampl[i, n] = a_i(ϵ, t_n)` and `dampl[i, n] = ∂a_i/∂ϵ` from its
closures (reference `get_control_derivs`, src/workspace.jl:283-285; `evaluate(μ)`, src/optimize.jl:946-951, with
`dampl = 0` for the `isnothing(μ)` branch) and adds the returned slot gradients per control.
"""
function evaluate_functional_amplitudes(ampl::Vector{Float64}, b::B200Workspace)
    rc = GC.@preserve ampl b ccall((:grape_b200_eval_f_amplitudes, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{ComplexF64}), b.handle, ampl, b.J_parts, b.tau_vals)
    _check(rc, b.handle)
    return sum(b.J_parts)
end

function evaluate_gradient_amplitudes!(G_slots::Vector{Float64}, ampl::Vector{Float64}, dampl::Vector{Float64},
                                       b::B200Workspace)
    rc = GC.@preserve G_slots ampl dampl b ccall((:grape_b200_eval_fg_amplitudes, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{ComplexF64}),
        b.handle, ampl, dampl, G_slots, b.J_parts, b.tau_vals)
    _check(rc, b.handle)
    return sum(b.J_parts)
end

"`wrk.fw_propagators[k].state` for all k (read by `update_result!`, src/optimize.jl:187-189)."
function final_states(b::B200Workspace)
    out = zeros(ComplexF64, b.N, b.K)
    _check(ccall((:grape_b200_get_final_states, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), b.handle, out), b.handle)
    return [out[:, k] for k in 1:b.K]
end

"`wrk.fw_storage[k]` as an `N × (NT+1)` matrix (src/workspace.jl:215); fetched lazily."
function stored_states(b::B200Workspace, k::Integer)
    out = zeros(ComplexF64, b.N, b.NT + 1)
    _check(ccall((:grape_b200_get_stored_states, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{ComplexF64}),
                 b.handle, Int32(k - 1), out), b.handle)
    return out
end

"`wrk.chi_states`, `wrk.chi_states_norm` (src/optimize.jl:867-869)."
function chi_states(b::B200Workspace)
    chi = zeros(ComplexF64, b.N, b.K); rho = zeros(b.K)
    _check(ccall((:grape_b200_get_chi_states, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}),
                 b.handle, chi, rho), b.handle)
    return [chi[:, k] for k in 1:b.K], rho
end

"""
    dense_orders(b) -> (economised::Bool, orders::Vector{Int32})

Polynomial degree of every time step in the last call of the dense path (N > 32). `economised` is true when the
chains summed the Chebyshev-cut polynomial of `exp(-i H dt)` -- what `prop_method = Cheby` is in the reference
(docs/src/tutorial.md:308, 432) -- which the library selects by itself for Hermitian generators.
"""
function dense_orders(b::B200Workspace)
    orders = zeros(Int32, b.NT)
    rc = ccall((:grape_b200_dense_orders, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), b.handle, orders)
    rc < 0 && _check(-rc, b.handle)
    return rc == 1, orders
end

# ------------------------------------------------------------------------------------------------------------
# Several GPUs from this one Julia process (the reference threads over trajectories, src/optimize.jl:720, 876):
# `grape_b200_multi_create` splits the trajectories into contiguous blocks, one per device; the shards reduce the
# tau sums and the gradient among themselves over NVLink peer memory (csrc/xchg.cuh) -- nothing here but two ccalls.
# ------------------------------------------------------------------------------------------------------------
mutable struct B200MultiWorkspace
    handle::Ptr{Cvoid}
    K::Int
    N::Int
    L::Int
    NT::Int
    J_parts::Vector{Float64}
    tau_vals::Vector{ComplexF64}
    grad_J_Tb::Vector{Float64}
    grad_J_a::Vector{Float64}
end

function _multi_error(h::Ptr{Cvoid})
    p = ccall((:grape_b200_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), h)
    return p == C_NULL ? "" : unsafe_string(p)
end

"""
    B200MultiWorkspace(devices, tlist, H0, Hc, psi0, tgt; kwargs...)

Same arguments as `B200Workspace` for the WHOLE ensemble plus the list of CUDA device ordinals.
"""
function B200MultiWorkspace(devices::Vector{<:Integer}, tlist::Vector{Float64}, H0, Hc, psi0, tgt;
        gen_of_traj = nothing, shape = nothing, weights = nothing, functional = JT_SM, gradient_method = :gradgen,
        J_a_fluence::Bool = false, lambda_a = 1.0, g_b_D = nothing, lambda_b = 1.0, chi_min_norm = 1e-100,
        taylor_grad_max_order = 100, taylor_grad_tolerance = 1e-16, taylor_grad_check_convergence = true)
    K = length(psi0); N = length(psi0[1]); G = length(H0); L = length(Hc[1])
    L == 0 && error("no controls in trajectories: cannot optimize")
    NT = length(tlist) - 1
    h0 = ComplexF64[x for g in 1:G for x in vec(Matrix{ComplexF64}(H0[g]))]
    hc = ComplexF64[x for g in 1:G for l in 1:L for x in vec(Matrix{ComplexF64}(Hc[g][l]))]
    p0 = ComplexF64[x for k in 1:K for x in psi0[k]]
    tg = ComplexF64[x for k in 1:K for x in tgt[k]]
    gen = isnothing(gen_of_traj) ? (G == 1 ? zeros(Int32, K) : Int32.(0:K-1)) : Int32.(gen_of_traj .- 1)
    shp = isnothing(shape) ? Float64[] : Float64[shape[l][n] for l in 1:L for n in 1:NT]
    w = isnothing(weights) ? Float64[] : Float64.(weights)
    Ds = isnothing(g_b_D) ? ComplexF64[] :
         (g_b_D isa AbstractMatrix ? ComplexF64[x for x in vec(Matrix{ComplexF64}(g_b_D))] :
          ComplexF64[x for D in g_b_D for x in vec(Matrix{ComplexF64}(D))])
    nD = isnothing(g_b_D) ? 0 : (g_b_D isa AbstractMatrix ? 1 : length(g_b_D))
    devs = Int32.(devices)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = GC.@preserve tlist h0 hc p0 tg gen shp w Ds devs begin
        prob = Problem(ABI_VERSION, K, N, L, NT, G, Int32(0), Int32(0),
            pointer(tlist), pointer(gen),
            Ptr{Float64}(pointer(h0)), Ptr{Float64}(pointer(hc)),
            isempty(shp) ? Ptr{Float64}(C_NULL) : pointer(shp),
            Ptr{Float64}(pointer(p0)), Ptr{Float64}(pointer(tg)),
            isempty(w) ? Ptr{Float64}(C_NULL) : pointer(w),
            Int32(functional), gradient_method == :taylor ? TAYLOR : GRADGEN,
            J_a_fluence ? JA_FLUENCE : JA_NONE, nD > 0 ? GB_QUADFORM : GB_NONE,
            Float64(lambda_a), Float64(lambda_b),
            nD > 0 ? Ptr{Float64}(pointer(Ds)) : Ptr{Float64}(C_NULL), Int32(nD),
            Int32(taylor_grad_max_order), Float64(taylor_grad_tolerance),
            Int32(taylor_grad_check_convergence), Int32(0), Float64(chi_min_norm))
        ccall((:grape_b200_multi_create, LIB), Cint, (Ref{Problem}, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}),
              prob, devs, Int32(length(devs)), out)
    end
    rc == OK || error(_multi_error(Ptr{Cvoid}(C_NULL)))
    m = B200MultiWorkspace(out[], K, N, L, NT, zeros(3), zeros(ComplexF64, K), zeros(L * NT), zeros(L * NT))
    finalizer(m) do x
        x.handle == C_NULL || ccall((:grape_b200_multi_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle)
        x.handle = C_NULL
    end
    return m
end

function evaluate_functional(pulsevals::Vector{Float64}, m::B200MultiWorkspace)
    rc = GC.@preserve pulsevals m ccall((:grape_b200_multi_eval_f, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{ComplexF64}), m.handle, pulsevals, m.J_parts, m.tau_vals)
    rc == OK || error(_multi_error(m.handle))
    return sum(m.J_parts)
end

function evaluate_gradient!(G::Vector{Float64}, pulsevals::Vector{Float64}, m::B200MultiWorkspace)
    rc = GC.@preserve G pulsevals m ccall((:grape_b200_multi_eval_fg, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Float64}),
        m.handle, pulsevals, G, m.J_parts, m.tau_vals, m.grad_J_Tb, m.grad_J_a)
    rc == OK || error(_multi_error(m.handle))
    return sum(m.J_parts)
end

function final_states(m::B200MultiWorkspace)
    out = zeros(ComplexF64, m.N, m.K)
    rc = ccall((:grape_b200_multi_get_final_states, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), m.handle, out)
    rc == OK || error(_multi_error(m.handle))
    return [out[:, k] for k in 1:m.K]
end

end # module
