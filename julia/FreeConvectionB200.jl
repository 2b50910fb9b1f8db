# FreeConvectionB200.jl -- thin Julia host for libnfc_b200.so (include/nfc.h).
#
# Drop-in for the numerical work behind src/free_convection_nde.jl, src/convective_adjustment_nde.jl, src/solve.jl and
# src/training.jl of NeuralFreeConvection.jl: the reference's exported names and signatures are kept, only the bodies
# call the C ABI.  This file cannot be exercised in the build image (no Julia); it is the binding a maintainer adds.
module FreeConvectionB200

using Flux
using JLD2: jldopen
using Oceananigans: interior, znodes              # the reference's datasets are FieldTimeSeries (src/data.jl)
using Printf

export solve_nde_history, FreeConvectionNDE, ConvectiveAdjustmentNDE, solve_nde, nde_loss_and_gradient, diagnose_flux,
       oceananigans_convective_adjustment_with_neural_network,
       FreeConvectionNDEParameters, ConvectiveAdjustmentNDEParameters, train_neural_differential_equation!, inscribe_history,
       Tsit5, ROCK4, RKC2, nfc_comm_unique_id, attach_communicator!

const LIB = get(ENV, "LIBNFC_B200", joinpath(@__DIR__, "..", "neuralfreeconvection.jl_b200", "libnfc_b200.so"))
const NFC_MAX_LAYERS = 8

# mirrors `struct nfc_config` field for field (include/nfc.h)
struct NfcConfig
    struct_size::Int32; Nz::Int32; conv_width::Int32; n_layers::Int32
    layer_dims::NTuple{9,Int32}; activations::NTuple{8,Int32}
    nde_type::Int32; alg::Int32; reltol::Float32; abstol::Float32; fixed_dt::Float32
    max_iters::Int32; n_save::Int32; saveat::Ptr{Float32}; device::Int32
    adjoint_max_steps::Int32; adjoint_reltol::Float32; adjoint_abstol::Float32; reserved::NTuple{8,Int32}
end

mutable struct NDEHandle          # what FreeConvectionNDE(NN, ds) returns instead of an ODEProblem
    ptr::Ptr{Cvoid}
    nde_type::Int32
    saveat::Vector{Float32}
    Nz::Int
    alg::Int32                    # NFC_ALG_TSIT5 = 0, NFC_ALG_RKC2 = 1 (fixed per C handle)
    make::Function                # alg -> a fresh handle with the same Chain / save times (the reference passes alg per solve call)
    others::Dict{Int32,Any}       # handles of the other algorithms, created on first use
end

# Time steppers.  The host does not need OrdinaryDiffEq any more; these stand in for its algorithm objects so that
# `algorithm = Meta.parse(args["time-stepper"] * "()") |> eval` (train_free_convection_nde.jl:92) keeps working.
struct Tsit5 end
struct RKC2 end          # stabilised explicit Runge-Kutta-Chebyshev, in the role ROCK4 plays in the reference's production runs
const ROCK4 = RKC2       # train_free_convection_nde.sh:3: ROCK4's tables are OrdinaryDiffEq's; the library runs RKC2 (include/nfc.h)
alg_id(::Tsit5) = Int32(0)
alg_id(::RKC2) = Int32(1)
alg_id(x) = error("time stepper $(typeof(x)): libnfc_b200 implements Tsit5() and the stabilised explicit RKC2() / ROCK4()")

"The handle that runs `alg` (one C handle per algorithm, sharing the Chain layout and the save times)."
function handle_for(nde::NDEHandle, alg)
    a = alg_id(alg)
    a == nde.alg && return nde
    return get!(() -> nde.make(a), nde.others, a)
end

check(h, rc, what) = rc == 0 || error("$what failed ($rc): " * unsafe_string(ccall((:nfc_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

function chain_layout(NN::Chain, Nz)
    conv = 0; dims = Int32[]; acts = Int32[]
    for l in NN.layers
        if l isa Conv
            conv = size(l.weight, 1)
        elseif l isa Dense
            isempty(dims) && push!(dims, size(l.weight, 2))
            push!(dims, size(l.weight, 1)); push!(acts, l.σ === relu ? 1 : 0)
        end
    end
    return conv, dims, acts
end

function make_nde(nde_type, NN, ds; iterations=nothing, device=0, alg=Int32(0), adjoint_max_steps=0)
    T = ds["T"]; Nz = size(T, 3); Nt = size(T, 4)
    iterations = isnothing(iterations) ? (1:Nt) : iterations
    saveat = collect(Float32, range(0f0, Float32(maximum(iterations) / Nt), length=length(iterations)))   # src/free_convection_nde.jl:40-41
    conv, dims, acts = chain_layout(NN, Nz)
    pad(v, n) = ntuple(i -> i <= length(v) ? Int32(v[i]) : Int32(0), n)
    function create(a)
        cfg = Ref(NfcConfig(sizeof(NfcConfig), Nz, conv, length(acts), pad(dims, 9), pad(acts, 8), nde_type, a, 1f-4, 1f-6, 0f0,
                            100_000, length(saveat), pointer(saveat), device, adjoint_max_steps, 0f0, 0f0, ntuple(_ -> Int32(0), 8)))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve saveat begin
            rc = ccall((:nfc_create, LIB), Int32, (Ref{NfcConfig}, Ref{Ptr{Cvoid}}), cfg, out)
        end
        check(C_NULL, rc, "nfc_create")
        h = NDEHandle(out[], nde_type, saveat, Nz, a, create, Dict{Int32,Any}())
        finalizer(x -> ccall((:nfc_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
    return create(Int32(alg))
end

FreeConvectionNDE(NN, ds; kw...) = make_nde(Int32(0), NN, ds; kw...)            # src/free_convection_nde.jl:1
ConvectiveAdjustmentNDE(NN, ds; kw...) = make_nde(Int32(1), NN, ds; kw...)      # src/convective_adjustment_nde.jl:1

# nde_params: 6- or 7-vector (or 6|7 x B matrix) exactly as FreeConvectionNDEParameters / ConvectiveAdjustmentNDEParameters build it
function split_params(nde, p::AbstractVecOrMat)
    P = p isa AbstractVector ? reshape(p, :, 1) : p
    want = nde.nde_type == 0 ? 6 : 7
    size(P, 1) == want || throw(ArgumentError("need $want trailing parameters, got $(size(P, 1))"))
    colp = zeros(Float32, 3, size(P, 2)); colp[1:2, :] .= P[1:2, :]
    want == 7 && (colp[3, :] .= P[7, :])
    return colp, Float32.(P[3:6, 1])
end

"solve_nde(nde, NN, T₀, alg, nde_params) |> Array  (src/solve.jl:1-6).  T₀: Nz or Nz x B; returns Nz x n_save (x B)."
function solve_nde(nde::NDEHandle, NN, T₀, alg, nde_params)
    nde = handle_for(nde, alg)
    θ, _ = Flux.destructure(NN)
    check(nde.ptr, ccall((:nfc_set_weights, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), nde.ptr, θ, length(θ)), "nfc_set_weights")
    T0 = Float32.(T₀ isa AbstractVector ? reshape(T₀, :, 1) : T₀); B = size(T0, 2)
    colp, glob = split_params(nde, nde_params)
    out = Array{Float32}(undef, nde.Nz, length(nde.saveat), B)
    na = zeros(Int32, B); nr = zeros(Int32, B); st = zeros(Int32, B)
    rc = ccall((:nfc_solve, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int64),
               nde.ptr, T0, colp, glob, out, na, nr, st, B)
    check(nde.ptr, rc, "nfc_solve")
    any(!=(0), st) && @warn "solver status != 0 for $(count(!=(0), st)) columns: their trajectories are NaN after the failure"   # reference: retcode warnings, never exceptions
    return T₀ isa AbstractVector ? out[:, :, 1] : out
end

"""
    solve_nde(ds, NN, NDEType, nde_params, algorithm, T_scaling, wT_scaling; T₀=nothing) -> (T=..., wT=...)

src/solve.jl:8-51: solution and diagnosed flux of one simulation in physical units (`Nz x Nt` and `(Nz+1) x Nt`, Float64 like the
reference's `zeros(Nz+1, Nt)`).  The reference reads `nde_params[7]` whatever the NDE type (src/solve.jl:30), so `nde_params` must be the
7-vector of ConvectiveAdjustmentNDEParameters; a FreeConvectionNDE gets its first six entries.
"""
function solve_nde(ds, NN, NDEType, nde_params, algorithm, T_scaling, wT_scaling; T₀=nothing)
    nde = handle_for(NDEType(NN, ds), algorithm)
    isnothing(T₀) && (T₀ = T_scaling.(interior(ds["T"])[1, 1, :, 1]))
    K_CA = nde_params[7]                                        # throws for a 6-vector exactly like the reference
    p = nde.nde_type == 0 ? nde_params[1:6] : nde_params
    T = solve_nde(nde, NN, Float32.(T₀), algorithm, Float32.(p))               # Nz x Nt, scaled units
    wT = diagnose_flux(nde, NN, reshape(T, size(T, 1), size(T, 2), 1), Float32.(p))[:, :, 1]
    return (T=inv(T_scaling).(Float64.(T)), wT=inv(wT_scaling).(Float64.(wT)))
end

"nfc_set_option: kernel selection and scratch budgets (the names are listed in include/nfc.h)."
set_option(nde::NDEHandle, name::AbstractString, value::Integer) =
    check(nde.ptr, ccall((:nfc_set_option, LIB), Int32, (Ptr{Cvoid}, Cstring, Int64), nde.ptr, name, value), "nfc_set_option")

"Loss and dL/dθ of src/training.jl:45-48,70 for all simulations in one call; Ttrue is Nz x n_save x B (already T_scaling'ed)."
function nde_loss_and_gradient(nde::NDEHandle, θ::Vector{Float32}, T₀::Matrix{Float32}, nde_params, Ttrue::Array{Float32,3})
    check(nde.ptr, ccall((:nfc_set_weights, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), nde.ptr, θ, length(θ)), "nfc_set_weights")
    colp, glob = split_params(nde, nde_params); B = size(T₀, 2)
    loss = Ref{Float64}(0); grad = similar(θ); st = zeros(Int32, B)
    call() = check(nde.ptr, ccall((:nfc_loss_grad, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ref{Float64}, Ptr{Float32}, Ptr{Int32}, Int64),
                                  nde.ptr, T₀, colp, glob, Ttrue, loss, grad, st, B), "nfc_loss_grad")
    call()
    # a failed column is excluded from the gradient while its residuals stay in the loss: never train on that silently.  Status 4 = a column
    # accepted more steps than the dense adjoint tape holds (adjoint_max_steps rows per column): the compact tape has no such cap
    if any(==(4), st)
        set_option(nde, "tape_mode", 2)
        call()
        any(==(4), st) && error("adjoint tape overflow for $(count(==(4), st)) columns (more than 32 765 accepted steps)")
    end
    any(!=(0), st) && @warn "nfc_loss_grad: $(count(!=(0), st)) of $B columns failed (status $(unique(st[st .!= 0]))) and are missing from the gradient"
    return loss[], grad
end

# ---- the trailing-parameter vectors (src/free_convection_nde.jl:49-62, src/convective_adjustment_nde.jl:59-72) ----
function FreeConvectionNDEParameters(ds, T_scaling, wT_scaling)
    wT = ds["wT"]
    H, τ = abs(znodes(wT)[1]), wT.times[end]                   # domain depth, simulation length
    fluxes = (wT_scaling(interior(wT)[1, 1, 1, 1]), wT_scaling(ds.metadata["temperature_flux"]))      # bottom, top
    return eltype(wT).([fluxes..., T_scaling.σ, wT_scaling.σ, H, τ])
end

ConvectiveAdjustmentNDEParameters(ds, T_scaling, wT_scaling, K_CA) =
    vcat(FreeConvectionNDEParameters(ds, T_scaling, wT_scaling), eltype(ds["wT"])(K_CA))

# ---- training (src/training.jl:3-12, 34-78) ----
function inscribe_history(history_filepath, epoch; kwargs...)
    jldopen(history_filepath, "a") do file
        for (name, val) in pairs(kwargs)
            file["$name/$epoch"] = val
        end
    end
    return nothing
end
inscribe_history(::Nothing, args...; kwargs...) = nothing

"""
    train_neural_differential_equation!(NN, NDEType, nde_params, algorithm, datasets, T_scaling, iterations, opt, epochs; history_filepath=nothing)

src/training.jl:34-78 with the same arguments and the same return value (the Chain rebuilt from the best weights seen,
GalacticOptim's `sol.minimizer`).  One epoch = ONE nfc_loss_grad call over all training simulations (the reference solves them one after
the other and differentiates with Zygote + DiffEqSensitivity) + one `Flux.Optimise.update!` with `opt` (`ADAM()` in
train_free_convection_nde.jl:245).  The loss recorded for epoch e is the loss of the weights the gradient was taken at -- the reference's
callback re-solves all simulations for the same number (src/training.jl:53-54); `inscribe_history` gets the same three entries.
Training runs on Tsit5 or, with `ROCK4()` / `RKC2()`, on the stabilised explicit integrator.
"""
function train_neural_differential_equation!(NN, NDEType, nde_params, algorithm, datasets, T_scaling, iterations, opt, epochs; history_filepath=nothing, device=0)
    ids = sort(collect(keys(datasets)))
    T₀ = Float32.(hcat([T_scaling.(interior(datasets[id]["T"])[1, 1, :, 1]) for id in ids]...))                    # Nz x B
    true_sols = Float32.(cat([T_scaling.(interior(datasets[id]["T"])[1, 1, :, iterations]) for id in ids]..., dims=3))   # Nz x n_save x B
    P = Float32.(hcat([nde_params[id] for id in ids]...))                                                      # 6|7 x B
    nde = handle_for(NDEType(NN, datasets[first(ids)]; iterations, device), algorithm)                        # one handle serves every simulation
    weights, reconstruct = Flux.destructure(NN)
    weights = Float32.(weights)
    best_loss, best_weights = Inf, copy(weights)
    t₀ = time_ns()
    for epoch in 1:epochs
        mse_loss, grad = nde_loss_and_gradient(nde, weights, T₀, P, true_sols)
        @info @sprintf("Training free convection NDE epoch %d/%d... MSE loss: μ_loss::%s = %.12e", epoch, epochs, typeof(mse_loss), mse_loss)
        mse_loss < best_loss && ((best_loss, best_weights) = (mse_loss, copy(weights)))
        runtime = (time_ns() - t₀) * 1e-9
        inscribe_history(history_filepath, epoch, neural_network=reconstruct(weights), mean_loss=mse_loss; runtime)
        t₀ = time_ns()
        Flux.Optimise.update!(opt, weights, grad)
    end
    return reconstruct(best_weights)
end

# ---- multi-GPU (include/nfc.h: nfc_comm_*): one Julia process per GPU; the host ships the 128-byte id (e.g. MPI.Bcast!) ----
function nfc_comm_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:nfc_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id)
    rc == 0 || error("nfc_comm_unique_id failed ($rc): " * unsafe_string(ccall((:nfc_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    return id
end

"After this (collective) call nde_loss_and_gradient returns the loss and gradient of ALL ranks' simulations on every rank."
function attach_communicator!(nde::NDEHandle, id::Vector{UInt8}, rank::Integer, world::Integer)
    check(nde.ptr, ccall((:nfc_comm_init, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), nde.ptr, id, rank, world), "nfc_comm_init")
    return nde
end

"""
The epoch x simulation loop of compute_nde_solution_history (src/testing.jl:62-73) as one call.  `thetas` is n_params x epochs
(column e = Flux.destructure(history["neural_network/\$e"])[1]), T₀ is Nz x S (one column per dataset, T_scaling'ed),
nde_params 7 x S.  Returns Nz x n_save x S x epochs (scaled units; apply inv(T_scaling) as src/solve.jl:50 does).
"""
function solve_nde_history(nde::NDEHandle, thetas::Matrix{Float32}, T₀::Matrix{Float32}, nde_params)
    E = size(thetas, 2); S = size(T₀, 2)
    colp, glob = split_params(nde, nde_params)
    T0e = repeat(T₀, 1, E); colpe = repeat(colp, 1, E)                     # member e owns columns (e-1)S+1 : eS
    out = Array{Float32}(undef, nde.Nz, length(nde.saveat), S, E)
    na = zeros(Int32, S, E); nr = zeros(Int32, S, E); st = zeros(Int32, S, E)
    rc = ccall((:nfc_solve_ensemble, LIB), Int32,
               (Ptr{Cvoid}, Ptr{Float32}, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int64),
               nde.ptr, thetas, E, T0e, colpe, glob, out, na, nr, st, S)
    check(nde.ptr, rc, "nfc_solve_ensemble")
    any(!=(0), st) && @warn "solver status != 0 for $(count(!=(0), st)) (simulation, epoch) pairs"
    return out
end

"""
The flux diagnosis loop of solve_nde(ds, ...) (src/solve.jl:32-48) for a solution `T` (Nz x n_save x B, scaled units):
wT[:, n] = [bottom; NN(T[:, n]); top] - min(0, K_CA dT/dz), returned as (Nz+1) x n_save x B (scaled units; apply inv(wT_scaling)).
"""
function diagnose_flux(nde::NDEHandle, NN, T::Array{Float32,3}, nde_params)
    θ, _ = Flux.destructure(NN)
    check(nde.ptr, ccall((:nfc_set_weights, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), nde.ptr, θ, length(θ)), "nfc_set_weights")
    colp, _ = split_params(nde, nde_params); B = size(T, 3)
    wT = Array{Float32}(undef, nde.Nz + 1, size(T, 2), B)
    rc = ccall((:nfc_diagnose_flux, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Int64), nde.ptr, T, colp, wT, B)
    check(nde.ptr, rc, "nfc_diagnose_flux")
    return wT
end

"""
oceananigans_convective_adjustment_with_neural_network (src/oceananigans_nn.jl:158-282) for B columns at once: explicit neural-network flux
step + implicit convective adjustment per fixed step `Δt`.  `T₀` is Nz x B (physical temperatures), `heat_flux` the B surface temperature
fluxes; `K` is non-dimensionalised as the reference does (:195).  Returns the reference's `(T, wT)`: Nz x n_out x B and (Nz+1) x n_out x B
(n_out = n_steps ÷ save_every + 1) in physical units.
"""
function oceananigans_convective_adjustment_with_neural_network(nde::NDEHandle, NN, T₀::Matrix{Float32}, heat_flux::Vector{Float32},
                                                                T_scaling, wT_scaling; Lz, stop_time, Δt=600.0, K=10.0, ∂T∂z_deep=0.0, save_every=1)
    θ, _ = Flux.destructure(NN)
    check(nde.ptr, ccall((:nfc_set_weights, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), nde.ptr, θ, length(θ)), "nfc_set_weights")
    B = size(T₀, 2); n_steps = round(Int, stop_time / Δt)
    Knd = wT_scaling.σ / T_scaling.σ * stop_time / Lz * K
    scal = Float32[T_scaling.μ, T_scaling.σ, wT_scaling.μ, wT_scaling.σ]
    out = Array{Float32}(undef, nde.Nz, n_steps ÷ save_every + 1, B)
    rc = ccall((:nfc_embedded_solve, LIB), Int32,
               (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Float32, Float32, Float32, Float32, Int32, Int32, Ptr{Float32}, Int64),
               nde.ptr, T₀, heat_flux, scal, Lz, Δt, Knd, ∂T∂z_deep, n_steps, save_every, out, B)
    check(nde.ptr, rc, "nfc_embedded_solve")
    # second output of the reference, diagnose_wT_NN (src/oceananigans_nn.jl:200-218): (Nz+1) x n_out x B
    wT = Array{Float32}(undef, nde.Nz + 1, size(out, 2), B)
    rc = ccall((:nfc_embedded_diagnose_flux, LIB), Int32,
               (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Float32, Float32, Int32, Ptr{Float32}, Int64),
               nde.ptr, out, heat_flux, scal, Lz, Knd, size(out, 2), wT, B)
    check(nde.ptr, rc, "nfc_embedded_diagnose_flux")
    return (T=out, wT=wT)
end

end # module
