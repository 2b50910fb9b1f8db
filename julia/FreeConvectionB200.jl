# FreeConvectionB200.jl -- thin Julia host for libnfc_b200.so (include/nfc.h).
#
# Drop-in for the numerical work behind src/free_convection_nde.jl, src/convective_adjustment_nde.jl, src/solve.jl and
# src/training.jl of NeuralFreeConvection.jl: the reference's exported names and signatures are kept, only the bodies
# call the C ABI.  This file cannot be exercised in the build image (no Julia); it is the binding a maintainer adds.
module FreeConvectionB200

using Flux
export solve_nde_history, FreeConvectionNDE, ConvectiveAdjustmentNDE, solve_nde, nde_loss_and_gradient, diagnose_flux,
       oceananigans_convective_adjustment_with_neural_network

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

function make_nde(nde_type, NN, ds; iterations=nothing, device=0)
    T = ds["T"]; Nz = size(T, 3); Nt = size(T, 4)
    iterations = isnothing(iterations) ? (1:Nt) : iterations
    saveat = collect(Float32, range(0f0, Float32(maximum(iterations) / Nt), length=length(iterations)))   # src/free_convection_nde.jl:40-41
    conv, dims, acts = chain_layout(NN, Nz)
    pad(v, n) = ntuple(i -> i <= length(v) ? Int32(v[i]) : Int32(0), n)
    cfg = Ref(NfcConfig(sizeof(NfcConfig), Nz, conv, length(acts), pad(dims, 9), pad(acts, 8), nde_type, 0, 1f-4, 1f-6, 0f0,
                        100_000, length(saveat), pointer(saveat), device, 0, 0f0, 0f0, ntuple(_ -> Int32(0), 8)))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve saveat begin
        rc = ccall((:nfc_create, LIB), Int32, (Ref{NfcConfig}, Ref{Ptr{Cvoid}}), cfg, out)
    end
    check(C_NULL, rc, "nfc_create")
    h = NDEHandle(out[], nde_type, saveat, Nz)
    finalizer(x -> ccall((:nfc_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
    return h
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
    θ, _ = Flux.destructure(NN)
    check(nde.ptr, ccall((:nfc_set_weights, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), nde.ptr, θ, length(θ)), "nfc_set_weights")
    T0 = Float32.(T₀ isa AbstractVector ? reshape(T₀, :, 1) : T₀); B = size(T0, 2)
    colp, glob = split_params(nde, nde_params)
    out = Array{Float32}(undef, nde.Nz, length(nde.saveat), B)
    na = zeros(Int32, B); nr = zeros(Int32, B); st = zeros(Int32, B)
    rc = ccall((:nfc_solve, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int64),
               nde.ptr, T0, colp, glob, out, na, nr, st, B)
    check(nde.ptr, rc, "nfc_solve")
    any(!=(0), st) && @warn "solver status != 0 for $(count(!=(0), st)) columns"      # reference: retcode warnings, never exceptions
    return T₀ isa AbstractVector ? out[:, :, 1] : out
end

"Loss and dL/dθ of src/training.jl:45-48,70 for all simulations in one call; Ttrue is Nz x n_save x B (already T_scaling'ed)."
function nde_loss_and_gradient(nde::NDEHandle, θ::Vector{Float32}, T₀::Matrix{Float32}, nde_params, Ttrue::Array{Float32,3})
    check(nde.ptr, ccall((:nfc_set_weights, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64), nde.ptr, θ, length(θ)), "nfc_set_weights")
    colp, glob = split_params(nde, nde_params); B = size(T₀, 2)
    loss = Ref{Float64}(0); grad = similar(θ); st = zeros(Int32, B)
    rc = ccall((:nfc_loss_grad, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ref{Float64}, Ptr{Float32}, Ptr{Int32}, Int64),
               nde.ptr, T₀, colp, glob, Ttrue, loss, grad, st, B)
    check(nde.ptr, rc, "nfc_loss_grad")
    return loss[], grad
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
fluxes; `K` is non-dimensionalised as the reference does (:195).  Returns Nz x (n_steps ÷ save_every + 1) x B in physical units.
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
    return out
end

end # module
