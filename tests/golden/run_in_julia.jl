# run_in_julia.jl -- feeds the committed golden fixtures through the REAL reference stack (Flux Chain, OceanParameterizations operators,
# OrdinaryDiffEq Tsit5, DiffEqSensitivity + Zygote) and reports the distance to the stored vectors.  The build image has no Julia
# (SURVEY.md 8c), so parity is "vs the CPU restatement" until this script has run once on a machine with the reference's environment:
#
#     julia --project=/path/to/NeuralFreeConvection.jl tests/golden/run_in_julia.jl [tests/golden]
#
# It writes tests/golden/julia_report.json; every entry must be below the tolerance printed next to it (north_star: RHS 1e-5,
# trajectories 1e-3, loss and gradient 1e-3) for the oracle -- and with it the CUDA path, which is tested against the oracle -- to count
# as pinned against Julia.  UNTESTED here by construction; it uses only calls that appear in the reference at the cited lines.
using NPZ, JSON, LinearAlgebra, Printf, Statistics
using Flux, OrdinaryDiffEq, DiffEqSensitivity, Zygote, DiffEqFlux
using OceanParameterizations: Dᶜ, Dᶠ

dir = length(ARGS) >= 1 ? ARGS[1] : @__DIR__
relerr(a, b) = norm(Float64.(a) .- Float64.(b)) / norm(Float64.(b))
report = Dict{String,Any}()

# the five Chains of train_free_convection_nde.jl:125-153 with weights taken from a Flux.destructure vector
function chain(arch, Nz)
    arch == "dense-default" && return Chain(Dense(Nz, 4Nz, relu), Dense(4Nz, 4Nz, relu), Dense(4Nz, Nz - 1))
    arch == "dense-wider" && return Chain(Dense(Nz, 8Nz, relu), Dense(8Nz, 8Nz, relu), Dense(8Nz, Nz - 1))
    arch == "dense-deeper" && return Chain(Dense(Nz, 4Nz, relu), Dense(4Nz, 4Nz, relu), Dense(4Nz, 4Nz, relu), Dense(4Nz, Nz - 1))
    k = arch == "conv-2" ? 2 : 4
    return Chain(x -> reshape(x, Nz, 1, 1, 1), Conv((k, 1), 1 => 1, relu), x -> reshape(x, Nz - k + 1),
                 Dense(Nz - k + 1, 4Nz, relu), Dense(4Nz, 4Nz, relu), Dense(4Nz, Nz - 1))
end

# the right-hand side exactly as src/convective_adjustment_nde.jl:17-18,33-48 builds it (K_CA = 0 gives src/free_convection_nde.jl:29-38)
function make_problem(NN, Nz, saveat)
    Δẑ = 1 / Nz
    Dzᶜ, Dzᶠ = Dᶜ(Nz, Δẑ), Dᶠ(Nz, Δẑ)
    _, reconstruct = Flux.destructure(NN)
    function ∂T∂t(T, p, t)
        weights = p[1:end-7]
        bottom_flux, top_flux, σ_T, σ_wT, H, τ, K_CA = p[end-6:end]
        wT = [bottom_flux; reconstruct(weights)(T); top_flux]
        return σ_wT / σ_T * τ / H * (-(Dzᶜ * wT) .+ Dzᶜ * min.(0, K_CA * (Dzᶠ * T)))
    end
    ff = ODEFunction{false}(∂T∂t, tgrad=DiffEqFlux.basic_tgrad)
    return ODEProblem{false}(ff, nothing, (saveat[1], saveat[end]), saveat=saveat), ∂T∂t
end
params(colp, glob, b) = Float32[colp[b, 1], colp[b, 2], glob[1], glob[2], glob[3], glob[4], colp[b, 3]]     # [B][3] = (bottom, top, K_CA)

θ0 = npzread(joinpath(dir, "theta_dense_default_seed0.npy"))

# ---- (1) right-hand side: rhs_<arch>.npz, tolerance 1e-5 ----
for arch in ("dense-default", "dense-wider", "dense-deeper", "conv-2", "conv-4")
    g = npzread(joinpath(dir, "rhs_$arch.npz"))
    NN = chain(arch, 32); _, re = Flux.destructure(NN); NN = re(g["theta"])
    _, f = make_problem(NN, 32, Float32[0, 1])
    e = maximum(relerr(f(g["T"][b, :], [g["theta"]; params(g["colp"], g["glob"], b)], 0f0), g["dTdt"][b, :]) for b in 1:size(g["T"], 1))
    report["rhs_$arch"] = Dict("max_rel_err" => e, "tolerance" => 1e-5)
end

# ---- (2) trajectories: solve(...) of src/solve.jl:4 against the DOP853 truth, tolerance 1e-3 ----
for (tag, scale) in (("1e-5", 1f-5), ("0p3", 0.3f0), ("1p0", 1f0)), K in (0, 2)
    g = npzread(joinpath(dir, "traj_$(tag)_K$K.npz"))
    θ = θ0 .* scale
    NN = chain("dense-default", 32); _, re = Flux.destructure(NN); NN = re(θ)
    prob, _ = make_problem(NN, 32, g["saveat"])
    errs, steps = Float64[], Int[]
    for b in 1:size(g["T0"], 1)
        sol = solve(prob, Tsit5(), reltol=1e-4, u0=g["T0"][b, :], p=[θ; params(g["colp"], g["glob"], b)],
                    sense=InterpolatingAdjoint(autojacvec=ZygoteVJP(), checkpointing=true))
        push!(errs, relerr(Array(sol)', g["T_truth"][b, :, :])); push!(steps, sol.destats.naccept + sol.destats.nreject)
    end
    report["traj_$(tag)_K$K"] = Dict("max_rel_err" => maximum(errs), "tolerance" => 1e-3, "steps_julia" => steps,
                                     "steps_oracle" => g["naccept"] .+ g["nreject"])      # the controller restatement is pinned by the step counts
end

# ---- (3) loss and gradient: nde_loss + Zygote of src/training.jl:45-48,70 (adaptive Tsit5) against the float64 autograd gradient on a
#      fine fixed grid; the golden is a DISCRETE adjoint (kink effects of min(0, K dT/dz): 3e-3 for K = 2, see tests/test_gpu_adjoint.py) ----
for K in (0, 2)
    g = npzread(joinpath(dir, "grad_K$K.npz"))
    NN = chain("dense-default", 32)
    prob, _ = make_problem(NN, 32, g["saveat"])
    B = size(g["T0"], 1)
    truth = cat([g["Ttrue"][b, :, :]' for b in 1:B]..., dims=2)
    function nde_loss(weights)
        sols = cat([solve(prob, Tsit5(), reltol=1e-4, u0=g["T0"][b, :], p=[weights; params(g["colp"], g["glob"], b)],
                          sense=InterpolatingAdjoint(autojacvec=ZygoteVJP(), checkpointing=true)) |> Array for b in 1:B]..., dims=2)
        return Flux.Losses.mse(sols, truth)
    end
    L = nde_loss(θ0)
    dθ = Zygote.gradient(nde_loss, θ0)[1]
    report["grad_K$K"] = Dict("loss_rel_err" => abs(L - g["loss"]) / g["loss"], "grad_rel_err" => relerr(dθ, g["grad"]),
                              "tolerance_loss" => 1e-3, "tolerance_grad" => K == 0 ? 1e-3 : 5e-3)
end

open(joinpath(dir, "julia_report.json"), "w") do io
    JSON.print(io, report, 2)
end
for (k, v) in sort(collect(report), by=first)
    println(rpad(k, 22), v)
end
