# convert_formats.jl -- turns the .npz files written by neuralfreeconvection.jl_b200/formats.py (or by julia/FreeConvectionB200.jl) into
# exactly the JLD2 files the reference writes, so that every figure*.jl / movie*.jl / print_neural_network_runtimes.jl keeps working:
#
#   history_to_jld2(npz, jld2)     neural_network/<e> (Flux.Chain), mean_loss/<e>, runtime/<e>           src/training.jl:3-12, :62
#   network_to_jld2(npz, jld2)     grid_points, neural_network, T_scaling, wT_scaling                     train_free_convection_nde.jl:259-266
#   solutions_to_jld2(npz, jld2)   the above + true / nde / kpp / tke / convective_adjustment / oceananigans (Dict id => NamedTuple)
#                                  + solution_history / solution_loss_history (Dict id => Array)            train_free_convection_nde.jl:376-390
#
# A JLD2 file stores `Flux.Chain` as a serialised Julia struct: only Julia can write it (SURVEY.md row f4), which is why the B200 host
# side writes the same keys as plain arrays and this script does the last step.  Needs NPZ.jl, JLD2.jl, Flux.jl and
# OceanParameterizations (for ZeroMeanUnitVarianceScaling) -- the reference's own environment plus NPZ.
using NPZ, JLD2, Flux
using OceanParameterizations: ZeroMeanUnitVarianceScaling

# arch = [conv_width, in_0, out_0, ..., out_L]: the five Chains of train_free_convection_nde.jl:125-156
function chain_from_arch(arch::AbstractVector{<:Integer})
    cw, dims = arch[1], arch[2:end]
    layers = Any[]
    if cw > 0       # Conv((cw, 1), 1 => 1, relu) between the two reshapes of the reference's conv architectures
        Nz = dims[1] + cw - 1
        push!(layers, x -> reshape(x, Nz, 1, 1, 1), Conv((cw, 1), 1 => 1, relu), x -> reshape(x, Nz - cw + 1))
    end
    for i in 1:length(dims)-1
        push!(layers, Dense(dims[i], dims[i+1], i < length(dims) - 1 ? relu : identity))
    end
    return Chain(layers...)
end

network(theta::AbstractVector, arch) = Flux.destructure(chain_from_arch(Int.(arch)))[2](Float32.(theta))

function scaling(v::AbstractVector)              # (mu, sigma); the struct's fields are μ and σ (src/oceananigans_nn.jl:180-181)
    s = ZeroMeanUnitVarianceScaling(Float64[0, 1])
    return typeof(s)(v[1], v[2])
end

function history_to_jld2(npz_path, jld2_path; arch=nothing)
    f = npzread(npz_path)
    arch = something(arch, get(f, "arch", nothing))
    arch === nothing && error("the history has no `arch` entry: pass arch = [conv_width, in_0, out_0, ..., out_L]")
    epochs = sort([parse(Int, split(k, "/")[2]) for k in keys(f) if startswith(k, "neural_network/")])
    jldopen(jld2_path, "w") do file
        for e in epochs
            file["neural_network/$e"] = network(f["neural_network/$e"], arch)
            file["mean_loss/$e"] = f["mean_loss/$e"][]
            file["runtime/$e"] = f["runtime/$e"][]
        end
    end
    return length(epochs)
end

function write_header(file, f)
    file["grid_points"] = Int(f["grid_points"][])
    file["neural_network"] = network(f["neural_network/theta"], f["neural_network/arch"])
    file["T_scaling"] = scaling(f["T_scaling"])
    file["wT_scaling"] = scaling(f["wT_scaling"])
end

network_to_jld2(npz_path, jld2_path) = jldopen(file -> write_header(file, npzread(npz_path)), jld2_path, "w")

function solutions_to_jld2(npz_path, jld2_path)
    f = npzread(npz_path)
    groups = ("true", "nde", "kpp", "tke", "convective_adjustment", "oceananigans")
    jldopen(jld2_path, "w") do file
        write_header(file, f)
        for g in groups
            ids = unique(parse(Int, split(k, "/")[2]) for k in keys(f) if startswith(k, g * "/"))
            isempty(ids) && continue
            file[g] = Dict(id => (; (Symbol(split(k, "/")[3]) => f[k] for k in keys(f) if startswith(k, "$g/$id/"))...) for id in ids)
        end
        for g in ("solution_history", "solution_loss_history")
            ks = [k for k in keys(f) if startswith(k, g * "/")]
            isempty(ks) || (file[g] = Dict(parse(Int, split(k, "/")[2]) => f[k] for k in ks))
        end
    end
end

if abspath(PROGRAM_FILE) == @__FILE__
    length(ARGS) == 3 || error("usage: julia convert_formats.jl history|network|solutions IN.npz OUT.jld2")
    kind, src, dst = ARGS
    kind == "history" ? history_to_jld2(src, dst) : kind == "network" ? network_to_jld2(src, dst) : solutions_to_jld2(src, dst)
end
