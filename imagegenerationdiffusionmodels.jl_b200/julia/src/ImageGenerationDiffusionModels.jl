# Drop-in package module: the public API of ImageGenerationDiffusionModels.jl (README.md:16-30,47) with
# the DDPM hot path routed through libddpm.so.  Host-side concerns (MAT/BSON/PNG I/O, RNG draws,
# logging, early stopping) stay in Julia exactly as in the reference; Flux is only used as the
# *container* for the BSON model format (`Main.SimpleUNet` of Flux layers), never for arithmetic.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI (no Julia in the image); the Python mirror api.py is.
module ImageGenerationDiffusionModels

using MAT, Images, FileIO, Random
using Flux                      # layer structs only: BSON files store Flux.Conv / BatchNorm / Chain objects
using BSON: @save, @load

include("LibDDPM.jl")
using .LibDDPM

export generate_grid, apply_noise, train, denoise_image, generate_image, demo

# ---- constants of the scripts (src/train_brain.jl:17-24; T = 500 is the intended value, see SURVEY.md trap 2)
const D = 128
const T = 500
const β_min = Float32(1e-4)
const β_max = Float32(0.02)
const β     = collect(range(β_min, β_max, length=T))
const α     = 1 .- β
const α_cum = accumulate(*, α)

function timestep_embedding(t::Integer; D::Int=D)          # src/train_brain.jl:54-62, unchanged
    pe = zeros(Float32, D)
    for i in 1:(D ÷ 2)
        div = exp(-log(Float32(1e4)) * (2*(i-1)/(D-1)))
        pe[2*i-1] = sin(t * div)
        pe[2*i  ] = cos(t * div)
    end
    return pe
end
const PE = hcat((timestep_embedding(t) for t in 1:T)...)   # D x T, column t

# ---- the model container: identical struct and constructor to src/train_brain.jl:89-145.
# BSON.jl tags a struct with the module path of its type.  The shipped checkpoints were written by the driver SCRIPTS,
# so their model type is `Main.SimpleUNet` (SURVEY.md Appendix C): `@load` resolves that name in Main, and a struct
# defined inside this package would be SAVED as `ImageGenerationDiffusionModels.SimpleUNet` -- readable by nobody else.
# To keep the on-disk format unchanged in BOTH directions the struct is therefore defined in Main when the package
# loads (unless the user's script already did, as the reference's scripts do) and referred to through `unet_type()`.
function __init__()
    if !isdefined(Main, :SimpleUNet)
        Core.eval(Main, quote
            import Flux
            struct SimpleUNet
                down1::Flux.Chain
                down2::Flux.Chain
                mid::Flux.Chain
                up2::Flux.Chain
                up1::Flux.Chain
                final::Flux.Conv
            end
        end)
    end
end
unet_type() = getfield(Main, :SimpleUNet)

function SimpleUNet(channels::Int=1)                       # src/train_brain.jl:109-145
    down1 = Chain(Conv((3,3), channels + D => 64, pad=1), BatchNorm(64, relu), Conv((3,3), 64 => 64, pad=1), BatchNorm(64, relu))
    down2 = Chain(MaxPool((2,2)), Conv((3,3), 64 => 128, pad=1), BatchNorm(128, relu), Conv((3,3), 128 => 128, pad=1), BatchNorm(128, relu))
    mid   = Chain(Conv((3,3), 128 => 128, pad=1), BatchNorm(128, relu), Conv((3,3), 128 => 128, pad=1), BatchNorm(128, relu))
    up2   = Chain(ConvTranspose((2,2), 128 => 64, stride=2), Conv((3,3), 64 => 64, pad=1), BatchNorm(64, relu), Conv((3,3), 64 => 64, pad=1), BatchNorm(64, relu))
    up1   = Chain(Conv((3,3), 128 => 64, pad=1), BatchNorm(64, relu), Conv((3,3), 64 => 64, pad=1), BatchNorm(64, relu))
    final = Conv((1,1), 64 => 1)
    # the type was created by `__init__` at run time: construct it in the latest world
    return Base.invokelatest(unet_type(), down1, down2, mid, up2, up1, final)
end

"The 64 Float32 arrays in BSON order: conv (weight, bias); BatchNorm (β, γ, μ, σ²).  They alias the model."
function flux_arrays(m)
    out = Array{Float32}[]
    for chain in (m.down1, m.down2, m.mid, m.up2, m.up1)
        for l in chain.layers
            if l isa Conv || l isa ConvTranspose
                push!(out, l.weight, l.bias)
            elseif l isa BatchNorm
                push!(out, l.β, l.γ, l.μ, l.σ²)
            end
        end
    end
    push!(out, m.final.weight, m.final.bias)
    return out
end

const _engine = Ref{Union{Nothing,LibDDPM.Handle}}(nothing)
function engine()
    if _engine[] === nothing
        h = LibDDPM.Handle(T=T, D=D)
        LibDDPM.set_tables!(h, β, α_cum, PE)       # host-computed => bit-exact with this file by construction
        _engine[] = h
    end
    return _engine[]
end

const _model = Ref{Any}(nothing)
function default_model()
    if _model[] === nothing
        @load joinpath(@__DIR__, "..", "..", "..", "fixtures", "trained_model.bson") model
        _model[] = model
    end
    return _model[]
end

# ---- unchanged host helpers -------------------------------------------------------------------
function generate_grid()                                    # ImageGenerationDiffusionModels.jl:25-43
    data = matread(joinpath(@__DIR__, "..", "..", "..", "fixtures", "SyntheticImages500.mat"))
    images = reshape(data["syntheticImages"], 32, 32, 500)
    first64 = images[:, :, 1:64]
    canvas = zeros(Float32, 8 * 32, 8 * 32)
    for i in 0:7, j in 0:7
        canvas[i*32+1:(i+1)*32, j*32+1:(j+1)*32] .= first64[:, :, i * 8 + j + 1]
    end
    save("grid.png", colorview(Gray, clamp01.(canvas)))
    return canvas
end

function apply_noise(img; num_noise_steps = 500, beta_min = 0.0001, beta_max = 0.02)
    variance_schedule = beta_min : (beta_max - beta_min) / num_noise_steps : beta_max
    epsilon = randn(size(img))
    out = LibDDPM.apply_noise_f64(Float64.(img), epsilon, collect(variance_schedule))   # the loop of :65-67
    save("noisy_img.png", colorview(Gray, clamp01.(out)))
    return out
end

# ---- hot-path entry points --------------------------------------------------------------------
function generate_image(; model=default_model(), num_images=1, image_size=(32,32), seed=rand(UInt64))
    image_size == (32,32) || error("the trained U-Net only supports 32x32 images")
    h = engine()
    LibDDPM.set_weights!(h, flux_arrays(model))
    return LibDDPM.sample(h, num_images; seed=seed)          # generate_images.jl:231-245
end

function denoise_image(noisy_img::AbstractMatrix{<:Real}; model=default_model(), t_start::Int=100)
    size(noisy_img) == (32,32) || error("denoise_image expects a 32x32 matrix")
    h = engine()
    LibDDPM.set_weights!(h, flux_arrays(model))
    x = reshape(Float32.(2 .* noisy_img .- 1), 32, 32, 1, 1)
    out = LibDDPM.sample(h, 1; x_T=x, t_start=t_start, seed=rand(UInt64))
    denoised = clamp01.((out[:, :, 1, 1] .+ 1f0) ./ 2f0)
    save("denoised_img.png", colorview(Gray, denoised))
    return denoised
end

function batch_iterator(imgs::Array{Float32,4}, bs::Int)   # src/train_brain.jl:197-206, unchanged
    N = size(imgs, 4)
    return Channel{Array{Float32,4}}(c -> begin
        idx = randperm(N)
        for i in 1:bs:N
            put!(c, imgs[:, :, :, idx[i:min(i+bs-1, N)]])
        end
    end)
end

function train(data, lr=1f-4, epochs::Int=100, patience::Int=10, min_delta::Real=0.001; batch_size::Int=64)
    lr = Float32(lr)
    raw  = matread(data)["syntheticImages"]
    imgs = reshape(Float32.(raw), 32, 32, 1, :)
    imgs .*= 2; imgs .-= 1                                   # src/train_brain.jl:250-251
    model = SimpleUNet(1)
    opt   = Adam(lr)
    h = engine()
    LibDDPM.set_weights!(h, flux_arrays(model))
    LibDDPM.set_adam!(h, lr)
    losses = Float32[]; best_loss = Inf; epochs_no_improve = 0
    for epoch in 1:epochs
        total_loss, n = Float32(0), 0
        for x0 in batch_iterator(imgs, batch_size)
            B  = size(x0, 4)
            ts = Int32.(rand(1:T, B))                        # src/train_brain.jl:227
            ϵ  = randn(Float32, size(x0))                    # :228
            total_loss += LibDDPM.train_step!(h, x0, ts, ϵ)  # :230-241 + :267-272 in one call
            n += 1
        end
        epoch_loss = total_loss / n
        push!(losses, epoch_loss)
        @info "Epoch $epoch | avg loss = $epoch_loss"
        if epoch_loss < best_loss - min_delta
            best_loss = epoch_loss; epochs_no_improve = 0
        else
            epochs_no_improve += 1
        end
        if epochs_no_improve > patience
            @warn "Early stopping: No significant improvement for $(patience+1) epochs"
            break
        end
        if epoch % 5 == 0
            LibDDPM.get_weights!(h, flux_arrays(model))      # pull device weights into the Flux structs
            @save "ddpm_epoch_$epoch.bson" model opt epoch
        end
    end
    LibDDPM.get_weights!(h, flux_arrays(model))
    @save "trained_model.bson" model opt
    return model, losses
end

function demo()
    img = generate_grid()
    noisy = apply_noise(img[1:32, 1:32])
    den = denoise_image(clamp01.(img[1:32, 1:32]))
    new = generate_image()
    return (grid=img, noisy=noisy, denoised=den, generated=new)
end

end # module
