# LibDDPM.jl -- `ccall` bindings of libddpm.so (include/libddpm.h).  One function per C entry point;
# every wrapper turns a nonzero return code into `error(ddpm_last_error())`.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The identical signatures are
# exercised through ctypes (imagegenerationdiffusionmodels.jl_b200/capi.py) by tests/ and bench.py.
module LibDDPM

const libddpm = get(ENV, "LIBDDPM", joinpath(@__DIR__, "..", "..", "libddpm.so"))

const PREC_FP32 = Cint(0)
const PREC_FP16 = Cint(1)
const PREC_BF16 = Cint(2)
const PREC_TF32 = Cint(3)
const NUM_ARRAYS = 64

last_error() = unsafe_string(ccall((:ddpm_last_error, libddpm), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : error("libddpm: " * last_error())

mutable struct Handle
    ptr::Ptr{Cvoid}
    T::Int
    function Handle(; T::Int=500, D::Int=128, H::Int=32, W::Int=32, precision::Cint=PREC_FP16, device::Int=0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ddpm_create, libddpm), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint, Cint),
                    out, T, D, H, W, precision, device))
        h = new(out[], T)
        finalizer(h) do x
            x.ptr == C_NULL || ccall((:ddpm_destroy, libddpm), Cint, (Ptr{Cvoid},), x.ptr)
            x.ptr = C_NULL
        end
        return h
    end
end

device_count() = Int(ccall((:ddpm_device_count, libddpm), Cint, ()))

function array_lengths()
    lens = Vector{Int64}(undef, NUM_ARRAYS)
    check(ccall((:ddpm_array_lengths, libddpm), Cint, (Ptr{Int64},), lens))
    return lens
end

"beta, alpha_cum :: Vector{Float32}(T); pe :: Matrix{Float32}(D, T) (column t == timestep_embedding(t))"
function set_tables!(h::Handle, beta::Vector{Float32}, alpha_cum::Vector{Float32}, pe::Matrix{Float32})
    check(ccall((:ddpm_set_tables, libddpm), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
                h.ptr, beta, alpha_cum, pe))
end

"arrays: the 64 Float32 arrays of a SimpleUNet in BSON order (see flux_arrays in the package module)"
function set_weights!(h::Handle, arrays::Vector{<:Array{Float32}})
    lens = Int64[length(a) for a in arrays]
    ptrs = Ptr{Float32}[pointer(a) for a in arrays]
    GC.@preserve arrays check(ccall((:ddpm_set_weights, libddpm), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Float32}}, Ptr{Int64}, Cint), h.ptr, ptrs, lens, length(arrays)))
end

function get_weights!(h::Handle, arrays::Vector{<:Array{Float32}})
    lens = Int64[length(a) for a in arrays]
    ptrs = Ptr{Float32}[pointer(a) for a in arrays]
    GC.@preserve arrays check(ccall((:ddpm_get_weights, libddpm), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Float32}}, Ptr{Int64}, Cint), h.ptr, ptrs, lens, length(arrays)))
    return arrays
end

set_adam!(h::Handle, eta::Float32; beta=(0.9f0, 0.999f0), eps::Float32=1f-8) =
    check(ccall((:ddpm_set_adam, libddpm), Cint, (Ptr{Cvoid}, Cfloat, Cfloat, Cfloat, Cfloat), h.ptr, eta, beta[1], beta[2], eps))

"Adam moments (64 arrays each, weight order), (beta1^t, beta2^t) of the next update and the applied-step count."
function get_adam_state(h::Handle)
    lens = array_lengths()
    m = [zeros(Float32, n) for n in lens]; v = [zeros(Float32, n) for n in lens]
    pm = Ptr{Float32}[pointer(a) for a in m]; pv = Ptr{Float32}[pointer(a) for a in v]
    bt = zeros(Float32, 2); steps = Ref{Int64}(0)
    GC.@preserve m v check(ccall((:ddpm_get_adam_state, libddpm), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Float32}}, Ptr{Ptr{Float32}}, Ptr{Int64}, Cint, Ptr{Float32}, Ref{Int64}),
        h.ptr, pm, pv, lens, length(lens), bt, steps))
    return m, v, (bt[1], bt[2]), steps[]
end

function set_adam_state!(h::Handle, m::Vector{Vector{Float32}}, v::Vector{Vector{Float32}}, beta_t, steps::Integer)
    lens = Int64[length(a) for a in m]
    pm = Ptr{Float32}[pointer(a) for a in m]; pv = Ptr{Float32}[pointer(a) for a in v]
    bt = Float32[beta_t[1], beta_t[2]]
    GC.@preserve m v check(ccall((:ddpm_set_adam_state, libddpm), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Float32}}, Ptr{Ptr{Float32}}, Ptr{Int64}, Cint, Ptr{Float32}, Int64),
        h.ptr, pm, pv, lens, length(lens), bt, Int64(steps)))
end

function q_sample(h::Handle, x0::Array{Float32,4}, ts::Vector{Int32}, eps::Array{Float32,4})
    xt = similar(x0)
    check(ccall((:ddpm_q_sample, libddpm), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Int32}, Ptr{Float32}, Cint, Ptr{Float32}),
                h.ptr, x0, ts, eps, size(x0, 4), xt))
    return xt
end

function predict_eps(h::Handle, xt::Array{Float32,4}, ts::Vector{Int32}; train_mode::Bool=false)
    out = similar(xt)
    check(ccall((:ddpm_predict_eps, libddpm), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Int32}, Cint, Cint, Ptr{Float32}),
                h.ptr, xt, ts, size(xt, 4), train_mode ? 1 : 0, out))
    return out
end

"One iteration of the training-loop body (src/train_brain.jl:267-274); returns the Float32 loss."
function train_step!(h::Handle, x0::Array{Float32,4}, ts::Vector{Int32}, eps::Array{Float32,4})
    loss = Ref{Cfloat}(0)
    check(ccall((:ddpm_train_step, libddpm), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Int32}, Ptr{Float32}, Cint, Ref{Cfloat}),
                h.ptr, x0, ts, eps, size(x0, 4), loss))
    return loss[]
end

"""
    sample(h, N; x_T=nothing, z=nothing, seed=0, first_index=0, t_start=h.T)

The reverse loop of `generate_image` (src/generate_images.jl:231-245).  `x_T :: Array{Float32,4}(32,32,1,N)`
and `z :: Array{Float32,5}(32,32,1,N,t_start-1)` are optional host-supplied draws (`z[:,:,:,:,k]` is used at
step `t = t_start-k+1`); otherwise the device Philox generator keyed by (seed, first_index+i, step) is used.
"""
function sample(h::Handle, N::Integer; x_T=nothing, z=nothing, seed::Integer=0, first_index::Integer=0, t_start::Integer=h.T)
    out = Array{Float32,4}(undef, 32, 32, 1, N)
    xp = x_T === nothing ? Ptr{Float32}(C_NULL) : pointer(x_T)
    zp = z === nothing ? Ptr{Float32}(C_NULL) : pointer(z)
    GC.@preserve x_T z check(ccall((:ddpm_sample, libddpm), Cint,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, UInt64, Int64, Int64, Cint, Ptr{Float32}),
        h.ptr, xp, zp, UInt64(seed), Int64(N), Int64(first_index), Cint(t_start), out))
    return out
end

"Bulk dumps: sample N images on the device, then fetch them quantised to 8 bits ((x+1)/2 -> N0f8) -- 1 byte per pixel."
function sample_u8(h::Handle, N::Integer; seed::Integer=0, first_index::Integer=0, t_start::Integer=h.T)
    check(ccall((:ddpm_sample_device, libddpm), Cint, (Ptr{Cvoid}, UInt64, Int64, Int64, Cint),
                h.ptr, UInt64(seed), Int64(N), Int64(first_index), Cint(t_start)))
    out = Array{UInt8,4}(undef, 32, 32, 1, N)
    check(ccall((:ddpm_sample_fetch_u8, libddpm), Cint, (Ptr{Cvoid}, Int64, Ptr{UInt8}), h.ptr, Int64(N), out))
    return out
end

"`apply_noise` recurrence in Float64 over host-supplied betas (src/ImageGenerationDiffusionModels.jl:60-73)."
function apply_noise_f64(img::Array{Float64}, eps::Array{Float64}, betas::Vector{Float64})
    out = similar(img)
    check(ccall((:ddpm_apply_noise_f64, libddpm), Cint, (Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Cint, Ptr{Float64}),
                img, eps, length(img), betas, length(betas), out))
    return out
end

function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:ddpm_comm_unique_id, libddpm), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init!(h::Handle, id::Vector{UInt8}, rank::Integer, world::Integer; sync_bn::Bool=true) =
    check(ccall((:ddpm_comm_init, libddpm), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, Cint), h.ptr, id, rank, world, sync_bn ? 1 : 0))

end # module
