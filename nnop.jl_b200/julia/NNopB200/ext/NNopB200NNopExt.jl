# NNopB200NNopExt.jl -- the drop-in seam, in the reference's own style.
#
# NNop.jl plugs a backend in through a package extension keyed on a weak dependency
# (Project.toml:14-20; ext/NNopCUDAExt.jl:1-11 overrides `NNop._shared_memory(::CUDABackend, ...)`).
# This extension of NNopB200 is loaded the same way, when both `NNop` and `NNopB200` are in the
# session (`using NNop, NNopB200`), and adds CuArray methods to NNop's OWN launcher functions --
# the L2 layer of SURVEY.md section 1 -- so that unmodified user code written against the reference,
#
#     NNop.flash_attention(q, k, v; causal=true);  Zygote.gradient(... NNop.rms_norm ...)
#
# runs libnnop_b200.so instead of the KernelAbstractions kernels.  NNop's public wrappers and its
# ChainRules rrules (src/attention_crc.jl:4-31, src/softmax.jl:82-86, src/rms_norm.jl:171-185,
# src/layer_norm.jl:206-220, src/rope/llama_rope.jl:91-98) stay as they are: they call
# `_flash_attention` / `∇flash_attention` / `_rms_norm` / ... generically, and Julia's dispatch picks
# the more specific CuArray methods below.  Residual contents differ where they are private to the
# rrule closures: `ms` carries one Float32 log-sum-exp and `ls` is `nothing` or the forward workspace
# holding the head-major copy of `pair` (NNopB200.jl).  Without the reference installed the same code
# runs with `const NNop = NNopB200`.
module NNopB200NNopExt

using CUDA
using NNop
using NNopB200
using NNopB200: FloatT, Maybe

# cotangents may arrive as a `Fill` / broadcasted array (NNop's rrules only `unthunk`): make them dense
dense(like::CuArray, Δ) = Δ isa typeof(like) ? Δ : convert(typeof(like), Δ)

# ---- flash attention (src/attention.jl:133-137, src/attention_bwd.jl:199-206)
function NNop._flash_attention(
    q::CuArray{T,4}, k::CuArray{T,4}, v::CuArray{T,4}, pair::Maybe{CuArray{T,4}} = nothing;
    causal::Bool, kpad_mask::Maybe{CuMatrix{Bool}} = nothing,
) where T <: FloatT
    return NNopB200._flash_attention(q, k, v, pair; causal, kpad_mask)
end

function NNop.∇flash_attention(
    Δ::AbstractArray{<:Real,4}, o::CuArray{T,4}, ms, ls,
    q::CuArray{T,4}, k::CuArray{T,4}, v::CuArray{T,4}, pair::Maybe{CuArray{T,4}} = nothing;
    causal::Bool, kpad_mask::Maybe{CuMatrix{Bool}} = nothing,
) where T <: FloatT
    return NNopB200.∇flash_attention(dense(o, Δ), o, ms, ls, q, k, v, pair; causal, kpad_mask)
end

# ---- online softmax (src/softmax.jl:60, :70)
NNop.online_softmax(x::CuMatrix{T}) where T <: FloatT = NNopB200.online_softmax(x)
NNop.∇online_softmax(Δ::AbstractMatrix, y::CuMatrix{T}) where T <: FloatT =
    NNopB200.∇online_softmax(dense(y, Δ), y)

# ---- RMS norm (src/rms_norm.jl:117, :139)
NNop._rms_norm(x::CuMatrix{T}, w::CuVector{T}; ϵ::Float32, offset::Float32 = 0f0) where T <: FloatT =
    NNopB200._rms_norm(x, w; ϵ, offset)
NNop.∇rms_norm(Δ::AbstractMatrix, rms, x::CuMatrix{T}, w::CuVector{T}; offset::Float32) where T <: FloatT =
    NNopB200.∇rms_norm(dense(x, Δ), rms, x, w; offset)

# ---- layer norm (src/layer_norm.jl:150, :172)
NNop._layer_norm(x::CuMatrix{T}, w::CuVector{T}, b::CuVector{T}; ϵ::Float32 = 1f-6) where T <: FloatT =
    NNopB200._layer_norm(x, w, b; ϵ)
NNop.∇layer_norm(Δ::AbstractMatrix, μ, Σ, x::CuMatrix{T}, w::CuVector{T}, b::CuVector{T}) where T <: FloatT =
    NNopB200.∇layer_norm(dense(x, Δ), μ, Σ, x, w, b)

# ---- Llama RoPE (src/rope/llama_rope.jl:69); llama_rope / ∇llama_rope / the rrule call it generically
NNop._llama_rope(q::CuArray{T,4}, k::CuArray{T,4}, cos::CuArray{Float32,3}, sin::CuArray{Float32,3};
                 bwd::Bool) where T <: FloatT = NNopB200._llama_rope(q, k, cos, sin; bwd)

# ---- the reference's one backend hook (ext/NNopCUDAExt.jl:6-9): answered by nnop_device_info with the
# opt-in shared memory per block of the device (227 KB on B200) instead of the 48 KB static limit
NNop._shared_memory(::CUDABackend, device_id::Integer) =
    UInt64(NNopB200.device_info(device_id - 1).shared_mem_per_block_optin)

end # module
