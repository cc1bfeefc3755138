# NNopB200.jl -- drop-in for NNop.jl's public API on NVIDIA B200 (sm_100a).
#
# Same function names, signatures, column-major (E, L, H, B) layout and ChainRules rrules as
# pxl-th/NNop.jl v0.2.0; every body is: validate -> allocate outputs with `similar` -> one
# `ccall` into libnnop_b200.so (C ABI: include/nnop_b200.h) on the task-local CUDA stream.
# It replaces ext/NNopCUDAExt.jl and the KernelAbstractions kernels of src/*.jl; there is no
# AMDGPU dispatch and no CPU fallback.  Two ways in: with the reference installed,
# `using NNop, NNopB200` loads ext/NNopB200NNopExt.jl, which adds these bodies as CuArray methods of
# NNop's own functions (the reference's extension seam, Project.toml:14-20); without it,
# `const NNop = NNopB200` gives user code the same names.  (Written without a Julia toolchain at hand -- the image
# this repo is built in has none -- so it is kept deliberately thin; the same ABI is exercised
# end to end by the Python twin in nnop.jl_b200/nnop_b200/.)
module NNopB200

using CUDA
using BFloat16s: BFloat16        # the element type CUDA.jl itself uses for bf16 CuArrays
import ChainRulesCore as CRC

const libnnop_b200 = get(ENV, "NNOP_B200_LIB", joinpath(@__DIR__, "..", "..", "..", "lib", "libnnop_b200.so"))
const Maybe{T} = Union{Nothing, T}                                   # src/NNop.jl:13
const FloatT = Union{Float32, Float16, BFloat16}

dtype_code(::Type{Float32}) = Cint(0)
dtype_code(::Type{Float16}) = Cint(1)
dtype_code(::Type{BFloat16}) = Cint(2)

function check(status::Cint)
    status == 0 && return
    msg = unsafe_string(ccall((:nnop_last_error_string, libnnop_b200), Cstring, ()))
    error(msg)                                   # reference: error("...") (src/attention.jl:141-144)
end

ptr(x::CuArray) = Base.unsafe_convert(CuPtr{Cvoid}, x)
ptr(::Nothing) = CuPtr{Cvoid}(0)
stream() = CUDA.stream().handle

# Workspaces are released right after the (enqueue-only) library call.  That is safe because CUDA.jl's
# memory pool is stream-ordered: `unsafe_free!` hands the block back with `cuMemFreeAsync` semantics on
# the task-local stream -- the stream the kernels were just enqueued on -- so the memory cannot be
# reused before they have run.  With a non-stream-ordered pool (`JULIA_CUDA_MEMORY_POOL=none`) set
# `NNOP_B200_EAGER_FREE=0` and the garbage collector releases them instead.
const EAGER_FREE = get(ENV, "NNOP_B200_EAGER_FREE", "1") != "0"
release!(ws::CuArray) = (EAGER_FREE && CUDA.unsafe_free!(ws); nothing)
release!(::Nothing) = nothing

# nnop_device_info: replaces NNop.shared_memory / _shared_memory (src/NNop.jl:27-30, ext/NNopCUDAExt.jl:6-9)
struct DeviceInfo
    sm_count::Cint
    cc_major::Cint
    cc_minor::Cint
    shared_mem_per_block_optin::Csize_t
    l2_bytes::Csize_t
    hbm_bytes::Csize_t
end
function device_info(device::Integer = CUDA.deviceid(CUDA.device()))
    info = Ref{DeviceInfo}()
    check(ccall((:nnop_device_info, libnnop_b200), Cint, (Cint, Ptr{DeviceInfo}), device, info))
    return info[]
end

within_gradient(x) = false                                                     # src/attention_crc.jl:1-2
CRC.rrule(::typeof(within_gradient), x) = true, _ -> (CRC.NoTangent(), CRC.NoTangent())

# ------------------------------------------------------------------ flash attention
# workspace size with the pair extension: round_up(base, 256) + nnop_flash_attn_pair_workspace_bytes
function pair_workspace(base, ::Type{T}, QL, KL, QH, B, backward::Bool) where T
    extra = ccall((:nnop_flash_attn_pair_workspace_bytes, libnnop_b200), Csize_t,
        (Cint, Cint, Cint, Cint, Cint, Cint), dtype_code(T), QL, KL, QH, B, backward)
    return Csize_t((base + 255) & ~Csize_t(255)) + extra
end

# _flash_attention: src/attention.jl:133-177.  Residuals: (o, lse, ls) -- one Float32 log-sum-exp
# replaces the reference's `ms`; the `ls` slot is `nothing`, or with a pair bias the forward workspace
# holding the head-major copy of `pair`.  They are private to the rrule closure.
function _flash_attention(
    q::CuArray{T,4}, k::CuArray{T,4}, v::CuArray{T,4},
    pair::Maybe{CuArray{T,4}} = nothing;
    causal::Bool, kpad_mask::Maybe{CuMatrix{Bool}} = nothing,
) where T <: FloatT
    QE, QL, QH, B = size(q)
    KE, KL, KH, KB = size(k)
    QE == KE || error("Embedding dim of Q `$QE` must be the same as of K `$KE`.")
    size(k) == size(v) || error("Shapes of K `$(size(k))` and V `$(size(v))` must be the same.")
    ispow2(QE) || error("Only power-of-2 embedding dims are supported.")
    QH % KH == 0 || error("Number of query heads `$QH` must be divisible by number of KV heads `$KH`.")

    o = similar(q)
    lse = CUDA.zeros(Float32, QL, QH, B)
    # optional workspace: lets Float32 (E = 64) run on the tensor cores (split fp16 operands) and, with
    # the pair extension, keeps the additive bias on the tensor-core path (head-major copy of `pair`)
    base = ccall((:nnop_flash_attn_fwd_workspace_bytes, libnnop_b200), Csize_t,
        (Cint, Cint, Cint, Cint, Cint, Cint, Cint), dtype_code(T), QE, QL, KL, QH, KH, B)
    nbytes = isnothing(pair) ? base : pair_workspace(base, T, QL, KL, QH, B, false)
    ws = nbytes > 0 ? CuArray{UInt8}(undef, nbytes) : nothing
    check(ccall((:nnop_flash_attn_fwd_ws, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cfloat, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        ptr(o), ptr(lse), ptr(q), ptr(k), ptr(v), ptr(pair), ptr(kpad_mask),
        dtype_code(T), QE, QL, KL, QH, KH, B, causal, Float32(inv(sqrt(QE))), ptr(ws), nbytes, stream()))
    # third residual (the reference's `ls` slot): with a pair bias, the workspace that now holds the
    # head-major copy of `pair`, and its offset -- the backward takes it instead of making its own
    # ... but only when the tensor-core path actually ran and wrote that copy (the SIMT kernels read
    # `pair` in place and leave the workspace untouched)
    if isnothing(pair) || ccall((:nnop_last_attention_path, libnnop_b200), Cint, ()) != 1
        release!(ws)
        return o, lse, nothing
    end
    return o, lse, (ws, Int((base + 255) & ~Csize_t(255)))
end

# ∇flash_attention: src/attention_bwd.jl:199-275 (`ms` carries lse, `ls` the forward's pair copy or nothing)
function ∇flash_attention(
    Δ::CuArray{T,4}, o::CuArray{T,4}, ms, ls,
    q::CuArray{T,4}, k::CuArray{T,4}, v::CuArray{T,4},
    pair::Maybe{CuArray{T,4}} = nothing;
    causal::Bool, kpad_mask::Maybe{CuMatrix{Bool}} = nothing,
) where T <: FloatT
    QE, QL, QH, B = size(q)
    _, KL, KH, _ = size(k)
    dq, dk, dv = similar(q), similar(k), similar(v)          # fully overwritten by the library
    dpair = isnothing(pair) ? nothing : similar(pair)
    nbytes = ccall((:nnop_flash_attn_bwd_workspace_bytes, libnnop_b200), Csize_t,
        (Cint, Cint, Cint, Cint, Cint, Cint, Cint), dtype_code(T), QE, QL, KL, QH, KH, B)
    # `ls` carries the forward's (workspace, offset) when a pair bias was given: its head-major copy of
    # `pair` is reused and this workspace only needs the dpair staging area
    reuse = !isnothing(pair) && ls isa Tuple
    isnothing(pair) || (nbytes = pair_workspace(nbytes, T, QL, KL, QH, B, !reuse))
    pair_hm = reuse ? ptr(ls[1]) + ls[2] : ptr(nothing)
    ws = CuArray{UInt8}(undef, max(nbytes, 1))
    check(ccall((:nnop_flash_attn_bwd_reuse_pair, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cfloat, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid},
         CuPtr{Cvoid}),
        ptr(dq), ptr(dk), ptr(dv), ptr(dpair), ptr(Δ), ptr(o), ptr(ms), ptr(q), ptr(k), ptr(v),
        ptr(pair), ptr(kpad_mask), dtype_code(T), QE, QL, KL, QH, KH, B, causal,
        Float32(inv(sqrt(QE))), ptr(ws), nbytes, stream(), pair_hm))
    release!(ws)
    return dq, dk, dv, dpair
end

function flash_attention(q, k, v, pair::Maybe{AbstractArray{<:Real,4}} = nothing;
                         causal::Bool, kpad_mask::Maybe{AbstractMatrix{Bool}} = nothing)   # src/attention_crc.jl:4-14
    o = _flash_attention(q, k, v, pair; causal, kpad_mask)
    within_gradient(q) && return o
    return o[1]
end

function CRC.rrule(::typeof(_flash_attention), q, k, v, pair::Maybe{AbstractArray{<:Real,4}} = nothing;
                   causal::Bool, kpad_mask::Maybe{AbstractMatrix{Bool}} = nothing)         # src/attention_crc.jl:16-31
    o, lse, ls = _flash_attention(q, k, v, pair; causal, kpad_mask)   # ls: forward workspace with the pair copy
    function _pullback(Δ)
        Δd = convert(typeof(o), CRC.unthunk(Δ))      # Zygote may hand a Fill / thunk: materialise
        dq, dk, dv, dpair = ∇flash_attention(Δd, o, lse, ls, q, k, v, pair; causal, kpad_mask)
        return CRC.NoTangent(), dq, dk, dv, (isnothing(dpair) ? CRC.NoTangent() : dpair)
    end
    return o, _pullback
end

# ------------------------------------------------------------------ packed variable-length attention
# Additive API (the reference only has the dense `kpad_mask`, src/attention.jl:73-79): sequences are
# concatenated along L, q (E, total_q, QH), k/v (E, total_k, KH), cu_seqlens :: CuVector{Int32} of
# nseq+1 row offsets (0-based, cu[1] == 0, cu[end] == total).  Float16 / BFloat16, E in (64, 128).
function _flash_attention_varlen(
    q::CuArray{T,3}, k::CuArray{T,3}, v::CuArray{T,3},
    cu_seqlens_q::CuVector{Int32}, cu_seqlens_k::CuVector{Int32},
    max_seqlen_q::Integer, max_seqlen_k::Integer; causal::Bool,
) where T <: Union{Float16, BFloat16}
    E, TQ, QH = size(q)
    KE, TK, KH = size(k)
    E == KE || error("Embedding dim of Q `$E` must be the same as of K `$KE`.")
    size(k) == size(v) || error("Shapes of K `$(size(k))` and V `$(size(v))` must be the same.")
    nseq = length(cu_seqlens_q) - 1
    o = similar(q)
    lse = CUDA.zeros(Float32, TQ, QH)
    nbytes = ccall((:nnop_flash_attn_varlen_fwd_workspace_bytes, libnnop_b200), Csize_t,
        (Cint, Cint, Cint, Int64, Cint), dtype_code(T), E, nseq, TQ, QH)
    ws = nbytes > 0 ? CuArray{UInt8}(undef, nbytes) : nothing     # tile counter of the persistent forward
    check(ccall((:nnop_flash_attn_varlen_fwd_ws, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         Cint, Cint, Cint, Int64, Int64, Cint, Cint, Cint, Cint, Cint, Cfloat, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        ptr(o), ptr(lse), ptr(q), ptr(k), ptr(v), ptr(cu_seqlens_q), ptr(cu_seqlens_k),
        nseq, max_seqlen_q, max_seqlen_k, TQ, TK, dtype_code(T), E, QH, KH, causal,
        Float32(inv(sqrt(E))), ptr(ws), nbytes, stream()))
    release!(ws)
    return o, lse
end

function ∇flash_attention_varlen(
    Δ::CuArray{T,3}, o::CuArray{T,3}, lse, q::CuArray{T,3}, k::CuArray{T,3}, v::CuArray{T,3},
    cu_seqlens_q::CuVector{Int32}, cu_seqlens_k::CuVector{Int32},
    max_seqlen_q::Integer, max_seqlen_k::Integer; causal::Bool,
) where T <: Union{Float16, BFloat16}
    E, TQ, QH = size(q)
    _, TK, KH = size(k)
    nseq = length(cu_seqlens_q) - 1
    dq, dk, dv = similar(q), similar(k), similar(v)
    nbytes = ccall((:nnop_flash_attn_varlen_bwd_workspace_bytes, libnnop_b200), Csize_t,
        (Cint, Cint, Cint, Int64, Cint), dtype_code(T), E, nseq, TQ, QH)
    ws = CuArray{UInt8}(undef, max(nbytes, 1))
    check(ccall((:nnop_flash_attn_varlen_bwd, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         Cint, Cint, Cint, Int64, Int64, Cint, Cint, Cint, Cint, Cint, Cfloat, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        ptr(dq), ptr(dk), ptr(dv), ptr(Δ), ptr(o), ptr(lse), ptr(q), ptr(k), ptr(v),
        ptr(cu_seqlens_q), ptr(cu_seqlens_k), nseq, max_seqlen_q, max_seqlen_k, TQ, TK,
        dtype_code(T), E, QH, KH, causal, Float32(inv(sqrt(E))), ptr(ws), nbytes, stream()))
    release!(ws)
    return dq, dk, dv
end

flash_attention_varlen(q, k, v, cu_q, cu_k, max_q, max_k; causal::Bool) =
    _flash_attention_varlen(q, k, v, cu_q, cu_k, max_q, max_k; causal)[1]

function CRC.rrule(::typeof(flash_attention_varlen), q, k, v, cu_q, cu_k, max_q, max_k; causal::Bool)
    o, lse = _flash_attention_varlen(q, k, v, cu_q, cu_k, max_q, max_k; causal)
    function _pullback(Δ)
        Δd = convert(typeof(o), CRC.unthunk(Δ))
        dq, dk, dv = ∇flash_attention_varlen(Δd, o, lse, q, k, v, cu_q, cu_k, max_q, max_k; causal)
        nt = CRC.NoTangent()
        return nt, dq, dk, dv, nt, nt, nt, nt
    end
    return o, _pullback
end

# ------------------------------------------------------------------ sequence-sharded ("ring") attention
# Additive (the reference has no multi-GPU path): one process, rank r = the r-th array of each vector, each
# on its own device.  qs[r] (E, Ll, QH, B), ks[r] / vs[r] (E, Ll, KH, B); causal inputs in zig-zag order
# (rank r holds chunks r and 2W-1-r of 2W).  K / V blocks move device to device over NVLink inside
# nnop_ring_attn_fwd / _bwd (include/nnop_b200.h); every rank's work is enqueued on that device's
# task-local stream.
on_device(f, x::CuArray) = CUDA.device!(f, CUDA.device(x))
dev_ptrs(xs) = CuPtr{Cvoid}[ptr(x) for x in xs]
dev_ids(xs) = Cint[CUDA.deviceid(CUDA.device(x)) for x in xs]
dev_streams(xs) = Ptr{Cvoid}[on_device(() -> reinterpret(Ptr{Cvoid}, CUDA.stream().handle), x) for x in xs]

function _ring_flash_attention(qs::Vector{<:CuArray{T,4}}, ks::Vector{<:CuArray{T,4}},
                               vs::Vector{<:CuArray{T,4}}; causal::Bool) where T <: FloatT
    W = length(qs)
    E, Ll, QH, B = size(qs[1])
    KH = size(ks[1], 3)
    os = [on_device(() -> similar(q), q) for q in qs]
    lses = [on_device(() -> CUDA.zeros(Float32, Ll, QH, B), q) for q in qs]
    nbytes = ccall((:nnop_ring_attn_fwd_workspace_bytes, libnnop_b200), Csize_t,
        (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint), dtype_code(T), E, Ll, QH, KH, B, W, causal)
    wss = [on_device(() -> CuArray{UInt8}(undef, max(nbytes, 1)), q) for q in qs]
    check(ccall((:nnop_ring_attn_fwd, libnnop_b200), Cint,
        (Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}},
         Ptr{Cint}, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cfloat,
         Ptr{CuPtr{Cvoid}}, Csize_t, Ptr{Ptr{Cvoid}}),
        dev_ptrs(os), dev_ptrs(lses), dev_ptrs(qs), dev_ptrs(ks), dev_ptrs(vs),
        dev_ids(qs), W, dtype_code(T), E, Ll, QH, KH, B, causal, Float32(inv(sqrt(E))),
        dev_ptrs(wss), nbytes, dev_streams(qs)))
    foreach(ws -> on_device(() -> release!(ws), ws), wss)
    return os, lses
end

function ∇ring_flash_attention(Δs::Vector{<:CuArray{T,4}}, os, lses, qs::Vector{<:CuArray{T,4}},
                               ks::Vector{<:CuArray{T,4}}, vs::Vector{<:CuArray{T,4}}; causal::Bool) where T <: FloatT
    W = length(qs)
    E, Ll, QH, B = size(qs[1])
    KH = size(ks[1], 3)
    dqs = [on_device(() -> similar(q), q) for q in qs]
    dks = [on_device(() -> similar(k), k) for k in ks]
    dvs = [on_device(() -> similar(v), v) for v in vs]
    nbytes = ccall((:nnop_ring_attn_bwd_workspace_bytes, libnnop_b200), Csize_t,
        (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint), dtype_code(T), E, Ll, QH, KH, B, W, causal)
    wss = [on_device(() -> CuArray{UInt8}(undef, max(nbytes, 1)), q) for q in qs]
    check(ccall((:nnop_ring_attn_bwd, libnnop_b200), Cint,
        (Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}},
         Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}}, Ptr{CuPtr{Cvoid}},
         Ptr{Cint}, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cfloat,
         Ptr{CuPtr{Cvoid}}, Csize_t, Ptr{Ptr{Cvoid}}),
        dev_ptrs(dqs), dev_ptrs(dks), dev_ptrs(dvs), dev_ptrs(Δs), dev_ptrs(os), dev_ptrs(lses),
        dev_ptrs(qs), dev_ptrs(ks), dev_ptrs(vs),
        dev_ids(qs), W, dtype_code(T), E, Ll, QH, KH, B, causal, Float32(inv(sqrt(E))),
        dev_ptrs(wss), nbytes, dev_streams(qs)))
    foreach(ws -> on_device(() -> release!(ws), ws), wss)
    return dqs, dks, dvs
end

ring_flash_attention(qs, ks, vs; causal::Bool) = _ring_flash_attention(qs, ks, vs; causal)[1]

function CRC.rrule(::typeof(ring_flash_attention), qs, ks, vs; causal::Bool)
    os, lses = _ring_flash_attention(qs, ks, vs; causal)
    function _pullback(Δs)
        Δd = [convert(typeof(o), CRC.unthunk(Δ)) for (o, Δ) in zip(os, CRC.unthunk(Δs))]
        dqs, dks, dvs = ∇ring_flash_attention(Δd, os, lses, qs, ks, vs; causal)
        return CRC.NoTangent(), dqs, dks, dvs
    end
    return os, _pullback
end

# building blocks for hosts that drive the ring themselves (one process per GPU, NCCL.jl / MPI transport;
# the Python twin nnop_b200/ring.py does exactly that): fold a partial result, accumulate a partial gradient
function attn_merge!(o_acc::CuArray{Float32}, lse_acc::CuArray{Float32}, lse_out::CuArray{Float32},
                     o_part::CuArray{T}, lse_part::CuArray{Float32}; init::Bool) where T <: FloatT
    check(ccall((:nnop_attn_merge, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Cint, Int64, Cint, Ptr{Cvoid}),
        ptr(o_acc), ptr(lse_acc), ptr(lse_out), ptr(o_part), ptr(lse_part), dtype_code(T), size(o_part, 1),
        length(lse_part), init, stream()))
    return lse_out
end

function accumulate_f32!(acc::CuArray{Float32}, part::CuArray{T}; init::Bool) where T <: FloatT
    check(ccall((:nnop_accumulate_f32, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Int64, Cint, Ptr{Cvoid}),
        ptr(acc), ptr(part), dtype_code(T), length(part), init, stream()))
    return acc
end

# out[:, row_offset+1 : row_offset+rows, slab] = T(acc[:, :, slab]) for acc (E, rows, slabs...), out (E, out_rows, slabs...)
function store_rows_from_f32!(out::CuArray{T}, acc::CuArray{Float32}; row_offset::Integer = 0) where T <: FloatT
    E, rows = size(acc, 1), size(acc, 2)
    check(ccall((:nnop_store_rows_from_f32, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Int64, Ptr{Cvoid}),
        ptr(out), ptr(acc), dtype_code(T), E, length(acc) ÷ (E * rows), rows, size(out, 2), row_offset, stream()))
    return out
end

# ------------------------------------------------------------------ online softmax (src/softmax.jl:60-86)
function online_softmax(x::CuMatrix{T}) where T <: FloatT
    y = similar(x)
    check(ccall((:nnop_softmax_fwd, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}),
        ptr(y), ptr(x), dtype_code(T), size(x, 1), size(x, 2), stream()))
    return y
end

function ∇online_softmax(Δ::CuMatrix{T}, y::CuMatrix{T}) where T <: FloatT
    if within_gradient(y)      # second-order AD: stay differentiable, exactly as the reference (src/softmax.jl:71-74)
        tmp = Δ .* y
        return tmp .- y .* sum(tmp; dims=1)
    end
    dx = similar(y)
    check(ccall((:nnop_softmax_bwd, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}),
        ptr(dx), ptr(Δ), ptr(y), dtype_code(T), size(y, 1), size(y, 2), stream()))
    return dx
end

function CRC.rrule(::typeof(online_softmax), x)
    y = online_softmax(x)
    _pullback(Δ) = (CRC.NoTangent(), ∇online_softmax(convert(typeof(y), CRC.unthunk(Δ)), y))
    return y, _pullback
end

# ------------------------------------------------------------------ RMS norm (src/rms_norm.jl:117-185)
function norm_workspace(emb, n)
    nbytes = ccall((:nnop_norm_bwd_workspace_bytes, libnnop_b200), Csize_t, (Int64, Int64), emb, n)
    return CuArray{UInt8}(undef, max(nbytes, 1)), nbytes
end

function _rms_norm(x::CuMatrix{T}, w::CuVector{T}; ϵ::Float32, offset::Float32 = 0f0) where T <: FloatT
    emb, n = size(x)
    @assert emb == length(w)
    y = similar(x)
    rms = CUDA.zeros(Float32, n)
    check(ccall((:nnop_rms_norm_fwd, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Int64, Int64, Cfloat, Cfloat, Ptr{Cvoid}),
        ptr(y), ptr(rms), ptr(x), ptr(w), dtype_code(T), emb, n, ϵ, offset, stream()))
    return y, rms
end

function ∇rms_norm(Δ::CuMatrix{T}, rms, x::CuMatrix{T}, w::CuVector{T}; offset::Float32) where T <: FloatT
    emb, n = size(x)
    dx = similar(x)
    dw = CUDA.zeros(Float32, emb)                               # Float32 for every T (src/rms_norm.jl:146)
    ws, nbytes = norm_workspace(emb, n)
    check(ccall((:nnop_rms_norm_bwd, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         Cint, Int64, Int64, Cfloat, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        ptr(dx), ptr(dw), ptr(Δ), ptr(rms), ptr(x), ptr(w), dtype_code(T), emb, n, offset,
        ptr(ws), nbytes, stream()))
    release!(ws)
    return dx, dw
end

function rms_norm(x, w; ϵ::Float32 = 1f-6, offset::Float32 = 0f0)      # src/rms_norm.jl:171-176
    y = _rms_norm(x, w; ϵ, offset)
    within_gradient(x) && return y
    return y[1]
end

function CRC.rrule(::typeof(_rms_norm), x, w; ϵ::Float32 = 1f-6, offset::Float32 = 0f0)
    y, rms = _rms_norm(x, w; ϵ, offset)
    function _pullback(Δ)
        dx, dw = ∇rms_norm(convert(typeof(y), CRC.unthunk(Δ)), rms, x, w; offset)
        return CRC.NoTangent(), dx, dw
    end
    return y, _pullback
end

# ------------------------------------------------------------------ layer norm (src/layer_norm.jl:150-220)
function _layer_norm(x::CuMatrix{T}, w::CuVector{T}, b::CuVector{T}; ϵ::Float32) where T <: FloatT
    emb, n = size(x)
    @assert emb == length(w) == length(b)
    y = similar(x)
    μ = CUDA.zeros(Float32, n)
    Σ = CUDA.zeros(Float32, n)                                   # holds rstd (src/layer_norm.jl:50)
    check(ccall((:nnop_layer_norm_fwd, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         Cint, Int64, Int64, Cfloat, Ptr{Cvoid}),
        ptr(y), ptr(μ), ptr(Σ), ptr(x), ptr(w), ptr(b), dtype_code(T), emb, n, ϵ, stream()))
    return y, μ, Σ
end

function ∇layer_norm(Δ::CuMatrix{T}, μ, Σ, x::CuMatrix{T}, w::CuVector{T}, b::CuVector{T}) where T <: FloatT
    emb, n = size(x)
    dx, dw, db = similar(x), similar(w), similar(b)
    ws, nbytes = norm_workspace(emb, n)
    check(ccall((:nnop_layer_norm_bwd, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         CuPtr{Cvoid}, Cint, Int64, Int64, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        ptr(dx), ptr(dw), ptr(db), ptr(Δ), ptr(μ), ptr(Σ), ptr(x), ptr(w), dtype_code(T), emb, n,
        ptr(ws), nbytes, stream()))
    release!(ws)
    return dx, dw, db
end

function layer_norm(x, w, b; ϵ::Float32 = 1f-6)                       # src/layer_norm.jl:206-211
    y = _layer_norm(x, w, b; ϵ)
    within_gradient(x) && return y
    return y[1]
end

function CRC.rrule(::typeof(_layer_norm), x, w, b; ϵ::Float32 = 1f-6)
    y, μ, Σ = _layer_norm(x, w, b; ϵ)
    function _pullback(Δ)
        dx, dw, db = ∇layer_norm(convert(typeof(y), CRC.unthunk(Δ)), μ, Σ, x, w, b)
        return CRC.NoTangent(), dx, dw, db
    end
    return y, _pullback
end

# ------------------------------------------------------------------ Llama RoPE (src/rope/llama_rope.jl)
struct LlamaRotaryEmbedding{F <: AbstractVector{Float32}}
    inv_freq::F
    dim::Int
    base::Int
end

function LlamaRotaryEmbedding(dim::Int; base::Int = 10000)            # :7-11, host side, Float32
    ids = (0f0:2f0:Float32(dim - 1)) ./ Float32(dim)
    LlamaRotaryEmbedding(inv.(base .^ ids), dim, base)
end

function (emb::LlamaRotaryEmbedding)(position_ids::AbstractMatrix{Float32})   # :15-22
    position_ids = reshape(position_ids, 1, size(position_ids)...)
    freqs = emb.inv_freq .* position_ids
    freqs = vcat(freqs, freqs)
    return cos.(freqs), sin.(freqs)
end

function _llama_rope(q::CuArray{T,4}, k::CuArray{T,4}, cos::CuArray{Float32,3}, sin::CuArray{Float32,3};
                     bwd::Bool) where T <: FloatT
    @assert size(q, 1) == size(k, 1)
    @assert size(q, 2) == size(k, 2)
    @assert size(q, 4) == size(k, 4)
    qo, ko = similar(q), similar(k)            # out-of-place: fuses the reference's copy (:75-76)
    check(ccall((:nnop_llama_rope, libnnop_b200), Cint,
        (CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
         Cint, Cint, Int64, Cint, Cint, Cint, Cfloat, Ptr{Cvoid}),
        ptr(qo), ptr(ko), ptr(q), ptr(k), ptr(cos), ptr(sin), dtype_code(T),
        size(q, 1), size(q, 2), size(q, 3), size(k, 3), size(q, 4), bwd ? -1f0 : 1f0, stream()))
    return qo, ko
end

llama_rope(q, k; cos, sin) = _llama_rope(q, k, cos, sin; bwd=false)      # :91
∇llama_rope(dq, dk; cos, sin) = _llama_rope(dq, dk, cos, sin; bwd=true)  # :92

function CRC.rrule(::typeof(llama_rope), q, k; cos, sin)                   # :94-98
    qr, kr = llama_rope(q, k; cos, sin)
    function _pullback(Δ)
        dq, dk = CRC.unthunk.(Δ)
        return (CRC.NoTangent(), ∇llama_rope(convert(typeof(q), dq), convert(typeof(k), dk); cos, sin)...)
    end
    return (qr, kr), _pullback
end

# ------------------------------------------------------------------ diagnostics
# kernel-variant switches of the tcgen05 attention path (process-wide; include/nnop_b200_diag.h -- not part of
# the drop-in ABI)
set_attention_path(mode::Integer) = check(ccall((:nnop_set_attention_path, libnnop_b200), Cint, (Cint,), mode))
last_attention_path() = ccall((:nnop_last_attention_path, libnnop_b200), Cint, ())
set_fwd_mode(mode::Integer) = check(ccall((:nnop_set_fwd_mode, libnnop_b200), Cint, (Cint,), mode))
set_bwd_mode(mode::Integer) = check(ccall((:nnop_set_bwd_pair_mode, libnnop_b200), Cint, (Cint,), mode))

end # module
