"""nnop_b200 -- host-side mirror of NNop.jl's public API over libnnop_b200.so.

Same operator names, argument meaning and error behaviour as the reference
(`NNop.flash_attention`, `online_softmax`, `rms_norm`, `layer_norm`, `llama_rope`,
`LlamaRotaryEmbedding`; src/NNop.jl:15-25), with torch CUDA tensors standing in for CuArrays
and `torch.autograd.Function`s standing in for the ChainRules rrules.  A Julia `(E, L, H, B)`
array is passed as the row-major tensor `(B, H, L, E)` holding the same bytes.
"""
from ._lib import NNopError, LIB_PATH, lib  # noqa: F401
from .ops import (  # noqa: F401
    flash_attention, _flash_attention, grad_flash_attention,
    flash_attention_varlen, _flash_attention_varlen, grad_flash_attention_varlen,
    online_softmax, grad_online_softmax,
    rms_norm, _rms_norm, grad_rms_norm,
    layer_norm, _layer_norm, grad_layer_norm,
    llama_rope, grad_llama_rope, LlamaRotaryEmbedding,
    device_info, set_attention_path, last_attention_path, selftest_umma, set_bwd_pair_mode, set_fwd_mode,
    set_timing_events, HostAttentionPipeline,
)
from .sharding import shard_slices, shard_attention_inputs  # noqa: F401
from .ring import (  # noqa: F401
    ring_flash_attention, ring_attention_forward, ring_attention_backward, ring_schedule,
    zigzag_shard, zigzag_unshard, contiguous_shard, p2p_ring_attention_forward, p2p_ring_attention_backward,
)

__all__ = [
    "flash_attention", "_flash_attention", "grad_flash_attention", "online_softmax",
    "grad_online_softmax", "rms_norm", "_rms_norm", "grad_rms_norm", "layer_norm",
    "_layer_norm", "grad_layer_norm", "llama_rope", "grad_llama_rope", "LlamaRotaryEmbedding",
    "device_info", "set_attention_path", "last_attention_path", "selftest_umma",
    "shard_slices", "shard_attention_inputs", "NNopError", "set_timing_events",
    "HostAttentionPipeline", "flash_attention_varlen", "_flash_attention_varlen",
    "grad_flash_attention_varlen", "set_bwd_pair_mode", "set_fwd_mode", "ring_flash_attention", "ring_attention_forward",
    "ring_attention_backward", "ring_schedule", "zigzag_shard", "zigzag_unshard", "contiguous_shard",
    "p2p_ring_attention_forward", "p2p_ring_attention_backward",
]
# the reference spells its pullbacks with a nabla; reachable via getattr(nnop_b200, "∇flash_attention")
globals().update({
    "∇flash_attention": grad_flash_attention, "∇online_softmax": grad_online_softmax,
    "∇rms_norm": grad_rms_norm, "∇layer_norm": grad_layer_norm, "∇llama_rope": grad_llama_rope,
})
