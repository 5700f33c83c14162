"""Drop-in attention mechanisms (mirror of /root/reference/src/models/attention.py).

Same class names, constructor signature (`config` with attention_type / num_heads / temperature /
hidden_dim), parameter names (state_dict-compatible) and `forward(query, key, value,
key_padding_mask=None, **kwargs) -> (context, weights)` contract as the reference.  The arithmetic runs
in libcapdec (capdec_attention_forward): hoisted key/value projections + one fused
score/softmax/context kernel per image.  Supported call shape is the one the decoders use
(src/models/decoders.py:287-294): 2-D query [R,H] with key is value = region features [B,L,H],
R == B (or R == B*rows_per_image via the `rows_per_image` kwarg, rows grouped by image).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _capi
from .config import AttentionConfig, AttentionType, attention_kind
from .engine import Engine, EngineOwner


class AttentionMechanism(EngineOwner, nn.Module):
    """Base class (attention.py:9-35)."""

    _kind = None
    precision = "fp32"

    def _engine(self, device) -> Engine:
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (str(device), self.precision)
        if getattr(self, "_eng_sig", None) != sig:
            H = self.hidden_dim
            cfg = _capi.Config(arch=_capi.ARCH_LSTM, attention=_capi.ATT[self._kind], precision=_capi.PREC[self.precision],
                               vocab_size=4, hidden_dim=H, embed_dim=H, feature_dim=H, attention_dim=H, num_layers=1,
                               num_heads=int(self._num_heads), temperature=float(self._temperature),
                               pad_token_id=0, bos_token_id=1, eos_token_id=2)
            sd = {"attention." + k: v for k, v in self.state_dict().items()}
            object.__setattr__(self, "_eng", _AttentionOnlyEngine(cfg, sd, device))
            object.__setattr__(self, "_eng_sig", sig)
        return self._eng

    def forward(self, query, key, value, key_padding_mask=None, memory_state=None, cell_state=None,
                rows_per_image: int = 1, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        if query.dim() != 2:
            raise NotImplementedError("capdec attention supports the decoders' 2-D query [R,H] call shape only")
        if value is not key and not (value.data_ptr() == key.data_ptr() and value.shape == key.shape):
            raise NotImplementedError("capdec attention requires value is key (region features), as in decoders.py:287-290")
        if self._kind == "adaptive":
            assert memory_state is not None and cell_state is not None, \
                "AdaptiveAttention requires memory_state and cell_state"
        eng = self._engine(query.device)
        return eng.attention_forward(query, key, key_padding_mask, memory_state, cell_state, rows_per_image)


class _AttentionOnlyEngine(Engine):
    """Engine over an attention module alone: the decoder weights the handle would also want are absent,
    so bind zero placeholders for them (they are never read by capdec_attention_forward)."""

    def __init__(self, cfg, sd, device):
        H = cfg.hidden_dim
        z = lambda *s: torch.zeros(*s, device=device)
        full = {"embedding.weight": z(4, H), "output_layer.weight": z(4, H), "output_layer.bias": z(4),
                "init_h.weight": z(H, H), "init_h.bias": z(H), "init_c.weight": z(H, H), "init_c.bias": z(H),
                "lstm.weight_ih_l0": z(4 * H, 2 * H), "lstm.weight_hh_l0": z(4 * H, H),
                "lstm.bias_ih_l0": z(4 * H), "lstm.bias_hh_l0": z(4 * H)}
        full.update(sd)
        super().__init__(cfg, full, device)


class SoftAttention(AttentionMechanism):
    """Additive attention, attention.py:38-118."""
    _kind = "soft"

    def __init__(self, config: AttentionConfig):
        super().__init__()
        self.query_dim = self.key_dim = self.hidden_dim = config.hidden_dim
        self.query_proj = nn.Linear(self.query_dim, self.hidden_dim)
        self.key_proj = nn.Linear(self.key_dim, self.hidden_dim)
        self.energy = nn.Linear(self.hidden_dim, 1)
        self.temperature = config.temperature
        self._temperature, self._num_heads = config.temperature, 1


class MultiHeadAttention(AttentionMechanism):
    """Scaled dot-product multi-head attention, attention.py:121-218."""
    _kind = "multi_head"

    def __init__(self, config: AttentionConfig):
        super().__init__()
        self.num_heads = config.num_heads
        self.hidden_dim = config.hidden_dim
        assert self.hidden_dim % self.num_heads == 0, "Hidden dim must be divisible by num heads"
        self.head_dim = self.hidden_dim // self.num_heads
        self.temperature = config.temperature
        self.query_proj = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.key_proj = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.value_proj = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.output_proj = nn.Linear(self.hidden_dim, self.hidden_dim)
        self._temperature, self._num_heads = config.temperature, config.num_heads


def _base(config):
    # attention.py:229-230, 308-309
    return MultiHeadAttention(config) if config.num_heads > 1 else SoftAttention(config)


class AdaptiveAttention(AttentionMechanism):
    """Visual-sentinel adaptive attention, attention.py:221-294."""
    _kind = "adaptive"

    def __init__(self, config: AttentionConfig):
        super().__init__()
        self.hidden_dim = config.hidden_dim
        self.base_attention = _base(config)
        self.sentinel_gate = nn.Linear(self.hidden_dim * 2, self.hidden_dim)
        self.sentinel_proj = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.adaptive_weight = nn.Linear(self.hidden_dim * 2, 1)
        self._temperature, self._num_heads = config.temperature, config.num_heads


class AttentionOnAttention(AttentionMechanism):
    """Attention on Attention, attention.py:297-360."""
    _kind = "aoa"

    def __init__(self, config: AttentionConfig):
        super().__init__()
        self.hidden_dim = config.hidden_dim
        self.base_attention = _base(config)
        self.query_proj = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.info_vector_proj = nn.Sequential(nn.Linear(self.hidden_dim * 2, self.hidden_dim), nn.Tanh())
        self.info_gate_proj = nn.Sequential(nn.Linear(self.hidden_dim * 2, self.hidden_dim), nn.Sigmoid())
        self._temperature, self._num_heads = config.temperature, config.num_heads


def build_attention(config: AttentionConfig) -> AttentionMechanism:
    """Factory, attention.py:363-375 (same ValueError on unknown types)."""
    kind = attention_kind(config)
    if kind == AttentionType.SOFT.value:
        return SoftAttention(config)
    if kind == AttentionType.MULTI_HEAD.value:
        return MultiHeadAttention(config)
    if kind == AttentionType.ADAPTIVE.value:
        return AdaptiveAttention(config)
    if kind == AttentionType.AOA.value:
        return AttentionOnAttention(config)
    raise ValueError(f"Unsupported attention type: {config.attention_type}")
