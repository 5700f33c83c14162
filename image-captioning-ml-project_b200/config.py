"""Configuration surface of the decode path (mirror of /root/reference/src/config.py:7-58,93-124).

Same enum members, dataclass names, field names and defaults as the reference, so callers that build
`DecoderConfig` / `AttentionConfig` objects keep working.  Two reference defects are NOT reproduced
because they make the module unimportable / unusable: nested dataclass defaults use
`default_factory` (src/config.py:114-116 raises on Python >= 3.11), and `AttentionConfig` carries the
`hidden_dim` field every attention class reads (src/models/attention.py:45,130,229,305).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field


class DecoderType(enum.Enum):
    LSTM = "lstm"
    TRANSFORMER = "transformer"
    GPT2 = "gpt2"
    T5 = "t5"
    BART = "bart"


class AttentionType(enum.Enum):
    SOFT = "soft"
    MULTI_HEAD = "multi_head"
    ADAPTIVE = "adaptive"
    AOA = "aoa"
    OBJECT = "object"


@dataclass
class DecoderConfig:
    decoder_type: DecoderType = DecoderType.GPT2
    pretrained_model_name: str = "gpt2"
    hidden_dim: int = 768
    num_layers: int = 6
    num_heads: int = 8
    dropout: float = 0.1
    max_length: int = 50


@dataclass
class AttentionConfig:
    attention_type: AttentionType = AttentionType.MULTI_HEAD
    num_heads: int = 8
    temperature: float = 1.0
    use_geometric: bool = False
    hidden_dim: int = 768


@dataclass
class InferenceConfig:
    decoding_strategy: str = "beam"
    beam_size: int = 5
    top_p: float = 0.9
    temperature: float = 1.0
    min_length: int = 5
    max_length: int = 20
    length_penalty: float = 0.8
    num_beam_groups: int = 1
    diversity_penalty: float = 0.5
    use_clip_reranking: bool = False
    num_candidates: int = 5


@dataclass
class ModelConfig:
    decoder: DecoderConfig = field(default_factory=DecoderConfig)
    attention: AttentionConfig = field(default_factory=AttentionConfig)
    vocab_size: int = 50257
    pad_token_id: int = 0
    bos_token_id: int = 1
    eos_token_id: int = 2


def attention_kind(config) -> str:
    """Accept this package's enum, the reference's enum (same .value) or a plain string."""
    t = getattr(config, "attention_type", config)
    return t.value if hasattr(t, "value") else str(t)


def decoder_kind(config) -> str:
    t = getattr(config, "decoder_type", config)
    return t.value if hasattr(t, "value") else str(t)
