"""capdec-b200: B200-native batched caption decoding behind the reference's decoder API.

Importing this package loads image-captioning-ml-project_b200/csrc/libcapdec.so and fails if it is
missing -- there is no fallback implementation.
"""
from . import _capi                                   # noqa: F401  (loads the shared library)
from .config import AttentionConfig, AttentionType, DecoderConfig, DecoderType, InferenceConfig, ModelConfig  # noqa: F401
from .attention import (AttentionMechanism, AttentionOnAttention, AdaptiveAttention, MultiHeadAttention,  # noqa: F401
                        SoftAttention, build_attention)
from .decoders import CaptionDecoder, GPT2Decoder, LSTMDecoder, TransformerDecoder, build_decoder  # noqa: F401
from .legacy import Decoder  # noqa: F401
from .engine import Engine, launch_count  # noqa: F401
from .postprocess import coco_results, decode_captions, to_token_lists, trim_at_eos, write_results_json  # noqa: F401

__all__ = ["AttentionConfig", "AttentionType", "DecoderConfig", "DecoderType", "InferenceConfig", "ModelConfig",
           "AttentionMechanism", "SoftAttention", "MultiHeadAttention", "AdaptiveAttention", "AttentionOnAttention",
           "build_attention", "CaptionDecoder", "LSTMDecoder", "TransformerDecoder", "GPT2Decoder", "build_decoder", "Decoder", "Engine", "launch_count",
           "trim_at_eos", "to_token_lists", "decode_captions", "coco_results", "write_results_json"]
