"""Import shims for the UNMODIFIED reference modules (test infrastructure only).

The reference does not import as shipped on Python >= 3.11:
  * src/config.py:114-116,129-131 use dataclass instances as field defaults -> ValueError
  * AttentionConfig (src/config.py:52-58) lacks `hidden_dim`, which every attention class
    reads (src/models/attention.py:45,130,229,305)
  * models/decoder.py:2 imports pytorch_pretrained_bert (absent; unused when use_bert=False)

`load_reference()` registers a replacement `src.config` (same enums / dataclass field names and
defaults, plus AttentionConfig.hidden_dim) and a stub `pytorch_pretrained_bert` in sys.modules and
then imports the reference files from /root/reference WITHOUT editing them.  It returns a namespace
with the reference classes.  Nothing here exists on the GPU box (no /root/reference there):
callers must check `reference_available()` first.
"""
from __future__ import annotations

import enum
import importlib
import os
import sys
import types
from dataclasses import dataclass, field

REFERENCE_ROOT = os.environ.get("CAPDEC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models", "decoders.py"))


def _make_config_module() -> types.ModuleType:
    """Replacement for src/config.py (field names and defaults follow src/config.py:7-152)."""
    m = types.ModuleType("src.config")

    class EncoderType(enum.Enum):
        RESNET = "resnet"; VIT = "vit"; SWIN = "swin"; CONVNEXT = "convnext"
        EFFICIENTNET = "efficientnet"; CLIP = "clip"

    class DecoderType(enum.Enum):
        LSTM = "lstm"; TRANSFORMER = "transformer"; GPT2 = "gpt2"; T5 = "t5"; BART = "bart"

    class AttentionType(enum.Enum):
        SOFT = "soft"; MULTI_HEAD = "multi_head"; ADAPTIVE = "adaptive"; AOA = "aoa"; OBJECT = "object"

    @dataclass
    class EncoderConfig:
        encoder_type: EncoderType = EncoderType.VIT
        pretrained_model_name: str = "google/vit-base-patch16-224"
        freeze: bool = False
        feature_dim: int = 768
        use_object_features: bool = False

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
        hidden_dim: int = 768  # read by attention.py:45,130,229,305 but missing from the shipped dataclass

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
        encoder: EncoderConfig = field(default_factory=EncoderConfig)
        decoder: DecoderConfig = field(default_factory=DecoderConfig)
        attention: AttentionConfig = field(default_factory=AttentionConfig)
        projection_dim: int = 768
        use_q_former: bool = False
        q_former_num_queries: int = 32
        vocab_size: int = 50257
        pad_token_id: int = 0
        bos_token_id: int = 1
        eos_token_id: int = 2

    @dataclass
    class TrainingConfig:          # src/config.py:61-90 (only constructed, never read by the hot path)
        batch_size: int = 64
        num_epochs: int = 30
        learning_rate: float = 5e-5
        use_rl: bool = False
        rl_start_epoch: int = 20

    @dataclass
    class Config:                  # src/config.py:127-152
        model: ModelConfig = field(default_factory=ModelConfig)
        training: TrainingConfig = field(default_factory=TrainingConfig)
        inference: InferenceConfig = field(default_factory=InferenceConfig)
        device: str = "cpu"

    for k, v in dict(locals()).items():
        if k != "m":
            setattr(m, k, v)
    return m


_cached = None


def load_reference():
    """Return a namespace holding the reference's own classes (imported, not copied)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    # --- modern package: src.config shim + namespace packages rooted in the reference tree
    if "src" not in sys.modules or not getattr(sys.modules["src"], "__capdec_shim__", False):
        src_pkg = types.ModuleType("src")
        src_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "src")]
        src_pkg.__capdec_shim__ = True
        sys.modules["src"] = src_pkg
        models_pkg = types.ModuleType("src.models")
        models_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "src", "models")]
        sys.modules["src.models"] = models_pkg
        sys.modules["src.config"] = _make_config_module()
    attention = importlib.import_module("src.models.attention")
    decoders = importlib.import_module("src.models.decoders")
    config = sys.modules["src.config"]

    # --- legacy decoder: stub BERT dependency; `from constants import *` needs models/ on sys.path
    if "pytorch_pretrained_bert" not in sys.modules:
        stub = types.ModuleType("pytorch_pretrained_bert")
        stub.BertTokenizer = object
        stub.BertModel = object
        sys.modules["pytorch_pretrained_bert"] = stub
    legacy_dir = os.path.join(REFERENCE_ROOT, "models")
    if legacy_dir not in sys.path:
        sys.path.insert(0, legacy_dir)
    spec = importlib.util.spec_from_file_location("capdec_ref_legacy_decoder",
                                                  os.path.join(legacy_dir, "decoder.py"))
    legacy = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(legacy)

    ns = types.SimpleNamespace(config=config, attention=attention, decoders=decoders, legacy=legacy)
    _cached = ns
    return ns


_cached_trainer = None


def load_reference_trainer():
    """The reference's trainer / facade modules, imported unmodified: CaptioningTrainer (src/train/trainer.py, for its
    `_sample_captions` :383-438), QFormer (src/models/captioning_model.py:153-245) and ObjectRegionEncoder
    (src/models/encoders.py:233-296)."""
    global _cached_trainer
    if _cached_trainer is not None:
        return _cached_trainer
    load_reference()
    for name in ("train", "evaluate", "data"):
        key = "src." + name
        if key not in sys.modules:
            pkg = types.ModuleType(key)
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, "src", name)]
            sys.modules[key] = pkg
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):      # metrics.py prints a warning when pycocoevalcap is absent
        trainer = importlib.import_module("src.train.trainer")
    cm = importlib.import_module("src.models.captioning_model")
    enc = importlib.import_module("src.models.encoders")
    _cached_trainer = types.SimpleNamespace(trainer=trainer, CaptioningTrainer=trainer.CaptioningTrainer, QFormer=cm.QFormer,
                                            ObjectRegionEncoder=enc.ObjectRegionEncoder, captioning_model=cm, encoders=enc)
    return _cached_trainer
