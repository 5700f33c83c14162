"""Drop-in caption decoders (mirror of /root/reference/src/models/decoders.py, LSTM family).

`LSTMDecoder` keeps the reference's constructor signature, parameter names (state_dict-compatible
with src/models/decoders.py:92-117) and `generate(encoder_features, max_length, start_token_id=1,
**kwargs) -> (LongTensor, info)` contract; `generate` runs in libcapdec.  Extra keyword arguments
select the decode entry points the north star adds on the same step kernels:

    num_beams=k            HF-static beam search (the algorithm GPT2Decoder.generate reaches at
                           decoders.py:645); returns HF-cropped sequences, info["scores"], info["lengths"]
    do_sample=True         ancestral sampling rollout (trainer.py:383-438): num_samples rows per image
                           (+ with_greedy=True for the SCST baseline row), uniforms=... to fix the draws

`forward(encoder_features, captions=ids)` is the teacher-forced pass as ONE batched call of the step kernels with
forced tokens; it returns {"logits": [B,t,V], ...} like the reference, so CaptioningTrainer._sample_captions
(src/train/trainer.py:413-420) runs unmodified on these classes.  It is inference-only: no autograd graph is built and
dropout is the identity (the REINFORCE gradient stays with the PyTorch module); `score_tokens`
returns the per-token log-probs of given captions in the same single pass (trainer.py:366-378's sample_logprobs).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _capi
from .attention import build_attention
from .config import AttentionConfig, DecoderConfig, DecoderType, attention_kind, decoder_kind
from .engine import Engine, EngineOwner, Tiles


class CaptionDecoder(EngineOwner, nn.Module, ABC):
    """Base class for all caption decoders (decoders.py:20-69)."""

    @abstractmethod
    def forward(self, encoder_features, captions=None, caption_lengths=None, **kwargs) -> Dict[str, torch.Tensor]:
        ...

    @abstractmethod
    def generate(self, encoder_features, max_length: int, **kwargs) -> Tuple[torch.Tensor, Dict[str, Any]]:
        ...


def _reject_unknown(kwargs, where):
    """The reference forwards **kwargs to the generation backend (decoders.py:640-650).  Anything this path does not
    implement must fail loudly instead of silently decoding with defaults."""
    if kwargs:
        raise TypeError(f"{where}: unsupported generation argument(s) {sorted(kwargs)}; supported: num_beams, "
                        "length_penalty, do_sample, num_samples, with_greedy, uniforms, trace, return_scores")


def _padding_mask(encoder_features):
    """decoders.py:291 passes key_padding_mask = ~attention_mask.  The reference encoders emit a float
    mask (encoders.py:84) on which `~` raises; here any dtype is accepted, non-zero = valid region."""
    m = encoder_features.get("attention_mask", None)
    if m is None:
        return None
    return ~(m.bool())


class LSTMDecoder(CaptionDecoder):
    """LSTM decoder with attention (decoders.py:72-314)."""

    def __init__(self, config: DecoderConfig, attention_config: AttentionConfig, vocab_size: int, pad_token_id: int,
                 embedding_dim: int = None, bos_token_id: int = 1, eos_token_id: int = 2, precision: str = "fp32"):
        super().__init__()
        self.hidden_dim = config.hidden_dim
        self.embedding_dim = embedding_dim or config.hidden_dim
        self.num_layers = config.num_layers
        self.vocab_size = vocab_size
        self.dropout_p = config.dropout
        self.pad_token_id = pad_token_id
        self.bos_token_id, self.eos_token_id = bos_token_id, eos_token_id
        self.max_length = config.max_length
        self.precision = precision
        self._attention_kind = attention_kind(attention_config)
        self._attention_heads = attention_config.num_heads
        self._attention_temperature = attention_config.temperature
        # same construction order as the reference => identical random init under the same seed
        self.embedding = nn.Embedding(vocab_size, self.embedding_dim, padding_idx=pad_token_id)
        self.lstm = nn.LSTM(input_size=self.embedding_dim + self.hidden_dim, hidden_size=self.hidden_dim,
                            num_layers=self.num_layers, batch_first=True,
                            dropout=self.dropout_p if self.num_layers > 1 else 0)
        self.attention = build_attention(attention_config)
        self.output_layer = nn.Linear(self.hidden_dim, vocab_size)
        self.init_h = nn.Linear(self.hidden_dim, self.hidden_dim * self.num_layers)
        self.init_c = nn.Linear(self.hidden_dim, self.hidden_dim * self.num_layers)
        self.dropout = nn.Dropout(self.dropout_p)

    # ---- engine lifecycle: re-bind when parameters, device or precision change
    def _engine(self, device) -> Engine:
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (str(device), self.precision)
        if getattr(self, "_eng_sig", None) != sig:
            H = self.hidden_dim
            cfg = _capi.Config(arch=_capi.ARCH_LSTM, attention=_capi.ATT[self._attention_kind],
                               precision=_capi.PREC[self.precision], vocab_size=self.vocab_size, hidden_dim=H,
                               embed_dim=self.embedding_dim, feature_dim=H, attention_dim=H,
                               num_layers=self.num_layers, num_heads=int(self._attention_heads),
                               temperature=float(self._attention_temperature), pad_token_id=int(self.pad_token_id),
                               bos_token_id=int(self.bos_token_id), eos_token_id=int(self.eos_token_id))
            object.__setattr__(self, "_eng", Engine(cfg, self.state_dict(), device))
            object.__setattr__(self, "_eng_sig", sig)
        return self._eng

    def forward(self, encoder_features, captions=None, caption_lengths=None, **kwargs):
        """decoders.py:137-234: teacher forcing, step t consumes captions[:, t] and the previous context; logits[:, t] =
        output_layer(context_t).  `caption_lengths` is ignored: the reference sorts captions and features by it but
        takes h0 / c0 from the unsorted pooled features (:157-171), pairing rows with another image's initial state;
        its only caller on this path, _sample_captions, passes None."""
        if captions is None:
            return self.generate(encoder_features, self.max_length)   # reference hits a NameError here (:148)
        feats = encoder_features["features"]
        eng = self._engine(feats.device)
        logits, _, alpha = eng.forward_tokens(feats, encoder_features["pooled_features"], _padding_mask(encoder_features),
                                              captions, want_alpha=True)
        return {"logits": logits, "attention_weights": alpha}

    def score_tokens(self, encoder_features, tokens, rows_per_image: int = 1) -> torch.Tensor:
        """log p(tokens[:, t+1] | tokens[:, :t+1], image) for every t, one pass; tokens [B*rows_per_image, T]."""
        feats = encoder_features["features"]
        _, lp, _ = self._engine(feats.device).forward_tokens(feats, encoder_features["pooled_features"],
                                                             _padding_mask(encoder_features), tokens, rows_per_image,
                                                             want_logits=False, want_logprob=True)
        return lp

    def generate(self, encoder_features: Dict[str, torch.Tensor], max_length: int, start_token_id: int = 1,
                 num_beams: int = 1, do_sample: bool = False, num_samples: int = 1, with_greedy: bool = False,
                 uniforms: Optional[torch.Tensor] = None, length_penalty: float = 1.0, trace: bool = False,
                 return_scores: bool = True, **kwargs) -> Tuple[torch.Tensor, Dict[str, Any]]:
        _reject_unknown(kwargs, "LSTMDecoder.generate")
        feats = encoder_features["features"]
        pooled = encoder_features["pooled_features"]
        mask = _padding_mask(encoder_features)
        if (do_sample or num_beams > 1) and int(start_token_id) != int(self.bos_token_id):
            raise ValueError(f"beam search / sampling start from bos_token_id={self.bos_token_id}; "
                             f"start_token_id={start_token_id} is honoured by the greedy path only (decoders.py:240)")
        eng = self._engine(feats.device)
        if do_sample:
            B = feats.shape[0]
            k = num_samples + (1 if with_greedy else 0)
            if uniforms is None:
                uniforms = torch.rand(B * k, max_length - 1, device=feats.device)
            tok, lp = eng.decode_sample(feats, pooled, mask, num_samples, with_greedy, max_length, uniforms)
            return tok.long(), {"log_probs": lp}
        if num_beams > 1:
            out = eng.decode_beam(feats, pooled, mask, num_beams, max_length, length_penalty, trace=trace)
            seq = out["tokens"].long()
            seq = seq[:, : int(out["lengths"].max().item())]     # HF crops to the longest hypothesis
            info = {"scores": out["scores"], "lengths": out["lengths"].long()}
            if trace:
                info.update({k: out[k] for k in ("top_logprob", "top_token", "top_beam")})
            return seq, info
        tok, alpha = eng.decode_greedy(feats, pooled, mask, max_length, start_token_id)
        return tok.long(), {"attention_weights": alpha}


class TransformerDecoder(CaptionDecoder):
    """Transformer decoder over visual features (decoders.py:317-493), decoded with a self-attention KV cache and
    hoisted cross-attention K/V (mathematically identical to the reference's full-prefix recompute: the stack is
    causal).  Same constructor, parameter names and `generate` contract as the reference."""

    def __init__(self, config: DecoderConfig, vocab_size: int, pad_token_id: int, bos_token_id: int, eos_token_id: int,
                 precision: str = "fp32"):
        super().__init__()
        self.hidden_dim = config.hidden_dim
        self.num_layers = config.num_layers
        self.num_heads = config.num_heads
        self.vocab_size = vocab_size
        self.dropout_p = config.dropout
        self.pad_token_id, self.bos_token_id, self.eos_token_id = pad_token_id, bos_token_id, eos_token_id
        self.precision = precision
        self.embedding = nn.Embedding(vocab_size, self.hidden_dim, padding_idx=pad_token_id)
        self.position_encoding = nn.Embedding(config.max_length, self.hidden_dim)
        decoder_layer = nn.TransformerDecoderLayer(d_model=self.hidden_dim, nhead=self.num_heads,
                                                   dim_feedforward=self.hidden_dim * 4, dropout=self.dropout_p,
                                                   activation="gelu", batch_first=True)
        self.transformer_decoder = nn.TransformerDecoder(decoder_layer=decoder_layer, num_layers=self.num_layers)
        self.output_layer = nn.Linear(self.hidden_dim, vocab_size)
        self.visual_projection = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.dropout = nn.Dropout(self.dropout_p)

    def _engine(self, device) -> Engine:
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (str(device), self.precision)
        if getattr(self, "_eng_sig", None) != sig:
            H = self.hidden_dim
            cfg = _capi.Config(arch=_capi.ARCH_TRANSFORMER, attention=_capi.ATT["multi_head"],
                               precision=_capi.PREC[self.precision], vocab_size=self.vocab_size, hidden_dim=H,
                               embed_dim=H, feature_dim=H, attention_dim=H, num_layers=self.num_layers,
                               num_heads=int(self.num_heads), temperature=1.0, pad_token_id=int(self.pad_token_id),
                               bos_token_id=int(self.bos_token_id), eos_token_id=int(self.eos_token_id))
            object.__setattr__(self, "_eng", Engine(cfg, self.state_dict(), device))
            object.__setattr__(self, "_eng_sig", sig)
        return self._eng

    def forward(self, encoder_features, captions=None, caption_lengths=None, **kwargs):
        """decoders.py:377-438: full-prefix teacher forcing under the causal mask == the KV-cached step applied to the
        forced tokens position by position.  Pad tokens inside `captions` are masked as keys (tgt_key_padding_mask, :405);
        an `attention_mask` (True = valid region) masks region keys with -1e9 as the reference intends (:393-398; as shipped
        its [B,T,L] mask only has a valid shape for T == 1)."""
        if captions is None:
            return self.generate(encoder_features, 50)      # decoders.py:385-387
        feats = encoder_features["features"]
        if captions.shape[1] > self.position_encoding.num_embeddings:
            raise IndexError("caption length exceeds the position_encoding table (index out of range in self)")
        logits, _, _ = self._engine(feats.device).forward_tokens(feats, None, _padding_mask(encoder_features), captions)
        return {"logits": logits}          # ("hidden_states" of the reference are internal to the fused step)

    def score_tokens(self, encoder_features, tokens, rows_per_image: int = 1) -> torch.Tensor:
        feats = encoder_features["features"]
        _, lp, _ = self._engine(feats.device).forward_tokens(feats, None, _padding_mask(encoder_features), tokens,
                                                             rows_per_image, want_logits=False, want_logprob=True)
        return lp

    def generate(self, encoder_features: Dict[str, torch.Tensor], max_length: int, num_beams: int = 1,
                 do_sample: bool = False, num_samples: int = 1, with_greedy: bool = False,
                 uniforms: Optional[torch.Tensor] = None, length_penalty: float = 1.0, trace: bool = False,
                 return_scores: bool = True, memory_key_padding_mask: Optional[torch.Tensor] = None,
                 **kwargs) -> Tuple[torch.Tensor, Dict[str, Any]]:
        """decoders.py:439-493.  The reference's generate never looks at encoder_features["attention_mask"]; padded
        region sets (Q-Former / object-region encoders) pass `memory_key_padding_mask` [B,L] (True = padding) explicitly."""
        _reject_unknown(kwargs, "TransformerDecoder.generate")
        feats = encoder_features["features"]
        mkpm = memory_key_padding_mask
        if max_length > self.position_encoding.num_embeddings:
            raise IndexError("max_length exceeds the position_encoding table (index out of range in self)")
        eng = self._engine(feats.device)
        if do_sample:
            B = feats.shape[0]
            k = num_samples + (1 if with_greedy else 0)
            if uniforms is None:
                uniforms = torch.rand(B * k, max_length - 1, device=feats.device)
            tok, lp = eng.decode_sample(feats, None, mkpm, num_samples, with_greedy, max_length, uniforms)
            tok = tok.long()
            n = _all_eos_cut(tok, self.eos_token_id)          # trainer.py:435 batch-wide break
            return tok[:, :n], {"log_probs": lp[:, : n - 1]}
        if num_beams > 1:
            out = eng.decode_beam(feats, None, mkpm, num_beams, max_length, length_penalty, trace=trace)
            seq = out["tokens"].long()[:, : int(out["lengths"].max().item())]
            info = {"scores": out["scores"], "lengths": out["lengths"].long()}
            if trace:
                info.update({k: out[k] for k in ("top_logprob", "top_token", "top_beam")})
            return seq, info
        tok, _ = eng.decode_greedy(feats, None, mkpm, max_length, self.bos_token_id, want_alpha=False)
        tok = tok.long()
        return tok[:, : _all_eos_cut(tok, self.eos_token_id)], {}


class GPT2Decoder(CaptionDecoder):
    """GPT-2 decoder conditioned on a 10-token image prefix (decoders.py:496-656).

    Same constructor, attributes and parameter names as the reference (`self.model` IS a transformers
    GPT2LMHeadModel, so a reference state_dict loads unchanged); `generate(encoder_features, max_length,
    num_beams=4)` returns HF-style beam-search sequences.  As shipped the reference passes a list of
    [B,prefix,hidden] tensors as past_key_values, which no transformers version accepts; this
    drop-in implements the intended computation pinned in SURVEY section 8(c): the prefix, split into heads, is the past
    key AND value of every layer, positions continue after the prefix, HF beam search with its defaults."""

    def __init__(self, config: DecoderConfig, vocab_size: int = None, pad_token_id: int = None, bos_token_id: int = None,
                 eos_token_id: int = None, precision: str = "fp32"):
        super().__init__()
        from transformers import GPT2Config, GPT2LMHeadModel
        if config.pretrained_model_name:
            self.model = GPT2LMHeadModel.from_pretrained(config.pretrained_model_name)
            if vocab_size and vocab_size != self.model.config.vocab_size:
                self.model.resize_token_embeddings(vocab_size)
        else:
            self.model = GPT2LMHeadModel(GPT2Config(
                vocab_size=vocab_size, n_positions=config.max_length, n_ctx=config.max_length, n_embd=config.hidden_dim,
                n_layer=config.num_layers, n_head=config.num_heads, resid_pdrop=config.dropout,
                embd_pdrop=config.dropout, attn_pdrop=config.dropout))
        self.pad_token_id = pad_token_id or 0
        self.bos_token_id = bos_token_id or 1
        self.eos_token_id = eos_token_id or 2
        self.precision = precision
        self.visual_projection = nn.Linear(config.hidden_dim, self.model.config.n_embd)
        self.prefix_length = 10
        self.image_prefix = nn.Parameter(torch.randn(1, self.prefix_length, self.model.config.n_embd))
        self.image_to_prefix = nn.Linear(config.hidden_dim, self.prefix_length * self.model.config.n_embd)
        self._feature_dim = config.hidden_dim

    def _engine(self, device) -> Engine:
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (str(device), self.precision)
        if getattr(self, "_eng_sig", None) != sig:
            gc = self.model.config
            cfg = _capi.Config(arch=_capi.ARCH_GPT2, attention=_capi.ATT["multi_head"], precision=_capi.PREC[self.precision],
                               vocab_size=gc.vocab_size, hidden_dim=gc.n_embd, embed_dim=gc.n_embd,
                               feature_dim=self._feature_dim, attention_dim=gc.n_embd, num_layers=gc.n_layer,
                               num_heads=gc.n_head, temperature=1.0, pad_token_id=int(self.pad_token_id),
                               bos_token_id=int(self.bos_token_id), eos_token_id=int(self.eos_token_id))
            sd = {}
            for name, t in self.state_dict().items():
                if name.startswith("model.transformer.h.") and name.endswith(
                        ("attn.c_attn.weight", "attn.c_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight")):
                    t = t.t()                       # HF Conv1D stores [in, out]; libcapdec takes nn.Linear layout
                if name.endswith((".attn.bias", ".attn.masked_bias")) or name in ("image_prefix",) or name.startswith("visual_projection"):
                    continue                        # causal-mask buffers / parameters generate() never reads
                sd[name] = t
            object.__setattr__(self, "_eng", Engine(cfg, sd, device))
            object.__setattr__(self, "_eng_sig", sig)
        return self._eng

    def forward(self, encoder_features, captions=None, caption_lengths=None, **kwargs):
        """decoders.py:563-596 with the prefix bound as SURVEY section 8(c) pins it: logits of GPT2LMHeadModel over
        `captions` behind the 10-token image prefix (past K == V of every layer), pad tokens masked as keys
        (attention_mask = captions != pad, :581), and HF's shifted cross-entropy against labels = captions (:578,:589)."""
        if captions is None:
            return self.generate(encoder_features, 50)      # decoders.py:572-574
        pooled = encoder_features["pooled_features"]
        if self.prefix_length + captions.shape[1] > self.model.config.n_positions:
            raise IndexError("prefix + caption length exceeds GPT-2 n_positions (index out of range in self)")
        logits, lp, _ = self._engine(pooled.device).forward_tokens(None, pooled, None, captions, want_logprob=True)
        out = {"logits": logits}
        if lp.numel():
            out["loss"] = -lp.mean()
        return out

    def score_tokens(self, encoder_features, tokens, rows_per_image: int = 1) -> torch.Tensor:
        pooled = encoder_features["pooled_features"]
        _, lp, _ = self._engine(pooled.device).forward_tokens(None, pooled, None, tokens, rows_per_image,
                                                              want_logits=False, want_logprob=True)
        return lp

    # HF generate() arguments whose DEFAULT value is what this path implements; any other value, or any other argument,
    # raises (the reference forwards **kwargs to transformers' generate, decoders.py:640-650)
    _HF_DEFAULTS = {"early_stopping": False, "repetition_penalty": 1.0, "no_repeat_ngram_size": 0,
                    "num_return_sequences": 1, "min_length": 0, "temperature": 1.0, "top_k": 50, "top_p": 1.0,
                    "num_beam_groups": 1, "diversity_penalty": 0.0, "use_cache": True}

    def generate(self, encoder_features: Dict[str, torch.Tensor], max_length: int, num_beams: int = 4,
                 do_sample: bool = False, num_samples: int = 1, with_greedy: bool = False,
                 uniforms: Optional[torch.Tensor] = None, length_penalty: float = 1.0, trace: bool = False,
                 return_scores: bool = False, **kwargs) -> Tuple[torch.Tensor, Dict[str, Any]]:
        for key in list(kwargs):
            if key in self._HF_DEFAULTS and kwargs[key] == self._HF_DEFAULTS[key]:
                kwargs.pop(key)
            elif key in self._HF_DEFAULTS:
                raise NotImplementedError(f"GPT2Decoder.generate: {key}={kwargs[key]!r} is not implemented by the CUDA beam "
                                          f"search (only the HF default {self._HF_DEFAULTS[key]!r})")
        _reject_unknown(kwargs, "GPT2Decoder.generate")
        pooled = encoder_features["pooled_features"]
        B = pooled.shape[0]
        if self.prefix_length + max_length > self.model.config.n_positions:
            raise IndexError("prefix + max_length exceeds GPT-2 n_positions (index out of range in self)")
        eng = self._engine(pooled.device)
        dummy = torch.empty(B, 1, 4, device=pooled.device)          # region features are not used by this decoder
        if do_sample:
            k = num_samples + (1 if with_greedy else 0)
            if uniforms is None:
                uniforms = torch.rand(B * k, max_length - 1, device=pooled.device)
            tok, lp = eng.decode_sample(dummy, pooled, None, num_samples, with_greedy, max_length, uniforms)
            tok = tok.long()
            n = _all_eos_cut(tok, self.eos_token_id)
            return tok[:, :n], {"log_probs": lp[:, : n - 1]}
        if num_beams > 1:
            out = eng.decode_beam(dummy, pooled, None, num_beams, max_length, length_penalty, trace=trace)
            seq = out["tokens"].long()[:, : int(out["lengths"].max().item())]
            info = {"scores": out["scores"], "lengths": out["lengths"].long()} if trace or return_scores else {}
            if trace:
                info.update({k: out[k] for k in ("top_logprob", "top_token", "top_beam")})
            return seq, info
        # num_beams == 1: HF greedy search -- rows stop at their own EOS and are padded, the loop ends when all did
        tok, _ = eng.decode_greedy(dummy, pooled, None, max_length, self.bos_token_id, want_alpha=False)
        tok = tok.long()
        done = (tok[:, 1:] == self.eos_token_id).cumsum(dim=1) > 0
        after = torch.cat([torch.zeros_like(done[:, :1]), done[:, :-1]], dim=1)      # strictly after the first EOS
        tok[:, 1:][after] = self.pad_token_id
        all_done = done.all(dim=0)
        idx = torch.nonzero(all_done).flatten()
        n = int(idx[0].item()) + 2 if idx.numel() else tok.shape[1]
        return tok[:, :n], {}


def _all_eos_cut(tok: torch.Tensor, eos: int) -> int:
    """The reference stops only when EVERY row emits EOS at the same step (decoders.py:490, trainer.py:435);
    returns the number of columns it would have produced."""
    hit = (tok[:, 1:] == eos).all(dim=0)
    idx = torch.nonzero(hit).flatten()
    return int(idx[0].item()) + 2 if idx.numel() else tok.shape[1]


def build_decoder(config: DecoderConfig, attention_config: AttentionConfig, vocab_size: int, pad_token_id: int,
                  bos_token_id: int, eos_token_id: int) -> CaptionDecoder:
    """Factory (decoders.py:659-692).  Only the LSTM family is accelerated so far; other types raise."""
    kind = decoder_kind(config)
    if kind == DecoderType.LSTM.value:
        return LSTMDecoder(config=config, attention_config=attention_config, vocab_size=vocab_size,
                           pad_token_id=pad_token_id, bos_token_id=bos_token_id, eos_token_id=eos_token_id)
    if kind == DecoderType.TRANSFORMER.value:
        return TransformerDecoder(config=config, vocab_size=vocab_size, pad_token_id=pad_token_id,
                                  bos_token_id=bos_token_id, eos_token_id=eos_token_id)
    if kind == DecoderType.GPT2.value:
        return GPT2Decoder(config=config, vocab_size=vocab_size, pad_token_id=pad_token_id, bos_token_id=bos_token_id,
                           eos_token_id=eos_token_id)
    raise ValueError(f"Unsupported decoder type: {config.decoder_type}")
