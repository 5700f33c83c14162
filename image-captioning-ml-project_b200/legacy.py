"""Drop-in for the legacy Show-Attend-Tell decoder (mirror of /root/reference/models/decoder.py).

`Decoder(vocab_size, use_bert, device)` keeps the reference's constructor, parameter names and the
teacher-forced `forward(encoder_out, encoded_captions, caption_lengths) -> (predictions,
encoded_captions, dec_len, alphas)`; the step arithmetic (models/decoder.py:148-173) runs in libcapdec.
The free-running entry points the reference only plans (docs/technical_architecture.md:217) are added on
the same step kernels: `beam_search`, `greedy`, `sample`.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _capi
from .engine import Engine, EngineOwner, Tiles

PAD, START, END, UNK = 0, 1, 2, 3   # models/constants.py



class Decoder(EngineOwner, nn.Module):
    def __init__(self, vocab_size, use_bert=False, device="cuda", precision: str = "fp32"):
        super().__init__()
        if use_bert:
            raise NotImplementedError("use_bert=True (BERT caption embeddings, models/decoder.py:60-118) is a "
                                      "training-only text path outside the decode hot path")
        self.encoder_dim, self.attention_dim, self.embed_dim, self.decoder_dim = 2048, 512, 512, 512
        self.use_bert, self.device, self.vocab_size = use_bert, device, vocab_size
        self.precision = precision
        # same construction order as models/decoder.py:33-55 => same init under the same seed
        self.enc_att = nn.Linear(2048, 512)
        self.dec_att = nn.Linear(512, 512)
        self.att = nn.Linear(512, 1)
        self.relu = nn.ReLU()
        self.softmax = nn.Softmax(dim=1)
        self.dropout = nn.Dropout(p=0.5)
        self.decode_step = nn.LSTMCell(self.embed_dim + self.encoder_dim, self.decoder_dim, bias=True)
        self.h_lin = nn.Linear(self.encoder_dim, self.decoder_dim)
        self.c_lin = nn.Linear(self.encoder_dim, self.decoder_dim)
        self.f_beta = nn.Linear(self.decoder_dim, self.encoder_dim)
        self.sigmoid = nn.Sigmoid()
        self.fc = nn.Linear(self.decoder_dim, self.vocab_size)
        self.fc.bias.data.fill_(0)
        self.fc.weight.data.uniform_(-0.1, 0.1)
        self.embedding = nn.Embedding(vocab_size, self.embed_dim)
        self.embedding.weight.data.uniform_(-0.1, 0.1)

    def _engine(self, device) -> Engine:
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (str(device), self.precision)
        if getattr(self, "_eng_sig", None) != sig:
            cfg = _capi.Config(arch=_capi.ARCH_LEGACY_SAT, attention=_capi.ATT["soft"],
                               precision=_capi.PREC[self.precision], vocab_size=self.vocab_size,
                               hidden_dim=self.decoder_dim, embed_dim=self.embed_dim, feature_dim=self.encoder_dim,
                               attention_dim=self.attention_dim, num_layers=1, num_heads=1, temperature=1.0,
                               pad_token_id=PAD, bos_token_id=START, eos_token_id=END)
            object.__setattr__(self, "_eng", Engine(cfg, self.state_dict(), device))
            object.__setattr__(self, "_eng_sig", sig)
        return self._eng

    @staticmethod
    def _regions(encoder_out):
        if isinstance(encoder_out, Tiles):
            return encoder_out
        return encoder_out.reshape(encoder_out.size(0), -1, encoder_out.size(-1))   # models/decoder.py:127

    @torch.no_grad()
    def ingest(self, trunk_out: torch.Tensor, layout: str = "nchw") -> Tiles:
        """Encoder hand-off: the resnet trunk's [B,2048,14,14] output (models/encoder.py:13, BEFORE its permute; fp32 or the
        bf16/fp16 of an autocast encoder) -> the tile set `beam_search` streams.  One device pass does the permute, the
        region mean and the operand packing the decode prologue would otherwise redo from fp32 [B,14,14,2048]."""
        if layout == "bld" and trunk_out.dim() == 4:
            trunk_out = trunk_out.reshape(trunk_out.size(0), -1, trunk_out.size(-1))
        return self._engine(trunk_out.device).ingest_features(trunk_out, layout)

    def forward(self, encoder_out, encoded_captions, caption_lengths):
        """models/decoder.py:120-176 (eval semantics: dropout is the identity)."""
        dec_len = [int(x) - 1 for x in caption_lengths]
        enc = self._regions(encoder_out)
        preds, alphas = self._engine(enc.device).forward_teacher(enc, encoded_captions, dec_len)
        return preds, encoded_captions, dec_len, alphas

    @torch.no_grad()
    def beam_search(self, encoder_out, beam_size: int = 5, max_length: int = 20, length_penalty: float = 1.0,
                    trace: bool = False, crop: bool = True):
        """-> {"tokens" int32 [B,T], "lengths" int32 [B], "scores" float [B]} on the device (asynchronous), plus
        "sequences" = the int64 [B, longest] HF-style view when `crop` (cropping needs max(lengths) on the host, i.e.
        one synchronisation; batch pipelines pass crop=False and trim with `lengths`).  `encoder_out` is the encoder's
        [B,14,14,2048] output or a tile set from `ingest`."""
        enc = self._regions(encoder_out)
        out = self._engine(enc.device).decode_beam(enc, None, None, beam_size, max_length, length_penalty, trace=trace)
        if crop:
            out["sequences"] = out["tokens"].long()[:, : int(out["lengths"].max().item())]
        return out

    @torch.no_grad()
    def score_tokens(self, encoder_out, tokens, rows_per_image: int = 1) -> torch.Tensor:
        """log p(tokens[:, t+1] | tokens[:, :t+1], image), one teacher-forced pass (the SCST re-scoring of sampled
        captions, src/train/trainer.py:366-378, on the legacy step models/decoder.py:148-173)."""
        enc = self._regions(encoder_out)
        _, lp, _ = self._engine(enc.device).forward_tokens(enc, None, None, tokens, rows_per_image, want_logits=False,
                                                           want_logprob=True)
        return lp

    @torch.no_grad()
    def greedy(self, encoder_out, max_length: int = 20, start_token_id: int = START):
        enc = self._regions(encoder_out)
        tok, alpha = self._engine(enc.device).decode_greedy(enc, None, None, max_length, start_token_id)
        return tok.long(), alpha

    @torch.no_grad()
    def sample(self, encoder_out, num_samples: int = 1, with_greedy: bool = False, max_length: int = 20,
               uniforms: Optional[torch.Tensor] = None):
        enc = self._regions(encoder_out)
        k = num_samples + (1 if with_greedy else 0)
        if uniforms is None:
            uniforms = torch.rand(enc.shape[0] * k, max_length - 1, device=enc.device)
        tok, lp = self._engine(enc.device).decode_sample(enc, None, None, num_samples, with_greedy, max_length, uniforms)
        return tok.long(), lp
