"""Shared builders for tests, golden generation, smoke and bench: seeded weights and synthetic inputs.

Weights come from constructing the drop-in modules under torch.manual_seed(seed); the drop-ins create
their layers in the reference's order, so the result equals the reference module's own init under the
same seed (asserted against the real reference in tests/test_oracle_pin.py).  Inputs follow
SURVEY.md section 8(d): features = relu(randn(B,14,14,2048)) for the legacy path, randn(B,L,H) otherwise,
from torch.Generator().manual_seed(1234).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def legacy_weights(vocab_size=10000, seed=0):
    import capdec_b200 as cd
    torch.manual_seed(seed)
    m = cd.Decoder(vocab_size, False, "cpu").eval()
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def legacy_features(B, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(B, 14, 14, 2048, generator=g))


def lstm_decoder(kind, H=256, layers=2, heads=8, V=2000, seed=0, temperature=1.0, E=None):
    import capdec_b200 as cd
    torch.manual_seed(seed)
    dc = cd.DecoderConfig(decoder_type=cd.DecoderType.LSTM, hidden_dim=H, num_layers=layers, num_heads=8)
    ac = cd.AttentionConfig(attention_type=cd.AttentionType(kind), num_heads=heads, hidden_dim=H, temperature=temperature)
    m = cd.LSTMDecoder(dc, ac, vocab_size=V, pad_token_id=0, embedding_dim=E).eval()
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def lstm_inputs(B, L, H, seed=1234, ragged=False):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, L, H, generator=g)
    pooled = torch.randn(B, H, generator=g)
    mask = None
    if ragged:   # attention_mask: True = valid region
        mask = torch.ones(B, L, dtype=torch.bool)
        for b in range(B):
            keep = 1 + int(torch.randint(0, L, (1,), generator=g))
            mask[b, keep:] = False
    return feats, pooled, mask


def transformer_decoder(H=128, layers=2, heads=4, V=500, max_length=50, seed=0):
    import capdec_b200 as cd
    torch.manual_seed(seed)
    dc = cd.DecoderConfig(decoder_type=cd.DecoderType.TRANSFORMER, hidden_dim=H, num_layers=layers, num_heads=heads,
                          max_length=max_length)
    m = cd.TransformerDecoder(dc, vocab_size=V, pad_token_id=0, bos_token_id=1, eos_token_id=2).eval()
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def gpt2_decoder(H=64, layers=2, heads=4, V=300, max_length=64, seed=0, feature_dim=None):
    """GPT2Decoder drop-in with a random-init GPT2LMHeadModel (no pretrained weights offline)."""
    import capdec_b200 as cd
    torch.manual_seed(seed)
    dc = cd.DecoderConfig(decoder_type=cd.DecoderType.GPT2, pretrained_model_name="", hidden_dim=H, num_layers=layers,
                          num_heads=heads, dropout=0.0, max_length=max_length)
    m = cd.GPT2Decoder(dc, vocab_size=V, pad_token_id=0, bos_token_id=1, eos_token_id=2).eval()
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}
