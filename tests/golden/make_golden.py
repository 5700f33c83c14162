"""Generate the committed golden vectors by running the REFERENCE's own modules (needs /root/reference,
i.e. the build container; the GPU box only consumes the .pt files).

    python tests/golden/make_golden.py

Every fixture stores the seeds/config needed to rebuild weights and inputs (tests/helpers.py) plus the
outputs.  Weights are not stored (25M+ parameters); a CPU test asserts the drop-in modules' seeded init
equals the reference modules' init, so rebuilding from the seed is exact.
  legacy_teacher.pt   models/decoder.py::Decoder.forward (the reference module itself)
  legacy_beam{3,5}.pt oracle beam driver (pinned to HF) over the legacy step; the step restatement is
                      bit-identical to Decoder.forward (asserted here before writing)
  lstm_greedy_*.pt    src/models/decoders.py::LSTMDecoder.generate (the reference module itself)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import beam as obeam, legacy as olegacy, refshim  # noqa: E402
from tests.helpers import GOLDEN, legacy_features, legacy_weights, lstm_decoder, lstm_inputs, transformer_decoder  # noqa: E402


def main():
    ns = refshim.load_reference()
    torch.set_grad_enabled(False)

    # ---- legacy teacher-forced forward, straight from the reference module
    V = 10000
    torch.manual_seed(0)
    ref = ns.legacy.Decoder(V, False, "cpu").eval()
    _, sd = legacy_weights(V, 0)
    assert all(torch.equal(sd[k], v) for k, v in ref.state_dict().items())
    B = 5
    enc = legacy_features(B)
    g = torch.Generator().manual_seed(77)
    lens = [9, 8, 8, 5, 3]
    caps = torch.randint(0, V, (B, 9), generator=g)
    preds, _, dec_len, alphas = ref(enc, caps, lens)
    p2, a2, _ = olegacy.forward_teacher_forced(sd, enc, caps, lens)
    assert torch.equal(preds, p2) and torch.equal(alphas, a2), "oracle restatement drifted from the reference"
    torch.save(dict(vocab=V, seed=0, feat_seed=1234, B=B, cap_seed=77, lens=lens, caps=caps,
                    preds_sub=preds[:, :, ::97].clone(), argmax=preds.argmax(-1), alphas=alphas.clone()),
               os.path.join(GOLDEN, "legacy_teacher.pt"))

    # ---- legacy beam search (config-1 shape: beam 3, max_len 20, vocab 10k; and beam 5)
    for k, B in ((3, 8), (5, 6)):
        enc = legacy_features(B)
        out = obeam.beam_search(olegacy.LegacyStepper(sd, enc, k), B, k, 20, bos_token_id=1, eos_token_id=2,
                                pad_token_id=0, record_steps=True)
        torch.save(dict(vocab=V, seed=0, feat_seed=1234, B=B, k=k, T=20, sequences=out["sequences"],
                        lengths=out["lengths"], scores=out["scores"],
                        top_lp=torch.stack([s["top_lp"] for s in out["steps"]]),
                        top_tok=torch.stack([s["top_tok"] for s in out["steps"]]),
                        top_beam=torch.stack([s["top_beam"] for s in out["steps"]])),
                   os.path.join(GOLDEN, f"legacy_beam{k}.pt"))
        print("legacy beam", k, out["sequences"][:2].tolist(), out["scores"][:2].tolist())

    # ---- src LSTMDecoder.generate (greedy) from the reference module, every attention type
    C = ns.config
    cases = [("soft", 8, 512, 1, 196, 10000, False), ("soft", 8, 256, 2, 49, 2000, True),
             ("multi_head", 8, 256, 2, 49, 2000, True), ("aoa", 8, 256, 2, 49, 2000, False),
             ("aoa", 1, 256, 1, 49, 2000, True), ("adaptive", 8, 256, 2, 49, 2000, False),
             ("adaptive", 1, 256, 1, 36, 2000, True)]
    for kind, heads, H, layers, L, V2, ragged in cases:
        torch.manual_seed(0)
        dc = C.DecoderConfig(decoder_type=C.DecoderType.LSTM, hidden_dim=H, num_layers=layers, num_heads=8)
        ac = C.AttentionConfig(attention_type=C.AttentionType(kind), num_heads=heads, hidden_dim=H)
        ref = ns.decoders.LSTMDecoder(dc, ac, vocab_size=V2, pad_token_id=0).eval()
        _, sd2 = lstm_decoder(kind, H=H, layers=layers, heads=heads, V=V2, seed=0)
        assert all(torch.equal(sd2[k_], v) for k_, v in ref.state_dict().items())
        B = 6
        feats, pooled, mask = lstm_inputs(B, L, H, ragged=ragged)
        ef = {"features": feats, "pooled_features": pooled}
        if mask is not None:
            ef["attention_mask"] = mask
        ids, info = ref.generate(ef, 20)
        name = f"lstm_greedy_{kind}_h{heads}_H{H}_l{layers}_L{L}{'_ragged' if ragged else ''}.pt"
        torch.save(dict(kind=kind, heads=heads, H=H, layers=layers, L=L, vocab=V2, ragged=ragged, B=B, T=20, seed=0,
                        feat_seed=1234, ids=ids, attention_weights=info["attention_weights"]),
                   os.path.join(GOLDEN, name))
        print(name, ids[0].tolist()[:8])

    # ---- src TransformerDecoder.generate (greedy, full-prefix recompute) from the reference module
    for H, layers, heads, V2, L, B, T in ((128, 2, 4, 500, 49, 5, 12), (768, 6, 8, 10000, 196, 3, 20)):
        torch.manual_seed(0)
        dc = C.DecoderConfig(decoder_type=C.DecoderType.TRANSFORMER, hidden_dim=H, num_layers=layers, num_heads=heads,
                             max_length=50)
        ref = ns.decoders.TransformerDecoder(dc, vocab_size=V2, pad_token_id=0, bos_token_id=1, eos_token_id=2).eval()
        _, sd3 = transformer_decoder(H=H, layers=layers, heads=heads, V=V2, seed=0)
        assert all(torch.equal(sd3[k_], v) for k_, v in ref.state_dict().items())
        feats, _, _ = lstm_inputs(B, L, H)
        ids, _ = ref.generate({"features": feats}, T)
        name = f"transformer_greedy_H{H}_l{layers}_h{heads}_L{L}.pt"
        torch.save(dict(H=H, layers=layers, heads=heads, vocab=V2, L=L, B=B, T=T, seed=0, feat_seed=1234, ids=ids),
                   os.path.join(GOLDEN, name))
        print(name, ids[0].tolist()[:8])


if __name__ == "__main__":
    main()
