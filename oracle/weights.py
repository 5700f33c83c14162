"""Seeded random-init weights of the reference's legacy decoder built with plain torch.nn (TEST INFRASTRUCTURE ONLY).

`bench.py --impl reference` must not import the product package (its arm is the CPU oracle only), so the weights the
GPU arm gets from constructing `capdec_b200.Decoder` under torch.manual_seed(seed) are rebuilt here by instantiating the
same torch.nn layers in the same order as /root/reference/models/decoder.py:33-55 -- same RNG consumption, hence the same
tensors (asserted in tests/test_oracle_pin.py::test_plain_torch_weights_equal_dropin_init)."""
import torch
from torch import nn


def legacy_state_dict(vocab_size=10000, seed=0, encoder_dim=2048, attention_dim=512, embed_dim=512, decoder_dim=512):
    torch.manual_seed(seed)
    layers = {}
    layers["enc_att"] = nn.Linear(encoder_dim, attention_dim)            # models/decoder.py:33
    layers["dec_att"] = nn.Linear(decoder_dim, attention_dim)            # :35
    layers["att"] = nn.Linear(attention_dim, 1)                          # :37
    layers["decode_step"] = nn.LSTMCell(embed_dim + encoder_dim, decoder_dim, bias=True)   # :41
    layers["h_lin"] = nn.Linear(encoder_dim, decoder_dim)                # :42
    layers["c_lin"] = nn.Linear(encoder_dim, decoder_dim)                # :43
    layers["f_beta"] = nn.Linear(decoder_dim, encoder_dim)               # :45
    layers["fc"] = nn.Linear(decoder_dim, vocab_size)                    # :47
    layers["fc"].bias.data.fill_(0)                                      # :50-51
    layers["fc"].weight.data.uniform_(-0.1, 0.1)
    layers["embedding"] = nn.Embedding(vocab_size, embed_dim)            # :53-55
    layers["embedding"].weight.data.uniform_(-0.1, 0.1)
    sd = {}
    for name, mod in layers.items():
        for k, v in mod.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd
