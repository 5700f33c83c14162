"""oracle/ -- TEST INFRASTRUCTURE ONLY.  Not shipped, not measured, never on the product path.

CPU (torch fp32) restatement of the caption-decoding hot path of
thromel/Image-Captioning-ML-Project, used as the checker for the CUDA path:

  oracle/refshim.py    import shims that let the UNMODIFIED reference modules load in
                       this container (only used where /root/reference exists: fixture
                       generation and the "pin the oracle" tests).
  oracle/legacy.py     models/decoder.py::Decoder  (init state + per-step body + teacher-forced forward)
  oracle/attention.py  src/models/attention.py     (Soft / MultiHead / Adaptive / AoA forward)
  oracle/lstm.py       src/models/decoders.py::LSTMDecoder (init states, nn.LSTM single step, greedy generate)
  oracle/beam.py       transformers GenerationMixin._beam_search (the only beam search the reference
                       ever invokes, src/models/decoders.py:645) restated over a generic step function
  oracle/sample.py     src/train/trainer.py::_sample_captions restated with explicit uniforms

Parity status: PINNED.  Every restatement is checked (tests/test_oracle_pin.py, run in the build
container) against the reference's own modules imported from /root/reference through refshim, and
against golden vectors produced by those modules and committed under tests/golden/
(generator: tests/golden/make_golden.py).  The beam driver is additionally pinned against
transformers' own `generate(num_beams=k)` on a small random GPT-2.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (image-captioning-ml-project_b200/) never does.
"""
