"""Host-side restatement of the "p24" planar 24-bit tile format (csrc/common.cuh: p24_q_, p24_encode4_, p24_decode4_)
that the bf16x3 mode streams in the attention kernel: the integer identities the device code relies on and the error
bound DESIGN.md quotes (decoded value within 128.5 fp32 ulps, i.e. 2^-16 relative, of the original)."""
import numpy as np


def encode(x):
    bits = x.view(np.uint32)
    hi16 = (bits >> 16).astype(np.uint16)
    q = ((((bits & 0xFFFF).astype(np.uint64) + 128) * 65281) >> 24).astype(np.uint8)
    return hi16, q


def decode(hi16, q):
    return ((hi16.astype(np.uint32) << 16) | (q.astype(np.uint32) << 8) | q.astype(np.uint32)).view(np.float32)


def test_multiply_shift_equals_division_by_257():
    t = np.arange(0, 65536 + 128, dtype=np.uint64)
    assert np.array_equal((t * 65281) >> 24, t // 257)
    assert int(((np.uint64(65535 + 128) * 65281) >> 24)) == 255      # never overflows the byte, no carry into hi16


def test_round_trip_error_bound():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(1 << 20).astype(np.float32) * s for s in (1e-3, 1.0, 1e4)])
    x = np.concatenate([x, np.array([0.0, -0.0, 1.0, -1.0, np.float32(3.4e38), np.float32(1.2e-38), np.inf, -np.inf],
                                    dtype=np.float32)])
    hi16, q = encode(x)
    y = decode(hi16, q)
    fin = np.isfinite(x)
    ulps = np.abs(y[fin].view(np.uint32).astype(np.int64) - x[fin].view(np.uint32).astype(np.int64))
    assert ulps.max() <= 128
    nz = fin & (x != 0) & (np.abs(x) > 1e-30)
    assert (np.abs(y[nz].astype(np.float64) - x[nz]) / np.abs(x[nz].astype(np.float64))).max() <= 2.0 ** -16
    assert np.array_equal(np.isinf(y), np.isinf(x)) and np.array_equal(np.signbit(y), np.signbit(x))
    # the hi plane alone is the truncated bf16 the enc_att GEMM uses as its "hi" operand; the remainder is >= 0 in magnitude
    hi = (hi16.astype(np.uint32) << 16).view(np.float32)
    assert (np.abs(hi[fin]) <= np.abs(x[fin])).all()


def test_decode_is_monotone_in_q():
    hi16 = np.full(256, 0x3F80, dtype=np.uint16)      # 1.0 .. 1.0078
    q = np.arange(256, dtype=np.uint8)
    y = decode(hi16, q)
    assert (np.diff(y) > 0).all() and y[0] == 1.0
