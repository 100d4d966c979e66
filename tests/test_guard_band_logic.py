"""Host-side logic of the car-collision guard band (csrc/carfast.cuh: QMapView::amb_t3, dt_ball_fast), restated in
NumPy and checked exhaustively -- no GPU needed.

The kernel decides "this margin is within eps of its threshold" from the margin's top byte (sign | exponent >> 1) with
three integer instructions on the packed bytes of the three margins:

    clear = ((sg | 0x00808080) & 0x00ffffff) - T * 0x010101          (bit 7 of byte b survives <=> byte_b & 0x7f >= T)
    amb   = ~clear & word & 0x00808080

For the flags to stay bit-exact the test has to be CONSERVATIVE: every margin with |t| < eps must come out ambiguous
(then the float64 code decides).  These tests pin (1) the host's choice of T for every map size the library accepts,
(2) the byte arithmetic against a per-byte comparison, (3) the conservativeness over every float32 exponent."""
import math

import numpy as np


def host_T(rows, cols):
    """dt_qmap_view_of (carfast.cuh): eps of the map and the exponent threshold T derived from it."""
    mx = max(rows, cols)
    eps = np.float32(1.1920929e-7) * np.float32(mx + 4.0) + np.float32(0.075) * np.float32(1.5e-6) + np.float32(1.0e-7)
    eps = float(np.float32(eps))
    _, ex = math.frexp(eps)            # eps = m * 2^ex, m in [0.5, 1)
    T = (127 + ex + 1) // 2
    return eps, min(max(T, 1), 127)


def test_threshold_covers_eps_for_every_map_size():
    for n in range(1, 129):
        eps, T = host_T(n, n)
        assert 2.0 ** (2 * T - 127) >= eps, (n, eps, T)
        # ... and is not absurdly wide: at most a factor 4 above eps (the band costs exact-path decisions)
        assert 2.0 ** (2 * T - 127) < 4.0 * eps * 1.0000001, (n, eps, T)


def swar_clear(sg, T):
    return (((sg | np.uint32(0x00808080)) & np.uint32(0x00FFFFFF)) - np.uint32(T * 0x010101)) & np.uint32(0xFFFFFFFF)


def test_byte_arithmetic_matches_per_byte_comparison():
    rng = np.random.default_rng(0)
    sg = rng.integers(0, 2 ** 32, 200_000, dtype=np.uint64).astype(np.uint32)
    word = rng.integers(0, 2 ** 32, 200_000, dtype=np.uint64).astype(np.uint32)
    for T in (1, 37, 55, 56, 64, 127):
        clear = swar_clear(sg, T)
        amb = ~clear & word & np.uint32(0x00808080)
        want = np.zeros_like(sg)
        for b in range(3):
            byte = (sg >> np.uint32(8 * b)) & np.uint32(0x7F)
            flag = (word >> np.uint32(8 * b + 7)) & np.uint32(1)
            want |= ((byte < T) & (flag == 1)).astype(np.uint32) << np.uint32(8 * b + 7)
        assert np.array_equal(amb, want), T


def test_every_margin_inside_the_band_is_flagged():
    """All 2^9 (sign, exponent) combinations x a few mantissas: |t| < eps  =>  top byte & 0x7f < T."""
    eps, T = host_T(20, 20)
    mant = np.array([0, 1, 0x400000, 0x7FFFFF], dtype=np.uint32)
    for sign in (0, 1):
        for e in range(256):
            bits = (np.uint32(sign) << np.uint32(31)) | (np.uint32(e) << np.uint32(23)) | mant
            t = bits.view(np.float32)
            top = (bits >> np.uint32(24)) & np.uint32(0x7F)
            finite = np.isfinite(t)
            with np.errstate(invalid="ignore"):
                mag = np.abs(np.where(finite, t, np.float32(0)).astype(np.float64))
            inside = finite & (mag < eps)
            assert np.all(top[inside] < T), (sign, e)
            # the band the kernel actually applies: below 2^(2T - 126) at the widest
            flagged = top < T
            assert np.all(mag[flagged & finite] < 2.0 ** (2 * T - 126))
