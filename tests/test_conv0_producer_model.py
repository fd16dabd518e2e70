"""CPU checks of two facts the raw-image producer of conv0 relies on (rusty_sr_b200/csrc/conv_strip.cuh, raw_row_producer_role):
(1) u8 / 255 computed as x * (1/255) plus one Newton step is the IEEE-754 single-precision quotient for all 256 inputs, i.e. exactly
    img_to_data's `v as f32 / 255.0` (reference main.rs:170), without the division's slow path;
(2) the address a pixel's 16-byte pieces are written to is the 64-byte TMA swizzle (address bits 4-5 XOR bits 7-8), the layout a TMA
    load of an x-im2col tensor with 64-byte rows would have produced and the UMMA SWIZZLE_64B descriptor reads."""
from fractions import Fraction

import numpy as np


def fl32(fr: Fraction) -> np.float32:
    """round-to-nearest-even of an exact rational to float32"""
    x = np.float32(float(fr))
    cands = [x, np.nextafter(x, np.float32(np.inf)), np.nextafter(x, np.float32(-np.inf))]
    return np.float32(min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(np.float32(c).view(np.uint32)) & 1)))


def test_u8_over_255_newton_step_is_the_ieee_quotient():
    rcp = np.float32(1.0) / np.float32(255.0)
    plain_mismatches = 0
    for b in range(256):
        want = np.float32(b) / np.float32(255.0)
        q0 = fl32(Fraction(b) * Fraction(float(rcp)))                                   # x * rcp
        r = fl32(Fraction(b) - Fraction(float(q0)) * 255)                               # fmaf(-q0, 255, x): one rounding
        q1 = fl32(Fraction(float(r)) * Fraction(float(rcp)) + Fraction(float(q0)))      # fmaf(r, rcp, q0)
        assert q1 == want, (b, q0, q1, want)
        plain_mismatches += q0 != want
    assert plain_mismatches > 100      # the bare multiplication is NOT enough: the step is needed


def test_producer_rows_follow_the_64_byte_tma_swizzle():
    seen = set()
    for r in range(128):                       # pixel of the strip = 64-byte row of the slot
        for j in range(4):                     # 16-byte piece: bf16 hi 0-7, hi 8-15, mid 0-7, mid 8-15
            addr = r * 64 + ((j ^ ((r >> 1) & 3)) << 4)              # what raw_row_producer_role computes
            linear = r * 64 + j * 16
            swz = linear ^ (((linear >> 7) & 3) << 4)                # Swizzle<2, 4, 3>: bits 4-5 ^= bits 7-8
            assert addr == swz, (r, j)
            seen.add(addr)
    assert len(seen) == 512 and max(seen) == 128 * 64 - 16          # a permutation of the slot's 16-byte pieces
    # a warp-wide 16-byte store (lane = row, fixed piece) touches every bank group once per quarter warp: no conflicts
    for j in range(4):
        for q in range(4):
            banks = {((r * 64 + ((j ^ ((r >> 1) & 3)) << 4)) // 16) % 8 for r in range(8 * q, 8 * q + 8)}
            assert len(banks) == 8
