"""CPU model of the row-folded MMA schedule of conv_strip.cuh (ConvCfg::FOLD; experimental, off by default): replays
the kernel's loop — stage = tap column, one "MMA" per input row of the pass window whose N spans the output rows that row
feeds, weights of a tap column stored bottom tap first, first product into a row overwrites — on small random data and
checks it against the direct convolution.  Accumulators start as NaN, so an output row that were accumulated onto before
its overwriting product, or never written, fails the comparison."""
import numpy as np
import pytest


def folded_pass(win, wt, KH, KW, R, ncol):
    """win: [R + KH - 1][M + KW - 1][K] input window rows (zero padded); wt: [KH][KW][ncol][K].
    Returns D [M][R * ncol] the way the kernel's TMEM columns are laid out (row r at columns r*ncol..)."""
    WIN = R + KH - 1
    M = win.shape[1] - (KW - 1)
    D = np.full((M, R * ncol), np.nan)
    # fold-ordered weights: stage tb, block j = KH-1-ta  (rsr_b200.cu: w0f / w4f)
    wf = np.zeros((KW, KH, ncol, wt.shape[3]))
    for ta in range(KH):
        for tb in range(KW):
            wf[tb, KH - 1 - ta] = wt[ta, tb]

    def mma(dcol, a, b_rows, acc):   # D[:, dcol : dcol + rows] (+)= a @ b_rows^T
        n = b_rows.shape[0]
        prod = a @ b_rows.T
        if acc:
            D[:, dcol:dcol + n] += prod
        else:
            D[:, dcol:dcol + n] = prod

    for sp in range(KW):                       # stage = tap column
        for i in range(WIN):                   # input row of the window
            r_lo = max(i - (KH - 1), 0)
            r_hi = min(i, R - 1)
            n = r_hi - r_lo + 1
            a = win[i, sp:sp + M, :]           # descriptor start shifted by sp pixels
            blk = wf[sp].reshape(KH * ncol, -1)
            b0 = (KH - 1 - (i - r_lo)) * ncol  # first weight row of the block
            d = r_lo * ncol
            if sp == 0 and i < R:              # first product that reaches output row i
                if n > 1:
                    mma(d, a, blk[b0:b0 + (n - 1) * ncol], True)
                mma(d + (n - 1) * ncol, a, blk[b0 + (n - 1) * ncol:b0 + n * ncol], False)
            else:
                mma(d, a, blk[b0:b0 + n * ncol], True)
    return D


@pytest.mark.parametrize("KH,KW,R", [(3, 3, 2), (3, 3, 3), (3, 3, 4), (5, 1, 8), (5, 1, 4), (5, 1, 2), (3, 3, 1), (5, 5, 2)])
def test_folded_schedule_equals_direct_convolution(KH, KW, R):
    rng = np.random.default_rng(KH * 100 + KW * 10 + R)
    M, K, ncol = 7, 4, 3
    win = rng.standard_normal((R + KH - 1, M + KW - 1, K))
    wt = rng.standard_normal((KH, KW, ncol, K))
    D = folded_pass(win, wt, KH, KW, R, ncol)
    ref = np.zeros((M, R * ncol))
    for r in range(R):
        for ta in range(KH):
            for tb in range(KW):
                ref[:, r * ncol:(r + 1) * ncol] += win[r + ta, tb:tb + M, :] @ wt[ta, tb].T
    assert not np.isnan(D).any()
    assert np.allclose(D, ref, rtol=1e-12, atol=1e-12)
