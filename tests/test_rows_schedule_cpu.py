"""The issue schedule of the row-streaming conv3x3 kernel (csrc/conv3_rows.cuh), replayed in numpy: a CTA's range of row units
(image, 128-pixel strip, row) walked as segments, one halo row above and below each, the three 64-column groups of an N = 192 MMA
accumulating into the output rows y - 1, y, y + 1 of a ring of eight TMEM slots, split where the ring wraps, every slot drained once
and zeroed.  The replay must reproduce a direct 'same' convolution (UNet/model.py:30-35) for every split of the work over CTAs --
this pins the index arithmetic (tap order of the resident weights, group -> slot mapping, segment clipping, ragged last strip)
that the GPU parity cases (`rows_*` in tests/kernel_cases.py) then confirm on hardware."""
import numpy as np
import pytest

RW, RING = 128, 8


def replay(x, w, grid):
    """x [N, H, W, Cin], w [3, 3, Cin, Co] (out[h, w] = sum x[h + dh - 1, w + dw - 1] w[dh, dw]); returns (out, MMA pieces issued)"""
    N, H, W, Cin = x.shape
    Co = w.shape[-1]
    strips = (W + RW - 1) // RW
    total = N * strips * H
    out = np.full((N, H, W, Co), np.nan)
    pieces = 0
    for cta in range(grid):
        u, u1 = total * cta // grid, total * (cta + 1) // grid          # RowWalk
        tmem = np.zeros((RING, RW, Co))                                  # zeroed by the epilogue warps before the first MMA
        rc = 0
        while u < u1:
            strip = u // H
            hb = u - strip * H
            S = min(H - hb, u1 - u)
            img, w0 = strip // strips, (strip % strips) * RW
            u += S
            for i in range(-1, S + 1):
                g_lo = 1 - i if i < 1 else 0
                g_hi = S - i if S - i < 2 else 2
                row = np.zeros((RW + 2, Cin))                            # TMA box {64 ch, 130 px, 1 row}: out of bounds reads zero
                h = hb + i
                if 0 <= h < H:
                    lo, hi = max(w0 - 1, 0), min(w0 + RW + 1, W)
                    row[lo - (w0 - 1):hi - (w0 - 1)] = x[img, h, lo:hi]
                n = g_hi - g_lo + 1
                b = (rc + i - 1 + g_lo) & (RING - 1)
                n1 = min(n, RING - b)
                for dw in range(3):
                    a = row[dw:dw + RW]                                  # the A view: 128 rows starting at pixel dw
                    for slot0, g0, cnt in ((b, g_lo, n1), (0, g_lo + n1, n - n1)):
                        if cnt <= 0:
                            continue
                        pieces += 1
                        for k in range(cnt):                             # one MMA of N = 64 * cnt: group g = weights of filter row 2 - g
                            assert slot0 + k < RING
                            tmem[slot0 + k] += a @ w[2 - (g0 + k), dw]
                if i >= 1:                                               # tcgen05.commit -> rfull: output row i - 1 is complete
                    slot = (rc + i - 1) & (RING - 1)
                    hh = hb + i - 1
                    wv = min(RW, W - w0)                                 # TMA store clips the ragged strip
                    assert np.isnan(out[img, hh, w0:w0 + wv]).all(), "row written twice"
                    out[img, hh, w0:w0 + wv] = tmem[slot][:wv]
                    tmem[slot] = 0.0                                     # tcgen05.st of zeros before the slot is handed on
            rc += S
        assert not tmem.any()
    return out, pieces


def direct(x, w):
    N, H, W, _ = x.shape
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    out = np.zeros(x.shape[:3] + (w.shape[-1],))
    for dh in range(3):
        for dw in range(3):
            out += xp[:, dh:dh + H, dw:dw + W] @ w[dh, dw]
    return out


@pytest.mark.parametrize("N,H,W,grid", [(2, 11, 256, 5), (1, 1, 128, 1), (1, 2, 128, 2), (3, 7, 128, 4), (2, 33, 384, 7),
                                        (1, 16, 128, 16), (2, 9, 256, 36), (1, 10, 360, 3), (1, 5, 1000, 148 // 4)])
def test_rows_schedule_equals_direct_convolution(N, H, W, grid):
    rng = np.random.default_rng(N * 1000 + H * 10 + grid)
    x = rng.standard_normal((N, H, W, 6))
    w = rng.standard_normal((3, 3, 6, 4))
    got, _ = replay(x, w, grid)
    assert not np.isnan(got).any(), "an output row was never written"
    assert np.abs(got - direct(x, w)).max() < 1e-12


def test_rows_schedule_mma_count():
    """steady state: three column offsets per input row, one extra piece where the 3-slot window wraps around the ring (2 of 8 rows)"""
    x = np.zeros((1, 64, 128, 1))
    w = np.zeros((3, 3, 1, 1))
    _, pieces = replay(x, w, 1)
    rows_in = 64 + 2
    assert 3 * rows_in <= pieces <= 3 * rows_in + 3 * (rows_in // 4 + 2)
