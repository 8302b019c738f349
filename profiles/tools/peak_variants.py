"""Clocks per tcgen05.mma kind::i8 (M128 x N x K32) of the measurement kernel for several N, one variant per process
(MTM_B200_PEAK_VARIANT: bit 0 = B descriptor shifts 16 bytes per MMA, bit 1 = A walks over 8 slabs, bit 2 = a commit every 6 MMAs).

    MTM_B200_PEAK_VARIANT=1 python profiles/tools/peak_variants.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import MTM  # noqa: E402,F401
from mtm_b200 import _native  # noqa: E402

ctx = _native.default_context()
sm = 148
for n in [int(v) for v in os.environ.get("PEAK_N", "128,144,160,208,240,256").split(",")]:
    tmacs = ctx.measure_i8_peak(n, 4000)
    ms = sm * 4000 * 128.0 * n * 32.0 / (tmacs * 1e12) * 1e3
    print("variant %s N=%3d: %.3f TMAC/s = %.3f POP/s, %.4f ms per 4000 MMAs" % (os.environ.get("MTM_B200_PEAK_VARIANT", "0"), n, tmacs, 2 * tmacs / 1e3, ms))
