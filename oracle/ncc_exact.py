"""Exact restatement of ``cv2.matchTemplate`` for the MTM path (TEST INFRASTRUCTURE).

``cv2.matchTemplate(image, template, method)`` is the only arithmetic call of
``MTM.computeScoreMap`` (``MTM/__init__.py:92``).  OpenCV is a third-party
dependency that is not vendored under /root/reference (``setup.py:23``:
``opencv-python-headless>=4.5.4``; this image has 4.13.0), so its published
algorithm (``modules/imgproc/src/templmatch.cpp``: ``crossCorr`` followed by
``common_matchTemplate``) is restated here:

* numerator ``CC = sum I*T`` -- computed EXACTLY (int64 for uint8 inputs, float64
  for float32) where OpenCV uses an fp32 DFT / IPP;
* window sums ``S = sum I`` and ``Q = sum I^2`` per channel from summed-area tables
  (OpenCV: ``integral(img, sum, sqsum, CV_64F)``) -- exact integers here;
* the per-pixel float64 epilogue with OpenCV's clamps, then a cast to float32.

The restatement is pinned against live cv2 in tests/test_oracle.py (max
difference ~2e-6 on well-conditioned inputs; cv2's own fp32 noise).
"""
import ctypes
import os

import numpy as np

_FLT_EPSILON = float(np.finfo(np.float32).eps)
_DBL_EPSILON = float(np.finfo(np.float64).eps)

TM_SQDIFF, TM_SQDIFF_NORMED, TM_CCORR, TM_CCORR_NORMED, TM_CCOEFF, TM_CCOEFF_NORMED = range(6)

_lib = None


def _load():
    global _lib
    if _lib is None:
        from . import build as _b
        path = _b.LIB if os.path.exists(_b.LIB) else _b.build()
        lib = ctypes.CDLL(path)
        lib.oracle_cc_u8.restype = ctypes.c_int
        lib.oracle_cc_f32.restype = ctypes.c_int
        _lib = lib
    return _lib


def _as3(a):
    a = np.asarray(a)
    return a[:, :, None] if a.ndim == 2 else a


def cc_direct(image, templ):
    """Exact ``sum I*T`` by the direct double loop in C (int64 / float64)."""
    I = np.ascontiguousarray(_as3(image))
    T = np.ascontiguousarray(_as3(templ))
    H, W, C = I.shape
    h, w, c = T.shape
    assert c == C
    lib = _load()
    if I.dtype == np.uint8 and T.dtype == np.uint8:
        out = np.empty((H - h + 1, W - w + 1), np.int64)
        rc = lib.oracle_cc_u8(I.ctypes.data_as(ctypes.c_void_p), H, W, C,
                              T.ctypes.data_as(ctypes.c_void_p), h, w,
                              out.ctypes.data_as(ctypes.c_void_p))
    else:
        I = np.ascontiguousarray(I, np.float32)
        T = np.ascontiguousarray(T, np.float32)
        out = np.empty((H - h + 1, W - w + 1), np.float64)
        rc = lib.oracle_cc_f32(I.ctypes.data_as(ctypes.c_void_p), H, W, C,
                               T.ctypes.data_as(ctypes.c_void_p), h, w,
                               out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("oracle_cc: bad shapes")
    return out


def cc_fft(image, templ):
    """``sum I*T`` through a float64 FFT; for uint8 inputs the result is rounded
    to the nearest integer, which is exact while the FFT error stays < 0.5
    (checked against ``cc_direct`` in tests/test_oracle.py)."""
    from scipy import fft as sfft
    I = _as3(image)
    T = _as3(templ)
    H, W, C = I.shape
    h, w, _ = T.shape
    fh, fw = sfft.next_fast_len(H, real=True), sfft.next_fast_len(W, real=True)
    acc = np.zeros((H - h + 1, W - w + 1), np.float64)
    for c in range(C):
        Fi = sfft.rfft2(I[:, :, c].astype(np.float64), s=(fh, fw))
        Ft = sfft.rfft2(T[::-1, ::-1, c].astype(np.float64), s=(fh, fw))
        full = sfft.irfft2(Fi * Ft, s=(fh, fw))
        acc += full[h - 1:H, w - 1:W]
    if I.dtype == np.uint8 and T.dtype == np.uint8:
        return np.rint(acc).astype(np.int64)
    return acc


def window_sums(image, h, w):
    """Per-channel window sums ``S`` and all-channel ``Q`` (float64, exact for uint8)."""
    I = _as3(image)
    H, W, C = I.shape
    wide = np.int64 if I.dtype == np.uint8 else np.float64
    S = np.empty((C, H - h + 1, W - w + 1), np.float64)
    Q = np.zeros((H - h + 1, W - w + 1), np.float64)
    for c in range(C):
        p = I[:, :, c].astype(wide)
        sat = np.zeros((H + 1, W + 1), wide)
        sat[1:, 1:] = p.cumsum(0).cumsum(1)
        sq = np.zeros((H + 1, W + 1), wide)
        sq[1:, 1:] = (p * p).cumsum(0).cumsum(1)
        S[c] = sat[h:, w:] - sat[:-h, w:] - sat[h:, :-w] + sat[:-h, :-w]
        Q += sq[h:, w:] - sq[:-h, w:] - sq[h:, :-w] + sq[:-h, :-w]
    return S, Q


def epilogue(cc, S, Q, templ, method):
    """OpenCV ``common_matchTemplate`` (templmatch.cpp) restated, vectorised in float64.

    ``cc`` is the numerator map, ``S`` (C,mh,mw) per-channel window sums, ``Q`` the
    window sum of squares over all channels.  Returns float32.
    """
    T = _as3(templ).astype(np.float64)
    h, w, C = T.shape
    num = np.asarray(cc, np.float64).copy()
    if method == TM_CCORR:
        return num.astype(np.float32)
    area = float(h * w)
    inv_area = 1.0 / area
    num_type = 0 if method in (TM_CCORR, TM_CCORR_NORMED) else (1 if method in (TM_CCOEFF, TM_CCOEFF_NORMED) else 2)
    is_normed = method in (TM_SQDIFF_NORMED, TM_CCORR_NORMED, TM_CCOEFF_NORMED)

    t_mean = T.reshape(-1, C).sum(0) * inv_area                     # meanStdDev: mean
    t_var = np.maximum((T.reshape(-1, C) ** 2).sum(0) * inv_area - t_mean ** 2, 0.0)
    templ_norm = float(t_var.sum())
    templ_sum2 = 0.0
    if method != TM_CCOEFF:
        if templ_norm < _DBL_EPSILON and method == TM_CCOEFF_NORMED:
            return np.ones_like(num, dtype=np.float32)
        templ_sum2 = templ_norm + float((t_mean ** 2).sum())
        if num_type != 1:
            t_mean = np.zeros_like(t_mean)
            templ_norm = templ_sum2
        templ_sum2 /= inv_area
        templ_norm = np.sqrt(templ_norm)
        templ_norm /= np.sqrt(inv_area)

    wnd_mean2 = np.zeros_like(num)
    wnd_sum2 = np.zeros_like(num)
    if num_type == 1:
        for c in range(C):
            wnd_mean2 += S[c] * S[c]
            num -= S[c] * t_mean[c]
        wnd_mean2 *= inv_area
    if is_normed or num_type == 2:
        wnd_sum2 = Q.astype(np.float64)
        if num_type == 2:
            num = np.maximum(wnd_sum2 - 2.0 * num + templ_sum2, 0.0)
    if is_normed:
        diff2 = np.maximum(wnd_sum2 - wnd_mean2, 0.0)
        t = np.where(diff2 <= np.minimum(0.5, 10.0 * _FLT_EPSILON * wnd_sum2),
                     0.0, np.sqrt(diff2) * templ_norm)
        a = np.abs(num)
        safe_t = np.where(t > 0, t, 1.0)
        fallback = 1.0 if method == TM_SQDIFF_NORMED else 0.0
        num = np.where(a < t, num / safe_t,
                       np.where(a < t * 1.125, np.where(num > 0, 1.0, -1.0), fallback))
    return num.astype(np.float32)


def match_template_exact(image, templ, method=TM_CCOEFF_NORMED, use_fft=None):
    """Ground-truth score map for ``cv2.matchTemplate(image, templ, method)``."""
    I = _as3(image)
    T = _as3(templ)
    if not (I.dtype == np.uint8 and T.dtype == np.uint8):
        I = I.astype(np.float32)
        T = T.astype(np.float32)
    h, w, _ = T.shape
    if use_fft is None:
        macs = float(h) * w * (I.shape[0] - h + 1) * (I.shape[1] - w + 1) * I.shape[2]
        use_fft = macs > 2e10
    cc = cc_fft(I, T) if use_fft else cc_direct(I, T)
    S, Q = window_sums(I, h, w)
    return epilogue(cc, S, Q, T, method)


def _corr_valid_f64(I, K):
    """sum_c sum I[y+dy, x+dx, c] * K[dy, dx, c] ('valid' region) in float64 via FFT."""
    from scipy import fft as sfft
    I = _as3(I).astype(np.float64)
    K = _as3(K).astype(np.float64)
    H, W, C = I.shape
    h, w, _ = K.shape
    fh, fw = sfft.next_fast_len(H, real=True), sfft.next_fast_len(W, real=True)
    acc = np.zeros((H - h + 1, W - w + 1), np.float64)
    for c in range(C):
        Fi = sfft.rfft2(I[:, :, c], s=(fh, fw))
        Fk = sfft.rfft2(K[::-1, ::-1, c], s=(fh, fw))
        acc += sfft.irfft2(Fi * Fk, s=(fh, fw))[h - 1:H, w - 1:W]
    return acc


def match_template_masked_exact(image, templ, mask, method):
    """``cv2.matchTemplate(image, templ, method, mask=mask)`` for TM_SQDIFF (0) and TM_CCORR_NORMED (3):
    OpenCV ``matchTemplateMask`` (templmatch.cpp) restated in float64.  uint8 masks are binary
    (non-zero -> 1), float32 masks are weights.  Reached from ``MTM/__init__.py:92`` when a template
    tuple carries a mask and the method is 0 or 3 (``MTM/__init__.py:76-88, 213-217``)."""
    I = _as3(image).astype(np.float64)
    T = _as3(templ).astype(np.float64)
    M = _as3(mask)
    M = (M > 0).astype(np.float64) if M.dtype == np.uint8 else M.astype(np.float64)
    if M.shape[2] == 1 and T.shape[2] > 1:
        M = np.repeat(M, T.shape[2], axis=2)
    M2 = M * M
    cc = _corr_valid_f64(I, T * M2)
    q = _corr_valid_f64(I * I, M2)
    t2 = float(((T * M) ** 2).sum())
    if method == TM_SQDIFF:
        return (q - 2.0 * cc + t2).astype(np.float32)
    if method == TM_CCORR_NORMED:
        with np.errstate(divide="ignore", invalid="ignore"):
            return (cc / np.sqrt(t2 * q)).astype(np.float32)
    raise ValueError("masks are only defined for methods 0 and 3")
