"""Restated peak finders used by ``MTM._findLocalMax_`` (TEST INFRASTRUCTURE).

``skimage.feature.peak_local_max`` is called at ``MTM/__init__.py:45`` as
``peak_local_max(corrMap, threshold_abs=thr, exclude_border=False)`` and
``scipy.signal.find_peaks(row, height=thr)`` at ``MTM/__init__.py:34,40``.
scikit-image is not installed in this image, so its published algorithm
(skimage >= 0.19 ``feature/peak.py``: ``_get_peak_mask`` +
``_get_high_intensity_peaks``) is restated here for the argument subset MTM
uses (min_distance=1, 3x3 footprint, no labels, no num_peaks).  scipy is
present, so ``find_peaks_height`` is cross-checked against the live function in
tests/test_oracle.py.
"""
import numpy as np
from scipy import ndimage as ndi


def peak_local_max(image, threshold_abs, exclude_border=False):
    """Rows ``[r, c]`` of 3x3 local maxima strictly above ``threshold_abs``.

    Semantics restated (skimage >= 0.19):
    * a pixel is a candidate iff it equals the maximum of its 3x3 neighbourhood
      with ``mode='nearest'`` padding (== ignoring out-of-bounds neighbours);
    * if EVERY pixel is a candidate (constant map) there are no peaks;
    * candidates must be ``> threshold_abs`` (strict; numpy weak-scalar rules
      make this a float32 comparison for a float32 map);
    * output sorted by descending value, stable => row-major order on ties;
    * ``min_distance=1`` spacing filter only removes duplicates, so whole
      plateaus are returned.
    """
    if exclude_border:
        raise NotImplementedError("MTM always passes exclude_border=False")
    image = np.asarray(image)
    if image.size == 1:
        mask = image > threshold_abs
    else:
        image_max = ndi.maximum_filter(image, size=3, mode="nearest")
        mask = image == image_max
        if np.all(mask):
            mask[:] = False
        mask &= image > threshold_abs
    coord = np.nonzero(mask)
    intensities = image[coord]
    order = np.argsort(-intensities, kind="stable")
    return np.transpose(coord)[order]


def find_peaks_height(x, height):
    """Indices of 1-D peaks as ``scipy.signal.find_peaks(x, height=height)[0]``.

    scipy casts ``x`` to float64 first; a peak is a sample (or the floor-midpoint
    of a flat run) strictly above both neighbouring runs; the first and last
    samples are never peaks; ``height`` is inclusive (``>=``).  Order: ascending
    index.
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    out = []
    i = 1
    while i < n - 1:
        if x[i - 1] < x[i]:
            ahead = i + 1
            while ahead < n - 1 and x[ahead] == x[i]:
                ahead += 1
            if x[ahead] < x[i]:
                mid = (i + ahead - 1) // 2
                if x[mid] >= height:
                    out.append(mid)
                i = ahead
        i += 1
    return np.asarray(out, dtype=np.intp)
