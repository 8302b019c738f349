"""CPU port of the MTM hot path on live OpenCV (TEST INFRASTRUCTURE / CPU baseline).

Restates, in its own words, the orchestration of the reference
(``MTM/__init__.py:22-296`` and ``MTM/NMS.py:20-84``) on top of the SAME
third-party kernels the reference calls (``cv2.matchTemplate``,
``cv2.minMaxLoc``, ``cv2.dnn.NMSBoxes``, ``scipy.signal.find_peaks``) plus the
restated ``peak_local_max`` of ``oracle/peaks.py`` (scikit-image is absent here).
It exists because /root/reference cannot travel to the GPU box; it is pinned
against the unmodified reference by tests/golden/ref_*.json (made by
``oracle/make_golden.py``).  ``bench.py --impl reference`` times THIS code.

Differences from the reference, on purpose:
* hits are gathered in template order (the reference appends from pool threads
  in completion order, ``MTM/__init__.py:172-175,244`` -- any order it can
  produce is a permutation of this one; post-NMS results are order-independent
  except for exact score ties);
* ``workers`` is a parameter (reference: ``round(os.cpu_count()*.5)``).
"""
import os
import warnings
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np
from scipy.signal import find_peaks

from .peaks import peak_local_max

INF = float("inf")


def find_local_max(score_map, threshold):
    """``MTM/__init__.py:22-47``: branch on the map shape."""
    mh, mw = score_map.shape
    if (mh, mw) == (1, 1):
        return [[0, 0]] if score_map[0, 0] >= threshold else []
    if mh == 1:
        return [[0, int(i)] for i in find_peaks(score_map[0], height=threshold)[0]]
    if mw == 1:
        return [[int(i), 0] for i in find_peaks(score_map[:, 0], height=threshold)[0]]
    return peak_local_max(score_map, threshold_abs=threshold).tolist()


def find_local_min(score_map, threshold):
    """``MTM/__init__.py:51-53``."""
    return find_local_max(-score_map, -threshold)


def compute_score_map(template, image, method=cv2.TM_CCOEFF_NORMED, mask=None):
    """``MTM/__init__.py:56-92``: dtype policy + mask policy + cv2.matchTemplate."""
    if template.dtype == "float64" or image.dtype == "float64":
        raise ValueError("64-bit images not supported, max 32-bit")
    if not (template.dtype == "uint8" and image.dtype == "uint8"):
        template = np.float32(template)
        image = np.float32(image)
        if mask is not None:
            mask = np.float32(mask)
    if mask is not None:
        if method not in (0, 3):
            mask = None
            warnings.warn("Template matching method not compatible with use of mask (only 0/TM_SQDIFF or 3/TM_CCORR_NORMED).\n-> Ignoring mask.")
        elif not (mask.shape == template.shape and mask.dtype == template.dtype):
            mask = None
            warnings.warn("Mask does not have the same dimension or bit depth than the template.\n-> Ignoring mask.")
    return cv2.matchTemplate(image, template, method, mask=mask)


def _one_template(entry, image, method, n_object, threshold, x_off, y_off):
    """``MTM/__init__.py:179-244`` for one template; returns its hits."""
    name, template = entry[:2]
    mask = None
    if len(entry) >= 3:
        if method in (0, 3):
            mask = entry[2]
        else:
            warnings.warn("Template matching method not supporting the use of Mask. Use 0/TM_SQDIFF or 3/TM_CCORR_NORMED.")
    score_map = compute_score_map(template, image, method, mask=mask)
    if n_object == 1:
        _, _, min_loc, max_loc = cv2.minMaxLoc(score_map)
        loc = min_loc if method in (0, 1) else max_loc
        peaks = [loc[::-1]]
    elif method in (0, 1):
        peaks = find_local_min(score_map, threshold)
    else:
        peaks = find_local_max(score_map, threshold)
    th, tw = template.shape[0:2]
    return [(name, (int(p[1]) + x_off, int(p[0]) + y_off, tw, th), score_map[tuple(p)]) for p in peaks]


def find_matches(templates, image, method=cv2.TM_CCOEFF_NORMED, N_object=INF,
                 score_threshold=0.5, searchBox=None, workers=None):
    """``MTM/__init__.py:95-177``."""
    if N_object != INF and not isinstance(N_object, int):
        raise TypeError("N_object must be an integer")
    if image.shape[0] == 0:
        raise ValueError("Image has a height of 0.")
    if image.shape[1] == 0:
        raise ValueError("Image has a width of 0.")
    x_off = y_off = 0
    if searchBox is not None:
        x_off, y_off, bw, bh = searchBox
        image = image[y_off:y_off + bh, x_off:x_off + bw]
    for index, entry in enumerate(templates):
        if not isinstance(entry, tuple) or len(entry) < 2:
            raise ValueError("listTemplates should be a list of tuples as ('name','array') or ('name', 'array', 'mask')")
        name, arr = entry[0], entry[1]
        if arr.shape[0] == 0:
            raise ValueError(f"Template '{name}' has a height of 0.")
        if arr.shape[1] == 0:
            raise ValueError(f"Template '{name}' has a width of 0.")
        if not all(t <= i for t, i in zip(arr.shape, image.shape)):
            where = "searchBox" if searchBox is not None else "image"
            raise ValueError("Template '{}' at index {} in the list of templates is larger than {}.".format(name, index, where))
    if workers is None:
        workers = round(os.cpu_count() * .5)
    with ThreadPoolExecutor(max_workers=workers) as pool:
        per_template = list(pool.map(
            lambda e: _one_template(e, image, method, N_object, score_threshold, x_off, y_off),
            templates))
    hits = []
    for part in per_template:
        hits.extend(part)
    return hits


def nms(hits, scoreThreshold=0.5, sortAscending=False, N_object=INF, maxOverlap=0.5):
    """``MTM/NMS.py:20-84``."""
    if len(hits) <= 1:
        return hits[:]
    if N_object == 1:
        pick = min if sortAscending else max
        return [pick(hits, key=lambda hit: hit[2])]
    boxes = [hit[1] for hit in hits]
    scores = [hit[2] for hit in hits]
    if sortAscending:
        scores = [1 - s for s in scores]
        scoreThreshold = 1 - scoreThreshold
    keep = cv2.dnn.NMSBoxes(boxes, scores, scoreThreshold, maxOverlap)
    if N_object != INF:
        keep = keep[:N_object]
    return [hits[i] for i in keep]


def match_templates(templates, image, method=cv2.TM_CCOEFF_NORMED, N_object=INF,
                    score_threshold=0.5, maxOverlap=0.25, searchBox=None, workers=None):
    """``MTM/__init__.py:247-296``."""
    if maxOverlap < 0 or maxOverlap > 1:
        raise ValueError("Maximal overlap between bounding box is in range [0-1]")
    hits = find_matches(templates, image, method, N_object, score_threshold, searchBox, workers=workers)
    if method == 0:
        raise ValueError("The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.")
    return nms(hits, score_threshold, method == 1, N_object, maxOverlap)
