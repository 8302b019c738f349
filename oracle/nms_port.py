"""Restated ``cv2.dnn.NMSBoxes`` for integer rectangles (TEST INFRASTRUCTURE).

Call site: ``MTM/NMS.py:78`` -- ``cv2.dnn.NMSBoxes(listBoxes, listScores,
scoreThreshold, maxOverlap)``.  OpenCV is third-party (not vendored under
/root/reference); the published algorithm is ``dnn/src/nms.inl.hpp``
(``NMSFast_`` + ``GetMaxScoreIndex``) with ``rectOverlap = 1.f -
(float)jaccardDistance(a, b)`` on ``Rect_<int>``.  Pinned against the live
function in tests/test_oracle.py (random boxes, touching boxes, ties).
"""
import numpy as np


def _overlap(a, b):
    """``1.f - static_cast<float>(jaccardDistance(a, b))`` for int rects (x, y, w, h)."""
    area_a = a[2] * a[3]
    area_b = b[2] * b[3]
    if area_a + area_b <= 0:            # numeric_limits<int>::epsilon() == 0
        return np.float32(1.0)
    x1 = max(a[0], b[0])
    y1 = max(a[1], b[1])
    x2 = min(a[0] + a[2], b[0] + b[2])
    y2 = min(a[1] + a[3], b[1] + b[3])
    inter = float((x2 - x1) * (y2 - y1)) if (x2 > x1 and y2 > y1) else 0.0
    dist = 1.0 - inter / (float(area_a + area_b) - inter)       # double
    return np.float32(1.0) - np.float32(dist)


def nms_boxes(boxes, scores, score_threshold, nms_threshold, limit=None):
    """Indices kept, in descending-score order (ties keep input order)."""
    scores32 = np.asarray(scores, dtype=np.float32)
    thr = np.float32(score_threshold)
    nms_thr = np.float32(nms_threshold)
    cand = [i for i in range(len(scores32)) if scores32[i] > thr]
    cand.sort(key=lambda i: -float(scores32[i]))        # Python's sort is stable
    keep = []
    for i in cand:
        ok = True
        for k in keep:
            if not (_overlap(boxes[i], boxes[k]) <= nms_thr):
                ok = False
                break
        if ok:
            keep.append(i)
            if limit is not None and len(keep) >= limit:
                break
    return keep
