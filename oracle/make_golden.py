"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Writes
* ``tests/golden/fish_2048.npz``   -- ``images/Fish.tif`` pixels (lossless, compressed);
* ``tests/golden/ref_outputs.json`` -- reference outputs for every recipe in
  ``oracle/golden_cases.py`` (hit lists; for "map" cases the float32 map goes to
  ``tests/golden/<case>.npy``);
* records cv2/scipy/numpy versions used.
TEST INFRASTRUCTURE; never imported by product code.
"""
import json
import os
import sys

import numpy as np


def _plain(hits):
    return [[str(l), [int(v) for v in b], float(np.float32(s))] for (l, b, s) in hits]


def main():
    import cv2
    import scipy
    from . import golden_cases as gc, ref_loader
    os.makedirs(gc.GOLDEN_DIR, exist_ok=True)
    fish_path = os.path.join(gc.GOLDEN_DIR, "fish_2048.npz")
    if not os.path.exists(fish_path):
        fish = cv2.imread(os.path.join(ref_loader.REF_ROOT, "images", "Fish.tif"), -1)
        assert fish.shape == (2048, 2048) and fish.dtype == np.uint8
        np.savez_compressed(fish_path, fish=fish)
    ref = ref_loader.load()
    out = {"_meta": {"reference_version": ref.__version__, "cv2": cv2.__version__,
                     "scipy": scipy.__version__, "numpy": np.__version__,
                     "note": "peak_local_max inside the reference run is oracle/peaks.py (scikit-image absent)"}}
    for name in gc.CASES:
        kind, temps, img, kw = gc.build(name)
        if kind == "match":
            res = _plain(ref.matchTemplates(temps, img, **kw))
        elif kind == "find":
            res = _plain(ref.findMatches(temps, img, **kw))
        else:
            m = ref.computeScoreMap(temps[0][1], img, **kw)
            np.save(os.path.join(gc.GOLDEN_DIR, name + ".npy"), m)
            res = {"shape": list(m.shape), "max": float(m.max()), "argmax": int(m.argmax())}
        out[name] = res
        print(name, kind, (len(res) if isinstance(res, list) else res), file=sys.stderr)
    # MTM/NMS.py:86-96 self-demo
    demo = [("1", (780, 350, 700, 480), 0.8), ("1", (806, 416, 716, 442), 0.6), ("1", (1074, 530, 680, 390), 0.4)]
    out["nms_demo"] = _plain(ref.NMS(demo, scoreThreshold=0.3, sortAscending=False, maxOverlap=0.5, N_object=2))
    with open(os.path.join(gc.GOLDEN_DIR, "ref_outputs.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
