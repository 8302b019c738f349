"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Writes
* ``tests/golden/fish_2048.npz``   -- ``images/Fish.tif`` pixels (lossless, compressed);
* ``tests/golden/ref_outputs.json`` -- reference outputs for every recipe in
  ``oracle/golden_cases.py`` (hit lists; for "map" cases the float32 map goes to
  ``tests/golden/<case>.npy``);
* ``tests/golden/ref_outputs_f3.json`` -- the tutorials' augmentation / downscaling user code
  (``oracle/augment_port.py``) run with the UNMODIFIED reference's ``matchTemplates`` / ``NMS`` on the
  recipes of ``golden_cases.F3_CASES`` (``python -m oracle.make_golden --f3`` writes only this file);
* records cv2/scipy/numpy versions used.
TEST INFRASTRUCTURE; never imported by product code.
"""
import json
import os
import sys

import numpy as np


def _plain(hits):
    return [[str(l), [int(v) for v in b], float(np.float32(s))] for (l, b, s) in hits]


def make_f3():
    import cv2
    from . import augment_port as ap, golden_cases as gc, ref_loader
    ref = ref_loader.load()
    out = {"_meta": {"reference_version": ref.__version__, "cv2": cv2.__version__, "numpy": np.__version__,
                     "note": "tutorial user code (oracle/augment_port.py) on top of the unmodified reference's matchTemplates / NMS; "
                             "peak_local_max inside the reference run is oracle/peaks.py (scikit-image absent)"}}
    for name in gc.F3_CASES:
        case = gc.build_f3(name)
        if case[0] == "aug":
            _, temps, transforms, img, kw = case
            res = _plain(ap.reference_augmented(ref, temps, img, transforms, **kw))
        else:
            _, temps, img, f, refine, kw = case
            res = _plain(ap.match_templates_pyramid(temps, img, downscale=f, refine=refine, impl=ap.ReferenceImpl(ref), **kw))
        out[name] = res
        print(name, len(res), file=sys.stderr)
    with open(os.path.join(gc.GOLDEN_DIR, "ref_outputs_f3.json"), "w") as f:
        json.dump(out, f, indent=1)


def main():
    import cv2
    import scipy
    from . import golden_cases as gc, ref_loader
    os.makedirs(gc.GOLDEN_DIR, exist_ok=True)
    if "--f3" in sys.argv:
        make_f3()
        return
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
