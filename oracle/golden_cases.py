"""Recipes of the golden cases (TEST INFRASTRUCTURE).

Each case is reproducible from ``tests/golden/fish_2048.npz`` (the reference's
only offline data fixture, ``images/Fish.tif``, stored losslessly) or from the
seeded generator in ``oracle/synth.py``.  ``oracle/make_golden.py`` runs the
UNMODIFIED reference on every case in the build container and stores its
outputs in ``tests/golden/ref_outputs.json``; tests rebuild the inputs from the
recipe and compare.

Known answers owned by the reference itself (not just regenerated):
* ``t3_full`` / ``t3_searchbox`` / ``t3_downscaled`` -- stored notebook outputs
  ``tutorials/Tutorial3-SpeedingUp.ipynb`` cells 10 / 14 / 21;
* ``nms_demo`` -- ``MTM/NMS.py:86-96``.
"""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> stored notebook answer [(label, (x, y, w, h), score)]
NOTEBOOK_ANSWERS = {
    "t3_full": [("head", (528, 842, 196, 184), 1.0)],
    "t3_searchbox": [("head", (528, 842, 196, 184), 0.9999996)],
    "t3_downscaled": [("downsampled", (131, 210, 49, 46), 0.99999875)],
}

_fish = None


def fish():
    global _fish
    if _fish is None:
        _fish = np.load(os.path.join(GOLDEN_DIR, "fish_2048.npz"))["fish"]
    return _fish


def area_downscale(img, factor):
    """``cv2.resize(img, (W/f, H/f), interpolation=cv2.INTER_AREA)`` for an integer
    factor on uint8: rint of the f x f box mean (verified equal in tests/test_oracle.py)."""
    H, W = img.shape
    blocks = img.reshape(H // factor, factor, W // factor, factor).astype(np.float64)
    return np.rint(blocks.mean(axis=(1, 3))).astype(np.uint8)


def _synth_small(seed, H, W, sizes, n_plant, rot=False):
    from . import synth
    rng = np.random.default_rng(seed)
    temps = [synth.make_template(rng, h, w) for (h, w) in sizes]
    if rot:
        temps = [np.ascontiguousarray(np.rot90(t, k)) for t in temps for k in range(4)]
    image, _ = synth.make_scene(H, W, temps, n_plant, seed)
    return image, [("t%02d" % i, t) for i, t in enumerate(temps)]


def build(name):
    """Returns ``(kind, templates, image, kwargs)``; kind in {"match", "find", "map"}."""
    inf = float("inf")
    if name == "t3_full":
        img = fish()
        return "match", [("head", img[842:842 + 184, 528:528 + 196])], img, dict(N_object=1, method=5)
    if name == "t3_searchbox":
        img = fish()
        return "match", [("head", img[842:842 + 184, 528:528 + 196])], img, dict(
            N_object=1, method=5, searchBox=(76, 781, 1856, 353))
    if name == "t3_downscaled":
        small = area_downscale(fish(), 4)
        return "match", [("downsampled", small[210:210 + 46, 131:131 + 49])], small, dict(N_object=1, method=5)
    if name in ("c1_fish256_n1", "c1_fish256_inf", "c1_fish256_find", "c1_fish256_map"):
        small = area_downscale(fish(), 8)
        tmpl = [("head", small[84:148, 46:110])]
        if name == "c1_fish256_n1":
            return "match", tmpl, small, dict(N_object=1, method=5)
        if name == "c1_fish256_inf":
            return "match", tmpl, small, dict(N_object=inf, score_threshold=0.3, maxOverlap=0.25, method=5)
        if name == "c1_fish256_find":
            return "find", tmpl, small, dict(N_object=inf, score_threshold=0.3, method=5)
        return "map", tmpl, small, dict(method=5)
    if name in ("fish512_multi", "fish512_multi_n3", "fish512_find"):
        small = area_downscale(fish(), 4)
        temps = [("head", small[210:256, 131:180]), ("tail", small[230:262, 300:360]),
                 ("yolk", small[236:276, 190:240])]
        if name == "fish512_multi":
            return "match", temps, small, dict(N_object=inf, score_threshold=0.4, maxOverlap=0.3, method=5)
        if name == "fish512_multi_n3":
            return "match", temps, small, dict(N_object=3, score_threshold=0.3, maxOverlap=0.0, method=5)
        return "find", temps, small, dict(N_object=inf, score_threshold=0.5, method=5)
    if name == "synth_rot8":
        img, temps = _synth_small(11, 270, 480, [(32, 32), (32, 32)], 3, rot=True)
        return "match", temps, img, dict(N_object=inf, score_threshold=0.5, maxOverlap=0.25, method=5)
    if name == "synth_mixed":
        img, temps = _synth_small(12, 300, 420, [(24, 40), (33, 17), (48, 48), (20, 20), (64, 31)], 3)
        return "match", temps, img, dict(N_object=inf, score_threshold=0.45, maxOverlap=0.25, method=5)
    if name == "synth_mixed_n5":
        img, temps = _synth_small(12, 300, 420, [(24, 40), (33, 17), (48, 48), (20, 20), (64, 31)], 3)
        return "match", temps, img, dict(N_object=5, score_threshold=0.45, maxOverlap=0.1, method=5)
    if name == "synth_searchbox":
        img, temps = _synth_small(13, 256, 384, [(32, 32), (40, 24)], 4)
        return "match", temps, img, dict(N_object=inf, score_threshold=0.5, maxOverlap=0.25, method=5,
                                         searchBox=(37, 21, 301, 199))
    if name == "synth_exact_fit":
        # template exactly as large as the searchBox -> 1x1 score map (test.py:40-42)
        img, temps = _synth_small(14, 128, 160, [(40, 56)], 1)
        return "match", temps, img, dict(N_object=inf, score_threshold=-1.0, maxOverlap=0.25, method=5,
                                         searchBox=(10, 12, 56, 40))
    if name == "synth_row_map":
        # template as tall as the image -> 1 x n score map -> scipy find_peaks branch
        img, temps = _synth_small(15, 48, 400, [(48, 30)], 3)
        return "find", temps, img, dict(N_object=inf, score_threshold=0.2, method=5)
    if name == "synth_col_map":
        img, temps = _synth_small(16, 400, 40, [(26, 40)], 3)
        return "find", temps, img, dict(N_object=inf, score_threshold=0.2, method=5)
    raise KeyError(name)


def build_f3(name):
    """Recipes of the SURVEY 8 f3 golden cases (tutorial user code run with the unmodified reference).
    Returns ("aug", templates, transforms, image, kwargs) or ("pyr", templates, image, downscale, refine, kwargs)."""
    from . import synth
    inf = float("inf")
    if name in ("aug_rot4", "aug_flips_n3"):
        rng = np.random.default_rng(31)
        base = [synth.make_template(rng, 24, 40), synth.make_template(rng, 32, 32)]
        planted = [np.ascontiguousarray(np.rot90(base[0], 1)), base[1], np.ascontiguousarray(np.rot90(base[1], 2)), base[0]]
        img, _ = synth.make_scene(300, 420, planted, 2, seed=31)
        temps = [("a", base[0]), ("b", base[1])]
        if name == "aug_rot4":
            return "aug", temps, ("identity", "rot90", "rot180", "rot270"), img, dict(
                method=5, N_object=inf, score_threshold=0.5, maxOverlap=0.25)
        return "aug", temps, ("fliplr", "identity", "flipud", "transpose", "antitranspose"), img, dict(
            method=5, N_object=3, score_threshold=0.4, maxOverlap=0.1)
    if name in ("pyr_f4_refined", "pyr_f4_coarse", "pyr_f3_n5", "pyr_f2_sqdiff_n1"):
        rng = np.random.default_rng(3)
        ts = [synth.make_template(rng, 64, 64), synth.make_template(rng, 48, 80)]
        img, _ = synth.make_scene(600, 800, ts, 4, seed=3)
        temps = [("a", ts[0]), ("b", ts[1])]
        if name == "pyr_f4_refined":
            return "pyr", temps, img, 4, True, dict(method=5, N_object=inf, score_threshold=0.5, maxOverlap=0.25)
        if name == "pyr_f4_coarse":
            return "pyr", temps, img, 4, False, dict(method=5, N_object=inf, score_threshold=0.4, maxOverlap=0.25)
        if name == "pyr_f3_n5":
            return "pyr", temps, img, 3, True, dict(method=5, N_object=5, score_threshold=0.5, maxOverlap=0.1)
        return "pyr", temps, img, 2, True, dict(method=1, N_object=1, score_threshold=0.3, maxOverlap=0.25)
    if name == "pyr_fish_f4":
        img = fish()
        return "pyr", [("head", img[842:842 + 184, 528:528 + 196])], img, 4, True, dict(
            method=5, N_object=1, score_threshold=0.5, maxOverlap=0.25)
    raise KeyError(name)


F3_CASES = ["aug_rot4", "aug_flips_n3", "pyr_f4_refined", "pyr_f4_coarse", "pyr_f3_n5", "pyr_f2_sqdiff_n1", "pyr_fish_f4"]

CASES = ["t3_full", "t3_searchbox", "t3_downscaled",
         "c1_fish256_n1", "c1_fish256_inf", "c1_fish256_find", "c1_fish256_map",
         "fish512_multi", "fish512_multi_n3", "fish512_find",
         "synth_rot8", "synth_mixed", "synth_mixed_n5", "synth_searchbox",
         "synth_exact_fit", "synth_row_map", "synth_col_map"]
