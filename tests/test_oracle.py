"""CPU tests: pin the oracle (restated algorithms) against the golden vectors made by
the unmodified reference, the reference's own known answers, and live cv2/scipy."""
import os

import cv2
import numpy as np
import pytest

from helpers import assert_hits_equal


def _hits_from_json(rows):
    return [(r[0], tuple(r[1]), r[2]) for r in rows]


def test_area_downscale_matches_cv2_inter_area():
    from oracle import golden_cases as gc
    fish = gc.fish()
    for f in (4, 8):
        want = cv2.resize(fish, (2048 // f, 2048 // f), interpolation=cv2.INTER_AREA)
        assert np.array_equal(gc.area_downscale(fish, f), want)


@pytest.mark.parametrize("name", ["t3_downscaled", "c1_fish256_n1", "c1_fish256_inf", "fish512_multi",
                                  "fish512_multi_n3", "synth_rot8", "synth_mixed", "synth_mixed_n5",
                                  "synth_searchbox", "synth_exact_fit", "t3_searchbox"])
def test_port_reproduces_reference_match(golden, name):
    from oracle import golden_cases as gc, mtm_port
    kind, temps, img, kw = gc.build(name)
    got = mtm_port.match_templates(temps, img, **kw)
    assert_hits_equal(got, _hits_from_json(golden[name]), tol=1e-6)
    if name in gc.NOTEBOOK_ANSWERS:                      # stored notebook outputs (cv2 4.7 -> 4.13: ~1e-6)
        nb = gc.NOTEBOOK_ANSWERS[name]
        assert got[0][:2] == nb[0][:2] and abs(float(got[0][2]) - nb[0][2]) < 1e-5


@pytest.mark.parametrize("name", ["c1_fish256_find", "fish512_find", "synth_row_map", "synth_col_map"])
def test_port_reproduces_reference_find(golden, name):
    from oracle import golden_cases as gc, mtm_port
    kind, temps, img, kw = gc.build(name)
    got = mtm_port.find_matches(temps, img, **kw)
    assert_hits_equal(got, _hits_from_json(golden[name]), tol=1e-6, ordered=False)


def test_port_nms_demo(golden):
    from oracle import mtm_port
    demo = [("1", (780, 350, 700, 480), 0.8), ("1", (806, 416, 716, 442), 0.6), ("1", (1074, 530, 680, 390), 0.4)]
    got = mtm_port.nms(demo, scoreThreshold=0.3, sortAscending=False, maxOverlap=0.5, N_object=2)
    assert [g[1] for g in got] == [tuple(w[1]) for w in golden["nms_demo"]] == [demo[0][1], demo[2][1]]


def test_exact_ncc_matches_reference_map(golden):
    """Exact restatement vs the reference's own cv2 map (C1): cv2's fp32 noise only."""
    from oracle import golden_cases as gc, ncc_exact
    kind, temps, img, kw = gc.build("c1_fish256_map")
    ref = np.load(gc.GOLDEN_DIR + "/c1_fish256_map.npy")
    exact = ncc_exact.match_template_exact(img, temps[0][1])
    assert np.max(np.abs(exact - ref)) < 5e-5
    assert int(exact.argmax()) == golden["c1_fish256_map"]["argmax"]


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4, 5])
def test_exact_ncc_all_methods_vs_cv2(method):
    from oracle import ncc_exact, synth
    rng = np.random.default_rng(method)
    tmpl = synth.make_template(rng, 19, 27)
    img, _ = synth.make_scene(90, 110, [tmpl], 2, seed=method)
    exact = ncc_exact.match_template_exact(img, tmpl, method=method, use_fft=False)
    cv = cv2.matchTemplate(img, tmpl, method)
    scale = max(1.0, float(np.abs(cv).max()))
    assert np.max(np.abs(exact.astype(np.float64) - cv)) <= 2e-5 * scale


def test_exact_ncc_rgb_and_fft_twin():
    from oracle import ncc_exact
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (70, 95, 3), dtype=np.uint8)
    tmpl = rng.integers(0, 256, (12, 21, 3), dtype=np.uint8)
    assert np.array_equal(ncc_exact.cc_direct(img, tmpl), ncc_exact.cc_fft(img, tmpl))
    exact = ncc_exact.match_template_exact(img, tmpl, use_fft=False)
    assert np.max(np.abs(exact - cv2.matchTemplate(img, tmpl, cv2.TM_CCOEFF_NORMED))) < 2e-5
    g = rng.integers(0, 256, (200, 230), dtype=np.uint8)
    t = rng.integers(0, 256, (64, 50), dtype=np.uint8)
    assert np.array_equal(ncc_exact.cc_direct(g, t), ncc_exact.cc_fft(g, t))


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4, 5])
def test_exact_ncc_float32_vs_cv2(method):
    """float32 branch of the dtype policy (MTM/__init__.py:71-74): oracle in float64 vs cv2's float64 DFT."""
    from oracle import ncc_exact
    rng = np.random.default_rng(100 + method)
    img = (rng.random((80, 100)) * 4000 + 3000).astype(np.float32)
    tmpl = (img[20:44, 30:61] + rng.normal(0, 50, (24, 31))).astype(np.float32)
    exact = ncc_exact.match_template_exact(img, tmpl, method=method, use_fft=False)
    cv = cv2.matchTemplate(img, tmpl, method)
    scale = max(1.0, float(np.abs(cv).max()))
    assert np.max(np.abs(exact.astype(np.float64) - cv)) <= 2e-5 * scale


@pytest.mark.parametrize("method", [0, 3])
def test_masked_restatement_vs_cv2(method):
    """OpenCV matchTemplateMask (uint8 binary masks, float32 weight masks, 1 and 3 channels)."""
    from oracle import ncc_exact
    rng = np.random.default_rng(200 + method)
    for shape_c in ((), (3,)):
        img = rng.integers(0, 256, (50, 64) + shape_c, dtype=np.uint8)
        t = rng.integers(0, 256, (11, 15) + shape_c, dtype=np.uint8)
        m8 = (rng.random((11, 15) + shape_c) > 0.3).astype(np.uint8) * 255
        got = ncc_exact.match_template_masked_exact(img, t, m8, method)
        cv = cv2.matchTemplate(img, t, method, mask=m8)
        assert np.max(np.abs(got - cv)) <= 2e-6 * max(1.0, float(np.abs(cv).max()))
        mf = rng.random((11, 15) + shape_c).astype(np.float32)
        got = ncc_exact.match_template_masked_exact(img.astype(np.float32), t.astype(np.float32), mf, method)
        cv = cv2.matchTemplate(img.astype(np.float32), t.astype(np.float32), method, mask=mf)
        assert np.max(np.abs(got - cv)) <= 2e-6 * max(1.0, float(np.abs(cv).max()))


def test_exact_ncc_degenerate_rules():
    from oracle import ncc_exact
    img = np.full((30, 40), 9, np.uint8)
    img[5:12, 5:12] = 180
    assert np.all(ncc_exact.match_template_exact(img, np.full((6, 6), 3, np.uint8)) == 1.0)
    t = np.arange(36, dtype=np.uint8).reshape(6, 6)
    m = ncc_exact.match_template_exact(img, t)
    cv = cv2.matchTemplate(img, t, cv2.TM_CCOEFF_NORMED)
    assert m[20, 30] == 0.0 and cv[20, 30] == 0.0 and np.isfinite(m).all()


def test_nms_port_matches_cv2():
    from oracle import nms_port
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(1, 60))
        boxes = [(int(rng.integers(0, 100)), int(rng.integers(0, 100)), int(rng.integers(1, 40)), int(rng.integers(1, 40)))
                 for _ in range(n)]
        scores = [float(np.float32(rng.integers(0, 20) / 20.0)) for _ in range(n)]
        thr, ov = float(rng.integers(0, 10)) / 10, float(rng.integers(0, 9)) / 8
        want = list(cv2.dnn.NMSBoxes(boxes, scores, thr, ov))
        assert nms_port.nms_boxes(boxes, scores, thr, ov) == [int(i) for i in want]
    # touching boxes are kept at maxOverlap=0, a 1-px overlap is suppressed
    assert nms_port.nms_boxes([(0, 0, 10, 10), (10, 0, 10, 10), (9, 0, 10, 10)], [0.9, 0.8, 0.7], 0.1, 0.0) == [0, 1]


def test_peak_finders_restated():
    from scipy.signal import find_peaks
    from oracle import peaks
    rng = np.random.default_rng(1)
    for _ in range(100):
        x = rng.integers(0, 6, int(rng.integers(1, 40))).astype(np.float32) / 5
        h = float(rng.integers(0, 6)) / 5
        assert np.array_equal(peaks.find_peaks_height(x, h), find_peaks(x, height=h)[0])
    m = np.zeros((5, 6), np.float32)
    assert len(peaks.peak_local_max(m, -1.0)) == 0                       # constant map -> no peaks
    m[2, 3] = 0.9; m[0, 0] = 0.9; m[4, 5] = 0.5; m[2, 4] = 0.9          # plateau (2,3)-(2,4), border peaks
    got = peaks.peak_local_max(m, 0.4).tolist()
    assert got == [[0, 0], [2, 3], [2, 4], [4, 5]]
    assert peaks.peak_local_max(m, 0.5).tolist() == [[0, 0], [2, 3], [2, 4]]   # strict threshold


# ---- augmentation / pyramid restatements ---------------------------------------------------
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("channels", [1, 3, 4])
def test_area_downscale_restatement_equals_live_cv2(dtype, channels):
    """oracle.augment_port.area_downscale == cv2.resize(..., INTER_AREA) bit for bit (integer factors 2..16,
    odd output sizes, ties included) -- the rule transform.cu implements."""
    from oracle import augment_port as ap
    rng = np.random.default_rng(5)
    top = 255 if dtype == np.uint8 else 65535
    for f in range(2, 17):
        h, w = 23, 37
        shape = (h * f + f - 1, w * f + 1) + ((channels,) if channels > 1 else ())      # remainders are cropped
        img = rng.integers(0, top + 1, shape).astype(dtype)
        img[:f, :f] = top                                                                # saturated box
        got = ap.area_downscale(img, f)
        want = ap.cv_area_downscale(img, f)
        assert got.dtype == want.dtype and got.shape == want.shape
        assert np.array_equal(got, want), "factor %d: %d pixels differ" % (f, int((got != want).sum()))
    # narrow images take OpenCV's scalar tail: same rule
    for ow in range(1, 20):
        img = rng.integers(0, top + 1, (6, 2 * ow) + ((channels,) if channels > 1 else ())).astype(dtype)
        assert np.array_equal(ap.area_downscale(img, 2), ap.cv_area_downscale(img, 2))


def test_area_downscale_float32_close_to_live_cv2():
    from oracle import augment_port as ap
    rng = np.random.default_rng(6)
    for f in (2, 3, 4, 7, 8):
        img = (rng.random((f * 31, f * 45)) * 255).astype(np.float32)
        assert np.max(np.abs(ap.area_downscale(img, f) - ap.cv_area_downscale(img, f))) <= 1e-6 * 255


def test_expand_templates_equals_product_host_expansion(mtm):
    from oracle import augment_port as ap
    rng = np.random.default_rng(8)
    temps = [("a", rng.integers(0, 256, (5, 9), dtype=np.uint8)), ("b", rng.integers(0, 256, (7, 4, 3), dtype=np.uint8))]
    names = list(ap.HOST_TRANSFORMS)
    assert sorted(names) == sorted(mtm.TRANSFORMS)
    got = mtm.expandTemplates(temps, names)
    want = ap.expand_templates(temps, names)
    assert [g[0] for g in got] == [w[0] for w in want]
    assert all(np.array_equal(g[1], w[1]) for g, w in zip(got, want))
    assert [g[0] for g in got[:3]] == ["a", "a_rot90", "a_rot180"]
    # Tutorial2 cell 15: rotated = np.rot90(temp0, k=i+1)
    assert np.array_equal(got[1][1], np.rot90(temps[0][1], 1)) and np.array_equal(got[2][1], np.rot90(temps[0][1], 2))


def test_pyramid_port_reproduces_notebook_answers():
    """The coarse-to-fine specification gives the full-resolution answer stored in Tutorial3 cell 10, and its
    unrefined form on the notebook's own small pair gives cell 21."""
    from oracle import augment_port as ap, golden_cases as gc
    fish = gc.fish()
    head = [("head", fish[842:842 + 184, 528:528 + 196])]
    for f in (2, 4, 8):
        got = ap.match_templates_pyramid(head, fish, downscale=f, N_object=1)
        assert [(h[0], h[1]) for h in got] == [(h[0], h[1]) for h in gc.NOTEBOOK_ANSWERS["t3_full"]]
        assert abs(float(got[0][2]) - 1.0) <= 1e-5
    small = ap.cv_area_downscale(fish, 4)
    assert np.array_equal(small, gc.area_downscale(fish, 4))
    got = ap.match_templates_pyramid([("downsampled", small[210:210 + 46, 131:131 + 49])], small, downscale=1, N_object=1,
                                     refine=False)
    want = gc.NOTEBOOK_ANSWERS["t3_downscaled"]
    assert (got[0][0], got[0][1]) == (want[0][0], want[0][1]) and abs(float(got[0][2]) - want[0][2]) <= 1e-5


def test_pyramid_port_equals_full_search_on_planted_scenes():
    from oracle import augment_port as ap, mtm_port, synth
    rng = np.random.default_rng(3)
    temps = [synth.make_template(rng, 64, 64), synth.make_template(rng, 48, 80)]
    img, _ = synth.make_scene(600, 800, temps, 4, seed=3)
    labelled = [("a", temps[0]), ("b", temps[1])]
    full = mtm_port.match_templates(labelled, img, score_threshold=0.5, maxOverlap=0.25)
    for f in (2, 4):
        pyr = ap.match_templates_pyramid(labelled, img, downscale=f, score_threshold=0.5, maxOverlap=0.25)
        assert [(h[0], h[1]) for h in pyr] == [(h[0], h[1]) for h in full]


def _f3_golden():
    import json
    from oracle import golden_cases as gc
    with open(os.path.join(gc.GOLDEN_DIR, "ref_outputs_f3.json")) as f:
        return json.load(f)


def test_f3_port_reproduces_the_unmodified_reference():
    """oracle/augment_port.py on the CPU port == the same tutorial user code on the UNMODIFIED reference's
    matchTemplates / NMS (tests/golden/ref_outputs_f3.json, made by `python -m oracle.make_golden --f3`)."""
    from oracle import augment_port as ap, golden_cases as gc
    golden = _f3_golden()
    for name in gc.F3_CASES:
        case = gc.build_f3(name)
        if case[0] == "aug":
            _, temps, transforms, img, kw = case
            got = ap.match_templates_augmented(temps, img, transforms, **kw)
        else:
            _, temps, img, f, refine, kw = case
            got = ap.match_templates_pyramid(temps, img, downscale=f, refine=refine, **kw)
        want = golden[name]
        assert [(g[0], tuple(int(v) for v in g[1])) for g in got] == [(w[0], tuple(w[1])) for w in want], name
        assert max(abs(float(g[2]) - w[2]) for g, w in zip(got, want)) <= 1e-6, name
    assert golden["pyr_fish_f4"][0][:2] == ["head", [528, 842, 196, 184]]          # Tutorial3 cell 10, through the reference itself
