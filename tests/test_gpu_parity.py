"""-m gpu parity tests: the CUDA path (through the C ABI) against the oracle and the
golden vectors made by the unmodified reference.  Bars: score maps within 1e-4 of
the exact restatement (north_star), hit lists identical after NMS."""
import numpy as np
import pytest

from helpers import assert_hits_equal, assert_map_close

pytestmark = pytest.mark.gpu


def _hits_from_json(rows):
    return [(r[0], tuple(r[1]), r[2]) for r in rows]


def test_score_map_c1_fish(mtm, golden):
    from oracle import golden_cases as gc, ncc_exact
    kind, temps, img, kw = gc.build("c1_fish256_map")
    got = mtm.computeScoreMap(temps[0][1], img, **kw)
    ref = np.load(gc.GOLDEN_DIR + "/c1_fish256_map.npy")          # unmodified reference (cv2)
    exact = ncc_exact.match_template_exact(img, temps[0][1])
    assert got.dtype == np.float32 and got.shape == (193, 193)
    assert_map_close(got, exact, cv=ref)
    assert int(got.argmax()) == golden["c1_fish256_map"]["argmax"]


@pytest.mark.parametrize("shape,tshape,seed", [
    ((97, 131), (16, 16), 0), ((64, 64), (64, 64), 1), ((150, 70), (31, 7), 2), ((80, 300), (5, 64), 3),
    ((200, 200), (1, 1), 4), ((129, 257), (33, 65), 5), ((70, 90), (70, 13), 6), ((90, 70), (13, 70), 7),
    ((300, 340), (100, 130), 8),
])
def test_score_map_random(mtm, shape, tshape, seed):
    import cv2
    from oracle import ncc_exact
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    tmpl = rng.integers(0, 256, tshape, dtype=np.uint8)
    got = mtm.computeScoreMap(tmpl, img)
    exact = ncc_exact.match_template_exact(img, tmpl, use_fft=False)
    cv = cv2.matchTemplate(img, tmpl, cv2.TM_CCOEFF_NORMED)
    assert_map_close(got, exact, cv=cv)


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4, 5])
def test_score_map_all_methods(mtm, method):
    from oracle import ncc_exact, synth
    rng = np.random.default_rng(40 + method)
    tmpl = synth.make_template(rng, 21, 34)
    img, _ = synth.make_scene(120, 160, [tmpl], 2, seed=40 + method)
    got = mtm.computeScoreMap(tmpl, img, method=method)
    exact = ncc_exact.match_template_exact(img, tmpl, method=method, use_fft=False)
    assert_map_close(got, exact, tol=1e-4)


def test_score_map_rgb(mtm):
    import cv2
    from oracle import ncc_exact
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (90, 120, 3), dtype=np.uint8)
    tmpl = np.ascontiguousarray(img[20:45, 30:71]) // 2 + rng.integers(0, 100, (25, 41, 3), dtype=np.uint8)
    got = mtm.computeScoreMap(tmpl.astype(np.uint8), img)
    exact = ncc_exact.match_template_exact(img, tmpl.astype(np.uint8), use_fft=False)
    cv = cv2.matchTemplate(img, tmpl.astype(np.uint8), cv2.TM_CCOEFF_NORMED)
    assert_map_close(got, exact, cv=cv)


def test_flat_and_constant_inputs(mtm):
    """OpenCV rules: constant template -> all ones; exactly flat window -> 0; never NaN."""
    img = np.full((40, 50), 7, np.uint8)
    img[10:20, 10:20] = 200
    const_t = np.full((8, 8), 31, np.uint8)
    assert np.all(mtm.computeScoreMap(const_t, img) == 1.0)
    t = np.arange(64, dtype=np.uint8).reshape(8, 8)
    m = mtm.computeScoreMap(t, img)
    assert np.isfinite(m).all() and m[0, 40] == 0.0 and m[30, 0] == 0.0


GOLDEN_MATCH = ["t3_downscaled", "c1_fish256_n1", "c1_fish256_inf", "fish512_multi", "fish512_multi_n3",
                "synth_rot8", "synth_mixed", "synth_mixed_n5", "synth_searchbox", "synth_exact_fit"]


@pytest.mark.parametrize("name", GOLDEN_MATCH)
def test_match_templates_golden(mtm, golden, name):
    from oracle import golden_cases as gc
    kind, temps, img, kw = gc.build(name)
    assert kind == "match"
    got = mtm.matchTemplates(temps, img, **kw)
    assert_hits_equal(got, _hits_from_json(golden[name]))
    for h in got:
        assert isinstance(h[0], str) and isinstance(h[2], np.float32) and all(isinstance(v, int) for v in h[1])


@pytest.mark.parametrize("name", ["t3_full", "t3_searchbox"])
def test_match_templates_tutorial3_fullres(mtm, golden, name):
    """Known answers owned by the reference: Tutorial3-SpeedingUp.ipynb cells 10 / 14."""
    from oracle import golden_cases as gc
    kind, temps, img, kw = gc.build(name)
    got = mtm.matchTemplates(temps, img, **kw)
    assert_hits_equal(got, _hits_from_json(golden[name]))
    nb = gc.NOTEBOOK_ANSWERS[name]
    assert got[0][0] == nb[0][0] and got[0][1] == nb[0][1] and abs(float(got[0][2]) - nb[0][2]) <= 1e-4


@pytest.mark.parametrize("name", ["c1_fish256_find", "synth_row_map", "synth_col_map"])
def test_find_matches_golden(mtm, golden, name):
    from oracle import golden_cases as gc
    kind, temps, img, kw = gc.build(name)
    got = mtm.findMatches(temps, img, **kw)
    assert_hits_equal(got, _hits_from_json(golden[name]), ordered=False)


def test_find_matches_real_image_vs_reference(mtm, golden):
    """fish512_find (149 raw peaks on a real image).  Raw peak SETS are not stable across
    numerator implementations (cv2's own IPP / non-IPP paths differ by a few plateau pixels,
    SURVEY.md 7.3(1)); every reference peak scoring well above the plateau noise must be
    found at the same place, and the counts must agree within a few percent."""
    from oracle import golden_cases as gc
    kind, temps, img, kw = gc.build("fish512_find")
    got = {(h[0], h[1]): float(h[2]) for h in mtm.findMatches(temps, img, **kw)}
    ref = _hits_from_json(golden["fish512_find"])
    assert abs(len(got) - len(ref)) <= max(3, len(ref) // 20)
    missing = [r for r in ref if (r[0], r[1]) not in got]
    assert len(missing) <= max(3, len(ref) // 20), missing
    for r in ref:
        if (r[0], r[1]) in got:
            assert abs(got[(r[0], r[1])] - r[2]) <= 1e-4


def test_peak_extraction_matches_oracle_on_own_maps(mtm):
    """K5/K6 in isolation: peaks of the product's own score maps (bitwise the same input)
    must equal the restated peak_local_max / minMaxLoc exactly, including the order."""
    from oracle import golden_cases as gc, peaks
    kind, temps, img, kw = gc.build("fish512_find")
    for thr in (0.5, 0.2, -1.0):
        got = mtm.findMatches(temps, img, score_threshold=thr)
        want = []
        for name, t in temps:
            m = mtm.computeScoreMap(t, img)
            for (y, x) in peaks.peak_local_max(m, thr).tolist():
                want.append((name, (int(x), int(y), t.shape[1], t.shape[0]), m[y, x]))
        assert [(g[0], g[1]) for g in got] == [(w[0], w[1]) for w in want]
        assert all(g[2] == w[2] for g, w in zip(got, want))
    got1 = mtm.findMatches(temps, img, N_object=1)
    for (name, t), g in zip(temps, got1):
        m = mtm.computeScoreMap(t, img)
        y, x = np.unravel_index(int(m.argmax()), m.shape)
        assert g[1][:2] == (int(x), int(y)) and g[2] == m[y, x]


def test_nms_demo_and_random(mtm, golden):
    from oracle import mtm_port
    demo = [("1", (780, 350, 700, 480), 0.8), ("1", (806, 416, 716, 442), 0.6), ("1", (1074, 530, 680, 390), 0.4)]
    got = mtm.NMS(demo, scoreThreshold=0.3, sortAscending=False, maxOverlap=0.5, N_object=2)
    assert [(g[0], tuple(g[1])) for g in got] == [(w[0], tuple(w[1])) for w in golden["nms_demo"]]
    rng = np.random.default_rng(5)
    for trial in range(20):
        n = int(rng.integers(2, 200))
        hits = [("L%d" % i, (int(rng.integers(0, 300)), int(rng.integers(0, 300)), int(rng.integers(1, 80)),
                             int(rng.integers(1, 80))), np.float32(rng.integers(0, 50) / 50.0)) for i in range(n)]
        for asc in (False, True):
            for nobj in (float("inf"), 1, 3, 0):
                kw = dict(scoreThreshold=0.3, sortAscending=asc, N_object=nobj, maxOverlap=float(rng.integers(0, 5)) / 8)
                assert mtm.NMS(hits, **kw) == mtm_port.nms(hits, **kw), (trial, kw)


@pytest.mark.parametrize("cfg", ["C2", "C4"])
def test_baseline_configs_full_size(mtm, cfg):
    """BASELINE.json configs at full size against the CPU port (live cv2)."""
    from oracle import mtm_port, synth
    image, temps, params = synth.config(cfg)
    got = mtm.matchTemplates(temps, image, **params)
    want = mtm_port.match_templates(temps, image, **params)
    assert len(want) > 0
    assert_hits_equal(got, want)


def test_baseline_config_c5_one_image(mtm):
    """C5 (one 3840x2160 image of the batch, 64 templates of 64 different sizes 32..128, N_object=50): the
    mixed-size tensor-core groups and the 12-warp epilogue variant against the CPU port; a second image of the
    batch through the pipelined entry point must equal its own synchronous call."""
    from oracle import mtm_port, synth
    image, temps, params = synth.config("C5")
    got = mtm.matchTemplates(temps, image, **params)
    want = mtm_port.match_templates(temps, image, **params)
    assert len(want) == 50
    assert_hits_equal(got, want)
    image1, _, _ = synth.config("C5", image_index=1)
    batch = mtm.matchTemplatesBatch(temps, [image, image1], **params)
    assert_hits_equal(batch[0], got, tol=0.0)
    assert_hits_equal(batch[1], mtm.matchTemplates(temps, image1, **params), tol=0.0)


def test_c3_large_template_properties(mtm):
    """C3 (4096^2, 256^2 template): too large for the direct oracle -> FFT-exact oracle on the map."""
    from oracle import ncc_exact, synth
    image, temps, params = synth.config("C3")
    got = mtm.computeScoreMap(temps[0][1], image)
    exact = ncc_exact.match_template_exact(image, temps[0][1], use_fft=True)
    assert_map_close(got, exact)


def test_validation_errors_match_reference(mtm):
    img = np.zeros((50, 60), np.uint8)
    t = np.zeros((10, 10), np.uint8)
    with pytest.raises(ValueError, match="Maximal overlap"):
        mtm.matchTemplates([("a", t)], img, maxOverlap=1.5)
    with pytest.raises(TypeError, match="N_object must be an integer"):
        mtm.matchTemplates([("a", t)], img, N_object=np.int64(2))
    with pytest.raises(ValueError, match="larger than image"):
        mtm.matchTemplates([("big", np.pad(img, 1))], img)
    with pytest.raises(ValueError, match="larger than searchBox"):
        mtm.matchTemplates([("a", t)], img, searchBox=(0, 0, 5, 5))
    with pytest.raises(ValueError, match="list of tuples"):
        mtm.matchTemplates([["a", t]], img)
    with pytest.raises(ValueError, match="TM_SQDIFF is not supported"):
        mtm.matchTemplates([("a", t)], img, method=0)
    with pytest.raises(ValueError, match="64-bit"):
        mtm.computeScoreMap(t.astype(np.float64), img)


@pytest.mark.parametrize("method,thr", [(1, 0.35), (3, 0.92), (2, 0.0), (4, 0.0)])
def test_match_templates_other_methods_vs_port(mtm, method, thr):
    """the other score methods through the whole pipeline (method 1 searches
    local MINIMA and sorts ascending in the NMS, MTM/__init__.py:232-233, MTM/NMS.py:73-75)."""
    from oracle import mtm_port, ncc_exact, synth
    rng = np.random.default_rng(60 + method)
    temps = [("a", synth.make_template(rng, 28, 28)), ("b", synth.make_template(rng, 20, 36))]
    img, _ = synth.make_scene(180, 240, [t[1] for t in temps], 3, seed=60 + method)
    if method in (2, 4):        # un-normalised scores: compare maps and the N_object=1 answer only
        for name, t in temps:
            assert_map_close(mtm.computeScoreMap(t, img, method=method),
                             ncc_exact.match_template_exact(img, t, method=method, use_fft=False))
        got = mtm.matchTemplates(temps, img, method=method, N_object=1)
        want = mtm_port.match_templates(temps, img, method=method, N_object=1)
        assert got[0][:2] == want[0][:2]
        return
    got = mtm.matchTemplates(temps, img, method=method, score_threshold=thr, maxOverlap=0.2)
    want = mtm_port.match_templates(temps, img, method=method, score_threshold=thr, maxOverlap=0.2)
    assert len(want) >= 2
    assert_hits_equal(got, want)
    got1 = mtm.matchTemplates(temps, img, method=method, N_object=1)
    want1 = mtm_port.match_templates(temps, img, method=method, N_object=1)
    assert_hits_equal(got1, want1)


def test_hit_buffer_growth_beyond_device_capacity(mtm):
    """> 65536 raw peaks: the device hit blocks are re-allocated and the search re-run (general
    multi-kernel sort/NMS path); results must still equal the restated peak finder on the same map."""
    from oracle import mtm_port, peaks
    rng = np.random.default_rng(123)
    img = rng.integers(0, 256, (820, 830), dtype=np.uint8)
    t = rng.integers(0, 256, (3, 3), dtype=np.uint8)
    got = mtm.findMatches([("n", t)], img, score_threshold=-1.0)
    m = mtm.computeScoreMap(t, img)
    want = peaks.peak_local_max(m, -1.0)
    assert len(got) == len(want) > 65536
    assert [(g[1][1], g[1][0]) for g in got] == [tuple(int(v) for v in p) for p in want.tolist()]
    # and through NMS (N_object finite so the scan stops early)
    top = mtm.matchTemplates([("n", t)], img, score_threshold=0.0, maxOverlap=0.0, N_object=40)
    ref = mtm_port.nms([("n", (int(p[1]), int(p[0]), 3, 3), m[p[0], p[1]]) for p in want.tolist()], 0.0, False, 40, 0.0)
    assert [(a[1]) for a in top] == [(b[1]) for b in ref]


def test_match_templates_batch_equals_loop(mtm):
    """Pipelined batch entry point == the per-image calls (incl. a >1024-raw-peak image that
    falls back to the synchronous path, and searchBox offsets)."""
    from oracle import synth
    rng = np.random.default_rng(77)
    temps = [("t%d" % i, synth.make_template(rng, 24 + 4 * i, 30)) for i in range(5)]
    images = [synth.make_scene(200, 260, [t[1] for t in temps], 2, seed=100 + k)[0] for k in range(11)]
    for kw in (dict(score_threshold=0.5, maxOverlap=0.25), dict(score_threshold=0.4, N_object=3),
               dict(N_object=1), dict(score_threshold=-0.0, maxOverlap=0.3), dict(score_threshold=0.5, searchBox=(10, 20, 180, 150))):
        want = [mtm.matchTemplates(temps, im, **kw) for im in images]
        got = mtm.matchTemplatesBatch(temps, images, **kw)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert [(a[0], a[1]) for a in g] == [(b[0], b[1]) for b in w]
            assert all(a[2] == b[2] for a, b in zip(g, w))
