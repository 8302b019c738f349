"""-m gpu: masked template matching (the optional third element of MTM's template tuples,
MTM/__init__.py:76-88, 213-217; OpenCV matchTemplateMask for TM_SQDIFF / TM_CCORR_NORMED)."""
import warnings

import numpy as np
import pytest

from helpers import assert_hits_equal

pytestmark = pytest.mark.gpu


def _case(seed, dtype):
    from oracle import synth
    rng = np.random.default_rng(seed)
    temps = [synth.make_template(rng, 30, 34), synth.make_template(rng, 22, 22)]
    img, _ = synth.make_scene(150, 190, temps, 3, seed=seed)
    masks = []
    for t in temps:
        yy, xx = np.mgrid[:t.shape[0], :t.shape[1]]
        m = (((yy - t.shape[0] / 2) ** 2 + (xx - t.shape[1] / 2) ** 2) < (min(t.shape) / 2) ** 2)
        masks.append((m * 255).astype(np.uint8))
    if dtype == np.float32:
        img, temps = img.astype(np.float32), [t.astype(np.float32) for t in temps]
        masks = [(m / 255.0).astype(np.float32) * rng.random(m.shape).astype(np.float32) for m in masks]
    return img, temps, masks


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("method", [0, 3])
def test_masked_score_map(mtm, method, dtype):
    import cv2
    from oracle import ncc_exact
    img, temps, masks = _case(7, dtype)
    for t, m in zip(temps, masks):
        got = mtm.computeScoreMap(t, img, method=method, mask=m)
        exact = ncc_exact.match_template_masked_exact(img, t, m, method)
        cv = cv2.matchTemplate(img, t, method, mask=m)
        scale = max(1.0, float(np.abs(exact).max()))
        assert np.max(np.abs(got.astype(np.float64) - exact)) <= 1e-4 * scale
        assert np.max(np.abs(got.astype(np.float64) - cv)) <= 1e-4 * scale + np.max(np.abs(cv.astype(np.float64) - exact))


def test_masked_find_and_policy_warnings(mtm):
    from oracle import mtm_port
    img, temps, masks = _case(8, np.uint8)
    labelled = [("disc", temps[0], masks[0]), ("plain", temps[1])]           # mixed: one with, one without mask
    for kw in (dict(method=3, score_threshold=0.9), dict(method=3, N_object=1)):
        got = mtm.findMatches(labelled, img, **kw)
        want = mtm_port.find_matches(labelled, img, **kw)
        assert len(want) > 0
        assert_hits_equal(got, want, ordered=False)
    got = mtm.matchTemplates(labelled, img, method=3, score_threshold=0.9, maxOverlap=0.2)
    want = mtm_port.match_templates(labelled, img, method=3, score_threshold=0.9, maxOverlap=0.2)
    assert_hits_equal(got, want)
    got0 = mtm.findMatches(labelled, img, method=0, N_object=1)                # method 0: global minima
    want0 = mtm_port.find_matches(labelled, img, method=0, N_object=1)
    assert [(g[0], g[1]) for g in got0] == [(w[0], w[1]) for w in want0]
    with warnings.catch_warnings(record=True) as rec:                          # mask + method 5 -> warning, mask ignored
        warnings.simplefilter("always")
        a = mtm.matchTemplates(labelled, img, method=5, score_threshold=0.6)
        b = mtm.matchTemplates([(l[0], l[1]) for l in labelled], img, method=5, score_threshold=0.6)
    assert any("not supporting the use of Mask" in str(w.message) for w in rec)
    assert_hits_equal(a, b, tol=0)
    with warnings.catch_warnings(record=True) as rec:                          # wrong-shape mask -> warning + ignored
        warnings.simplefilter("always")
        m = mtm.computeScoreMap(temps[0], img, method=3, mask=masks[1])
    assert any("same dimension or bit depth" in str(w.message) for w in rec)
    assert np.allclose(m, mtm.computeScoreMap(temps[0], img, method=3), atol=1e-6)
