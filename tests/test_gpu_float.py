"""-m gpu: the float32 branch of MTM's dtype policy (MTM/__init__.py:67-74) -- uint16 / float32
images and templates are cast to float32 by the host layer and matched by the fp32 CUDA path.
Oracle: exact float64 restatement and live cv2 (which uses a float64 DFT for CV_32F inputs)."""
import numpy as np
import pytest

from helpers import assert_hits_equal, assert_map_close

pytestmark = pytest.mark.gpu


def _scene16(seed, H=200, W=260, sizes=((32, 32), (24, 40))):
    from oracle import synth
    rng = np.random.default_rng(seed)
    temps8 = [synth.make_template(rng, h, w) for (h, w) in sizes]
    img8, _ = synth.make_scene(H, W, temps8, 3, seed=seed)
    # 16-bit microscopy-like data: scale to 12 bits over a large pedestal
    img = (img8.astype(np.uint16) * 13 + 3000).astype(np.uint16)
    temps = [(t.astype(np.uint16) * 13 + 3000).astype(np.uint16) for t in temps8]
    return img, temps


@pytest.mark.parametrize("method", [5, 3, 1, 4, 2, 0])
def test_float_score_maps_all_methods(mtm, method):
    import cv2
    from oracle import ncc_exact
    img, temps = _scene16(3)
    imgf, tf = img.astype(np.float32), temps[0].astype(np.float32)
    got = mtm.computeScoreMap(temps[0], img, method=method)          # uint16 in -> float32 policy
    assert got.dtype == np.float32
    exact = ncc_exact.match_template_exact(imgf, tf, method=method, use_fft=False)
    cv = cv2.matchTemplate(imgf, tf, method)
    scale = max(1.0, float(np.abs(exact).max()))
    assert np.max(np.abs(got.astype(np.float64) - exact)) <= 1e-4 * scale
    assert np.max(np.abs(got.astype(np.float64) - cv)) <= 1e-4 * scale + np.max(np.abs(cv - exact))


def test_float_inputs_mixed_dtypes_and_rgb(mtm):
    import cv2
    from oracle import ncc_exact
    rng = np.random.default_rng(5)
    img = rng.random((90, 120)).astype(np.float32) * 100 + 20
    t = (img[30:52, 40:75] + rng.normal(0, 2, (22, 35))).astype(np.float32)
    assert_map_close(mtm.computeScoreMap(t, img), ncc_exact.match_template_exact(img, t, use_fft=False),
                     cv=cv2.matchTemplate(img, t, cv2.TM_CCOEFF_NORMED))
    # uint8 template on a float32 image -> both become float32 (MTM/__init__.py:71-74)
    img8 = rng.integers(0, 256, (80, 90), dtype=np.uint8)
    t8 = np.ascontiguousarray(img8[10:30, 20:44])
    got = mtm.computeScoreMap(t8, img8.astype(np.float32))
    assert_map_close(got, ncc_exact.match_template_exact(img8, t8, use_fft=False))
    rgb = (rng.random((70, 80, 3)) * 1000).astype(np.float32)
    trgb = np.ascontiguousarray(rgb[20:38, 25:50]) + rng.normal(0, 30, (18, 25, 3)).astype(np.float32)
    assert_map_close(mtm.computeScoreMap(trgb, rgb), ncc_exact.match_template_exact(rgb, trgb, use_fft=False),
                     cv=cv2.matchTemplate(rgb, trgb, cv2.TM_CCOEFF_NORMED))
    with pytest.raises(ValueError, match="64-bit"):
        mtm.computeScoreMap(t.astype(np.float64), img)


def test_float_match_templates_vs_port(mtm):
    from oracle import mtm_port
    img, temps = _scene16(9, 240, 320, ((32, 32), (24, 40), (48, 20)))
    labelled = [("t%d" % i, t) for i, t in enumerate(temps)]
    for kw in (dict(score_threshold=0.5, maxOverlap=0.25), dict(N_object=1), dict(N_object=4, score_threshold=0.3)):
        got = mtm.matchTemplates(labelled, img, **kw)
        want = mtm_port.match_templates(labelled, img, **kw)
        assert len(want) > 0
        assert_hits_equal(got, want)
