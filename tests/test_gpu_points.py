"""-m gpu parity tests of ncc_points.cu: tiny score maps of large uint8 templates (search boxes barely larger than the
template, the re-localisation step of matchTemplatesPyramid) against the exact restatement of cv2.matchTemplate, and
bit for bit against the dp4a kernel (same exact integer sums, same float64 epilogue)."""
import numpy as np
import pytest

from helpers import assert_hits_equal, assert_map_close

pytestmark = pytest.mark.gpu

CASES = [((140, 150), (128, 128)),          # 13 x 23 map; tensor-core range -> taken over from the tcgen05 branch
         ((300, 310), (300, 290)),          # beyond the 32-bit tensor range -> taken over from the dp4a branch
         ((150, 160, 3), (130, 140, 3)),    # RGB
         ((131, 135, 4), (128, 128, 4)),    # RGBA
         ((129, 700), (129, 128)),          # 1 x 573 map (find_peaks branch downstream)
         ((200, 200), (200, 200))]          # 1 x 1 map


def _pair(shape, tshape, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    tmpl = np.ascontiguousarray(img[:tshape[0], :tshape[1]] // 2 + rng.integers(0, 120, tshape, dtype=np.uint8))
    return img, tmpl


@pytest.mark.parametrize("shape,tshape", CASES)
def test_small_maps_of_large_templates_all_methods(mtm, shape, tshape):
    from mtm_b200 import _native
    from oracle import ncc_exact
    img, tmpl = _pair(shape, tshape, 17)
    dp4a = _native.Context(0)
    dp4a.set_path(_native.PATH_DIRECT)
    for method in range(6):
        got = mtm.computeScoreMap(tmpl, img, method)
        exact = ncc_exact.match_template_exact(img, tmpl, method=method, use_fft=False)
        assert_map_close(got, exact)
        same = mtm.computeScoreMap(tmpl, img, method, context=dp4a)
        assert np.array_equal(got, same), "method %d: differs from the dp4a kernel" % method
    dp4a.close()


def test_exact_fit_search_boxes_with_a_large_template(mtm):
    """MTM/__init__.py:140-144 + test.py:40-42 with a template large enough for the small-map kernel."""
    from oracle import mtm_port, synth
    rng = np.random.default_rng(23)
    temps = [synth.make_template(rng, 130, 140), synth.make_template(rng, 150, 128)]
    img, _ = synth.make_scene(400, 520, temps, 1, seed=23)
    labelled = [("a", temps[0]), ("b", temps[1])]
    full = mtm_port.match_templates(labelled, img, N_object=float("inf"), score_threshold=0.5, maxOverlap=0.25)
    assert len(full) == 2
    for hit in full:
        x, y, w, h = hit[1]
        one = [t for t in labelled if t[0] == hit[0]]
        for box in [(x, y, w, h), (max(0, x - 3), max(0, y - 2), w + 6, h + 5)]:
            box = (box[0], box[1], min(box[2], 520 - box[0]), min(box[3], 400 - box[1]))
            for kw in (dict(N_object=1), dict(score_threshold=0.3, maxOverlap=0.25)):
                got = mtm.matchTemplates(one, img, searchBox=box, **kw)
                want = mtm_port.match_templates(one, img, searchBox=box, **kw)
                assert_hits_equal(got, want)
