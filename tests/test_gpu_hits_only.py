"""-m gpu: hits-only searches (MODE 3 of the tcgen05 kernels: no score map is written; the peak search resolves the 3x3
maxima inside the list of above-threshold pixels, N_object == 1 takes the arg-max straight from the epilogue) must return
what the map-based route returns: the hit list the restated peak finders + NMS give on the product's OWN score maps
(``computeScoreMap`` still writes them: bitwise the same arithmetic), order and score bits included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _from_own_maps(mtm, temps, img, thr, overlap, n_object=float("inf")):
    """findMatches + NMS restated on the product's own maps (oracle/peaks.py, oracle/mtm_port.nms on live cv2)."""
    from oracle import mtm_port, peaks
    hits = []
    for name, t in temps:
        m = mtm.computeScoreMap(t, img)
        if n_object == 1:
            y, x = np.unravel_index(int(m.argmax()), m.shape)
            pts = [(y, x)]
        else:
            pts = peaks.peak_local_max(m, thr).tolist()
        for (y, x) in pts:
            hits.append((name, (int(x), int(y), t.shape[1], t.shape[0]), m[y, x]))
    return hits, mtm_port.nms(hits, thr, False, n_object, overlap)


def _ctx():
    from mtm_b200 import _native
    return _native.default_context()


@pytest.mark.parametrize("channels", [1, 3])
def test_hits_only_equals_the_map_route(mtm, channels):
    from oracle import synth
    rng = np.random.default_rng(31 + channels)
    sizes = [(40, 40), (33, 57), (40, 40), (64, 48), (25, 31)]            # mixed sizes: one mode-A group zero pads its members
    gray = [synth.make_template(rng, h, w) for h, w in sizes]
    scene, _ = synth.make_scene(600, 900, gray, 3, seed=31)
    if channels == 3:
        scene = np.stack([scene, np.roll(scene, 5, axis=1), 255 - scene], axis=2)
        temps = [("t%d" % k, np.ascontiguousarray(scene[60 * k + 10:60 * k + 10 + h, 100 * k + 7:100 * k + 7 + w])) for k, (h, w) in enumerate(sizes)]
    else:
        temps = [("t%d" % k, t) for k, t in enumerate(gray)]
    ctx = _ctx()
    for thr, overlap, nobj in ((0.5, 0.25, float("inf")), (0.3, 0.0, float("inf")), (0.5, 0.25, 3), (0.5, 0.25, 1)):
        before = ctx.counters()["hits_only_searches"]
        got_find = mtm.findMatches(temps, scene, score_threshold=thr, N_object=nobj)
        got = mtm.matchTemplates(temps, scene, score_threshold=thr, maxOverlap=overlap, N_object=nobj)
        # (>=: a hit list longer than the binding's buffer makes it repeat the call with a larger one)
        assert ctx.counters()["hits_only_searches"] >= before + 2, "the search did not take the hits-only kernels"
        want_find, want = _from_own_maps(mtm, temps, scene, thr, overlap, nobj)
        assert [(g[0], g[1]) for g in got_find] == [(w[0], w[1]) for w in want_find], (thr, nobj)
        assert all(g[2] == w[2] for g, w in zip(got_find, want_find))
        assert [(g[0], g[1]) for g in got] == [(w[0], w[1]) for w in want], (thr, overlap, nobj)
        assert all(g[2] == w[2] for g, w in zip(got, want)) and len(want) >= 1


def test_near_ties_keep_the_exact_order(mtm):
    """Two copies of a template, one of them with a single pixel off by one grey level: scores 1.0 and 1 - ~1e-6.  The
    fp32 epilogue must order them as the exact (float64) oracle does, with tol = 0 -- at the N_object cut a swap would
    change the returned box."""
    from oracle import ncc_exact, synth
    rng = np.random.default_rng(77)
    t = synth.make_template(rng, 48, 48)
    scene, _ = synth.make_scene(400, 700, [t], 0, seed=77)
    scene[50:98, 60:108] = t
    scene[250:298, 500:548] = t
    scene[270, 520] = np.uint8(int(scene[270, 520]) + (1 if scene[270, 520] < 255 else -1))
    exact = ncc_exact.match_template_exact(scene, t)
    a, b = float(exact[50, 60]), float(exact[250, 500])
    assert a == 1.0 and 0.0 < a - b < 2e-5
    temps = [("t", t)]
    got = mtm.matchTemplates(temps, scene, score_threshold=0.9, maxOverlap=0.0)
    assert [g[1][:2] for g in got[:2]] == [(60, 50), (500, 250)] and float(got[0][2]) > float(got[1][2])
    assert mtm.matchTemplates(temps, scene, score_threshold=0.9, maxOverlap=0.0, N_object=1)[0][1][:2] == (60, 50)
    # the same pair with the perturbed copy FIRST in raster order: the better one still wins (not "first occurrence")
    flipped = np.ascontiguousarray(scene[::-1, ::-1])
    tf = np.ascontiguousarray(t[::-1, ::-1])
    ex = ncc_exact.match_template_exact(flipped, tf)
    y, x = np.unravel_index(int(ex.argmax()), ex.shape)
    got1 = mtm.matchTemplates([("t", tf)], flipped, score_threshold=0.9, maxOverlap=0.0, N_object=1)
    assert got1[0][1][:2] == (int(x), int(y))
    # (live cv2 cannot referee this pair: its float32 numerator noise, ~1e-6, is as large as the gap -- the exact oracle does)
    got2 = mtm.matchTemplates([("t", tf)], flipped, score_threshold=0.9, maxOverlap=0.0)
    assert [g[1][:2] for g in got2[:2]] == [(700 - 60 - 48, 400 - 50 - 48), (700 - 500 - 48, 400 - 250 - 48)]
    assert ex[got2[0][1][1], got2[0][1][0]] > ex[got2[1][1][1], got2[1][1][0]]


def test_candidate_overflow_falls_back_to_the_maps(mtm):
    """More above-threshold pixels than the candidate list holds: the search is repeated with score maps (streaming peak
    pass); results equal the restated algorithm on the product's own maps."""
    from oracle import synth
    rng = np.random.default_rng(5)
    temps = [("a", synth.make_template(rng, 24, 24)), ("b", synth.make_template(rng, 20, 30))]
    scene, _ = synth.make_scene(400, 500, [t[1] for t in temps], 3, seed=5)
    got = mtm.matchTemplates(temps, scene, score_threshold=0.0, maxOverlap=0.3)        # about half of all pixels are candidates
    _, want = _from_own_maps(mtm, temps, scene, 0.0, 0.3)
    assert [(g[0], g[1]) for g in got] == [(w[0], w[1]) for w in want] and len(want) > 20
    assert all(g[2] == w[2] for g, w in zip(got, want))
    # and the context is healthy afterwards (the hash table of the list resolver is left empty)
    got = mtm.matchTemplates(temps, scene, score_threshold=0.5, maxOverlap=0.3)
    _, want = _from_own_maps(mtm, temps, scene, 0.5, 0.3)
    assert [(g[0], g[1]) for g in got] == [(w[0], w[1]) for w in want] and len(want) >= 4
