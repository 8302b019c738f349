"""-m gpu tests of the tcgen05 (tensor-core) numerator kernel against the exact oracle
and against the dp4a direct kernel (both are exact-integer numerators: the maps may
differ only by the fp32-vs-fp64 normalisation, far below the 1e-4 bar)."""
import numpy as np
import pytest

from helpers import assert_hits_equal, assert_map_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctxs():
    import MTM  # noqa: F401
    from mtm_b200 import _native
    t = _native.Context(0)
    t.set_path(_native.PATH_TENSOR)
    d = _native.Context(0)
    d.set_path(_native.PATH_DIRECT)
    yield t, d
    t.close()
    d.close()


def _maps(mtm, ctx, temps, img):
    return [mtm.computeScoreMap(t, img, context=ctx) for t in temps]


@pytest.mark.parametrize("shape,tshape,n_t,seed", [
    ((200, 300), (64, 64), 8, 0),       # mode A, full group
    ((130, 170), (32, 32), 5, 1),       # mode A, partial group
    ((300, 310), (48, 37), 1, 2),       # mode B
    ((90, 140), (16, 16), 1, 3),        # mode B, small
    ((260, 420), (100, 130), 2, 4),     # two templates -> mode B twice or A padded
    ((70, 75), (64, 64), 3, 5),         # image barely larger than the template
    ((400, 200), (17, 90), 11, 6),      # 8 + 3, non-square
    ((430, 520), (200, 210), 2, 7),     # between the BASELINE sizes: the launch weighs the persistent kernel against one tile per CTA
])
def test_tensor_path_maps_vs_exact(mtm, ctxs, shape, tshape, n_t, seed):
    from oracle import ncc_exact, synth
    ct, cd = ctxs
    rng = np.random.default_rng(seed)
    temps = [synth.make_template(rng, *tshape) for _ in range(n_t)]
    img, _ = synth.make_scene(shape[0], shape[1], temps[:3], 2, seed=seed)
    labelled = [("t%d" % i, t) for i, t in enumerate(temps)]
    # whole-list call exercises the grouped launch; single calls exercise mode B
    with ct.lock:
        ct.set_image(img)
        ct.set_templates(temps)
        got = [ct.score_map(i, 5, (shape[0] - tshape[0] + 1, shape[1] - tshape[1] + 1)) for i in range(n_t)]
    for i, t in enumerate(temps):
        exact = ncc_exact.match_template_exact(img, t, use_fft=False)
        assert_map_close(got[i], exact)
        direct = mtm.computeScoreMap(t, img, context=cd)
        assert np.max(np.abs(got[i] - direct)) <= 2e-6
    hits_t = mtm.matchTemplates(labelled, img, score_threshold=0.5, context=ct)
    hits_d = mtm.matchTemplates(labelled, img, score_threshold=0.5, context=cd)
    assert_hits_equal(hits_t, hits_d, tol=2e-6)


def test_tensor_path_mixed_size_groups(mtm, ctxs):
    """C5-style template sets: different sizes share one tcgen05 launch (zero-padded Toeplitz slabs)."""
    from oracle import ncc_exact, synth
    ct, cd = ctxs
    rng = np.random.default_rng(21)
    sides = np.linspace(16, 52, 13).round().astype(int)
    temps = [synth.make_template(rng, int(s), int(s + (i % 3) * 5)) for i, s in enumerate(sides)]
    order = rng.permutation(len(temps))                       # list order != size order
    temps = [temps[i] for i in order]
    img, _ = synth.make_scene(230, 310, temps[:4], 2, seed=21)
    labelled = [("t%d" % i, t) for i, t in enumerate(temps)]
    with ct.lock:
        ct.set_image(img)
        ct.set_templates(temps)
        got = [ct.score_map(i, 5, (img.shape[0] - t.shape[0] + 1, img.shape[1] - t.shape[1] + 1)) for i, t in enumerate(temps)]
    for i, t in enumerate(temps):
        assert_map_close(got[i], ncc_exact.match_template_exact(img, t, use_fft=False))
    hits_t = mtm.matchTemplates(labelled, img, score_threshold=0.45, context=ct)
    hits_d = mtm.matchTemplates(labelled, img, score_threshold=0.45, context=cd)
    assert len(hits_d) > 0
    assert_hits_equal(hits_t, hits_d, tol=2e-6)


@pytest.mark.parametrize("C,n_t,tshape,seed", [(3, 8, (32, 32), 31), (3, 3, (40, 21), 32), (4, 2, (16, 48), 33), (3, 1, (64, 64), 34)])
def test_tensor_path_rgb(mtm, ctxs, C, n_t, tshape, seed):
    """Interleaved RGB / RGBA uint8 on the tcgen05 kernel (x-step = C bytes in the Toeplitz band,
    per-channel window sums in the epilogue) against the exact oracle, cv2 and the dp4a kernel."""
    import cv2
    from oracle import ncc_exact
    ct, cd = ctxs
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (150, 200, C), dtype=np.uint8)
    temps = []
    for k in range(n_t):
        y, x = int(rng.integers(0, 150 - tshape[0])), int(rng.integers(0, 200 - tshape[1]))
        t = img[y:y + tshape[0], x:x + tshape[1]].astype(np.int32) + rng.integers(-40, 40, tshape + (C,))
        temps.append(np.clip(t, 0, 255).astype(np.uint8))
    with ct.lock:
        ct.set_image(img)
        ct.set_templates(temps)
        got = [ct.score_map(i, 5, (150 - tshape[0] + 1, 200 - tshape[1] + 1)) for i in range(n_t)]
    for i, t in enumerate(temps):
        exact = ncc_exact.match_template_exact(img, t, use_fft=False)
        assert_map_close(got[i], exact, cv=cv2.matchTemplate(img, t, cv2.TM_CCOEFF_NORMED))
        assert np.max(np.abs(got[i] - mtm.computeScoreMap(t, img, context=cd))) <= 2e-6
    labelled = [("t%d" % i, t) for i, t in enumerate(temps)]
    assert_hits_equal(mtm.matchTemplates(labelled, img, score_threshold=0.5, context=ct),
                      mtm.matchTemplates(labelled, img, score_threshold=0.5, context=cd), tol=2e-6)


@pytest.mark.parametrize("C", [1, 3])
@pytest.mark.parametrize("method", [0, 1, 2, 3, 4])
def test_tensor_path_other_methods(mtm, ctxs, method, C):
    """cv2 methods 0..4 on the tcgen05 numerator (float64 OpenCV epilogue on the summed-area tables) against
    the exact oracle and the dp4a kernel, which shares the epilogue: identical numerators -> identical maps."""
    from oracle import ncc_exact
    ct, cd = ctxs                                           # ct is forced onto the tensor path: unsupported -> error
    rng = np.random.default_rng(40 + method)
    shape = (120, 170) if C == 1 else (120, 170, C)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    temps = []
    for k in range(3):
        y, x = int(rng.integers(0, 90)), int(rng.integers(0, 130))
        t = img[y:y + 24, x:x + 31].astype(np.int32) + rng.integers(-30, 30, (24, 31) + shape[2:])
        temps.append(np.clip(t, 0, 255).astype(np.uint8))
    for t in temps:
        got_t = mtm.computeScoreMap(t, img, method=method, context=ct)
        got_d = mtm.computeScoreMap(t, img, method=method, context=cd)
        exact = ncc_exact.match_template_exact(img, t, method=method, use_fft=False)
        assert_map_close(got_t, exact)
        assert np.array_equal(got_t, got_d)
    if method != 0:                                         # matchTemplates rejects TM_SQDIFF like the reference
        labelled = [("t%d" % i, t) for i, t in enumerate(temps)]
        thr = {1: 0.3, 2: 0.0, 3: 0.9, 4: 0.0}[method]
        kw = dict(method=method, score_threshold=thr, N_object=5)
        assert_hits_equal(mtm.matchTemplates(labelled, img, context=ct, **kw), mtm.matchTemplates(labelled, img, context=cd, **kw), tol=0.0)


def test_tensor_path_uniform_noise_and_bright(mtm, ctxs):
    """Saturated inputs: 255*255*h*w up to 4.26e9 needs the full unsigned 32-bit accumulator range."""
    from oracle import ncc_exact
    ct, _ = ctxs
    rng = np.random.default_rng(11)
    img = rng.integers(250, 256, (300, 330), dtype=np.uint8)
    tmpl = rng.integers(250, 256, (256, 256), dtype=np.uint8)
    got = mtm.computeScoreMap(tmpl, img, context=ct)
    exact = ncc_exact.match_template_exact(img, tmpl, use_fft=True)
    assert_map_close(got, exact)
    img2 = rng.integers(0, 256, (150, 190), dtype=np.uint8)
    t2 = rng.integers(0, 256, (40, 40), dtype=np.uint8)
    assert_map_close(mtm.computeScoreMap(t2, img2, context=ct), ncc_exact.match_template_exact(img2, t2, use_fft=False))


def test_tensor_path_flat_and_constant(mtm, ctxs):
    ct, _ = ctxs
    img = np.full((60, 80), 7, np.uint8)
    img[10:30, 10:30] = 200
    assert np.all(mtm.computeScoreMap(np.full((8, 8), 31, np.uint8), img, context=ct) == 1.0)
    t = np.arange(64, dtype=np.uint8).reshape(8, 8)
    m = mtm.computeScoreMap(t, img, context=ct)
    assert np.isfinite(m).all() and m[0, 60] == 0.0 and m[45, 0] == 0.0


@pytest.mark.parametrize("cfg", ["C2", "C4"])
def test_tensor_path_baseline_configs(mtm, ctxs, cfg):
    from oracle import mtm_port, synth
    ct, _ = ctxs
    image, temps, params = synth.config(cfg)
    got = mtm.matchTemplates(temps, image, context=ct, **params)
    want = mtm_port.match_templates(temps, image, **params)
    assert len(want) > 0
    assert_hits_equal(got, want)


def test_tensor_path_c3(mtm, ctxs):
    from oracle import ncc_exact, synth
    ct, _ = ctxs
    image, temps, params = synth.config("C3")
    got = mtm.computeScoreMap(temps[0][1], image, context=ct)
    exact = ncc_exact.match_template_exact(image, temps[0][1], use_fft=True)
    assert_map_close(got, exact)
