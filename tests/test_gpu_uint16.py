"""-m gpu: 16-bit grayscale inputs (MTM_U16).  The reference casts uint16 to float32 (MTM/__init__.py:71-74) and OpenCV
correlates in floating point; here the numerator comes EXACTLY from four u8 x u8 tensor-core correlations of the
high/low byte planes (65536 hh + 256 (hl + lh) + ll), the window statistics and the epilogue stay float64/float32 as
on the float32 path.  Oracle: the exact float64 restatement, live cv2 and the fp32 CUDA kernel."""
import numpy as np
import pytest

from helpers import assert_hits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctxs():
    import MTM  # noqa: F401
    from mtm_b200 import _native
    t = _native.Context(0)
    t.set_path(_native.PATH_TENSOR)          # a call that cannot take the tensor cores raises instead of falling back
    d = _native.Context(0)
    d.set_path(_native.PATH_DIRECT)          # fp32 FFMA kernel
    yield t, d
    t.close()
    d.close()


def _scene16(seed, H=180, W=250, sizes=((32, 32), (24, 40), (32, 32)), full_range=False):
    from oracle import synth
    rng = np.random.default_rng(seed)
    temps8 = [synth.make_template(rng, h, w) for (h, w) in sizes]
    img8, _ = synth.make_scene(H, W, temps8, 3, seed=seed)
    if full_range:                                         # all 16 bits in use, both byte planes busy
        noise = rng.integers(0, 256, img8.shape).astype(np.uint16)
        img = (img8.astype(np.uint16) << 8) | noise
        temps = [((t.astype(np.uint16) << 8) | rng.integers(0, 256, t.shape).astype(np.uint16)) for t in temps8]
    else:                                                  # 12-bit data over a pedestal (microscopy-like)
        img = (img8.astype(np.uint16) * 13 + 3000).astype(np.uint16)
        temps = [(t.astype(np.uint16) * 13 + 3000).astype(np.uint16) for t in temps8]
    return img, temps


@pytest.mark.parametrize("full_range", [False, True])
@pytest.mark.parametrize("method", [5, 3, 1, 4, 2, 0])
def test_uint16_maps_tensor_vs_exact_cv2_and_fp32_kernel(mtm, ctxs, method, full_range):
    import cv2
    from oracle import ncc_exact
    ct, cd = ctxs
    img, temps = _scene16(7 + method, full_range=full_range)
    imgf = img.astype(np.float32)
    for t in temps[:2]:
        tf = t.astype(np.float32)
        got = mtm.computeScoreMap(t, img, method=method, context=ct)
        assert got.dtype == np.float32
        exact = ncc_exact.match_template_exact(imgf, tf, method=method, use_fft=False)
        cv = cv2.matchTemplate(imgf, tf, method)
        scale = max(1.0, float(np.abs(exact).max()))
        assert np.max(np.abs(got.astype(np.float64) - exact)) <= 1e-4 * scale
        assert np.max(np.abs(got.astype(np.float64) - cv)) <= 1e-4 * scale + np.max(np.abs(cv - exact))
        fp32 = mtm.computeScoreMap(t, img, method=method, context=cd)
        assert np.max(np.abs(got.astype(np.float64) - fp32)) <= 2e-4 * scale


def test_uint16_large_template_and_tile_edges(mtm, ctxs):
    """100 x 130 window at full 16-bit range: partial sums up to 255^2 * 13000 per plane, total ~5.6e13."""
    from oracle import ncc_exact
    ct, _ = ctxs
    rng = np.random.default_rng(2)
    img = rng.integers(0, 65536, (333, 417)).astype(np.uint16)
    t = np.ascontiguousarray(img[100:200, 150:280]).copy()
    t[::7, ::5] ^= 0x1234
    got = mtm.computeScoreMap(t, img, context=ct)
    exact = ncc_exact.match_template_exact(img.astype(np.float32), t.astype(np.float32), use_fft=False)
    assert np.max(np.abs(got.astype(np.float64) - exact)) <= 1e-4
    assert np.unravel_index(int(got.argmax()), got.shape) == (100, 150)


def test_uint16_match_templates_vs_port_and_fp32_kernel(mtm, ctxs):
    from oracle import mtm_port
    ct, cd = ctxs
    img, temps = _scene16(11, H=300, W=420, sizes=((32, 32), (24, 40), (32, 32), (48, 20)))
    labelled = [("t%d" % i, t) for i, t in enumerate(temps)]
    for kw in (dict(score_threshold=0.5, maxOverlap=0.25), dict(N_object=1), dict(method=1, score_threshold=0.4, N_object=6)):
        want = mtm_port.match_templates(labelled, img, **kw)
        assert len(want) > 0
        assert_hits_equal(mtm.matchTemplates(labelled, img, context=ct, **kw), want)
        assert_hits_equal(mtm.matchTemplates(labelled, img, context=cd, **kw), want)
    # searchBox: strided 16-bit view
    sb = (40, 30, 300, 200)
    assert_hits_equal(mtm.matchTemplates(labelled, img, searchBox=sb, context=ct), mtm_port.match_templates(labelled, img, searchBox=sb))
    # the pipelined entry point takes the same route
    batch = mtm.matchTemplatesBatch(labelled, [img, img[::-1].copy()], context=ct)
    assert_hits_equal(batch[0], mtm.matchTemplates(labelled, img, context=ct), tol=0.0)
    assert_hits_equal(batch[1], mtm.matchTemplates(labelled, img[::-1].copy(), context=ct), tol=0.0)


def test_uint16_mixed_inputs_keep_the_float32_route(mtm, ctxs):
    """uint16 image with a float32 template that is NOT integer-valued is plain float32 data: default context works, the
    tensor-only context refuses.  (An integer-valued float32 template takes the exact byte-plane route: next test.)"""
    from mtm_b200 import _native
    from oracle import ncc_exact
    ct, _ = ctxs
    img, temps = _scene16(13)
    tf = temps[0].astype(np.float32) + np.float32(0.25)
    got = mtm.computeScoreMap(tf, img)
    exact = ncc_exact.match_template_exact(img.astype(np.float32), tf, use_fft=False)
    assert np.max(np.abs(got.astype(np.float64) - exact)) <= 1e-4
    with pytest.raises(_native.NativeError):
        mtm.computeScoreMap(tf, img, context=ct)
    # and back to uint8 on the same contexts afterwards
    img8 = (img >> 8).astype(np.uint8) + np.uint8(3)
    t8 = np.ascontiguousarray(img8[20:52, 30:62])
    m = mtm.computeScoreMap(t8, img8, context=ct)
    assert np.unravel_index(int(m.argmax()), m.shape) == (20, 30) and abs(float(m.max()) - 1.0) < 1e-6


@pytest.mark.parametrize("method", [5, 1])
def test_float32_integer_valued_data_takes_the_exact_tensor_route(mtm, ctxs, method):
    """float32 arrays holding integers in [0, 65535] (uint16 data cast by the caller, or a uint16 image with a float32 template after
    the reference's float32 cast, MTM/__init__.py:71-74) are detected on upload and matched through the byte planes like MTM_U16:
    the tensor-only context accepts them and the map equals the MTM_U16 one bit for bit."""
    from mtm_b200 import _native
    from oracle import ncc_exact
    ct, cd = ctxs
    img, temps = _scene16(17 + method, full_range=True)
    imgf, tf = img.astype(np.float32), temps[0].astype(np.float32)
    want16 = mtm.computeScoreMap(temps[0], img, method=method, context=ct)
    for (t_in, i_in) in ((tf, imgf), (tf, img), (temps[0], imgf)):
        got = mtm.computeScoreMap(t_in, i_in, method=method, context=ct)
        assert np.array_equal(got, want16)
    exact = ncc_exact.match_template_exact(imgf, tf, method=method, use_fft=False)
    assert np.max(np.abs(want16.astype(np.float64) - exact)) <= 1e-4 * max(1.0, float(np.abs(exact).max()))
    # one fractional / negative / too large pixel anywhere in the image: plain float32 data again
    for bad in (0.5, -1.0, 65536.0, np.nan):
        img_bad = imgf.copy()
        img_bad[-1, -1] = bad
        with pytest.raises(_native.NativeError):
            mtm.computeScoreMap(tf, img_bad, method=method, context=ct)
    img_bad = imgf.copy()
    img_bad[3, 5] += 0.5
    got = mtm.computeScoreMap(tf, img_bad, method=method)
    exact = ncc_exact.match_template_exact(img_bad, tf, method=method, use_fft=False)
    assert np.max(np.abs(got.astype(np.float64) - exact)) <= 1e-4 * max(1.0, float(np.abs(exact).max()))
    # ... and the integer image right after it on the same context takes the planes again
    assert np.array_equal(mtm.computeScoreMap(tf, imgf, method=method, context=ct), want16)


def test_float32_integer_valued_match_templates(mtm, ctxs):
    from oracle import mtm_port
    ct, _ = ctxs
    img, temps = _scene16(19, H=300, W=420, sizes=((32, 32), (24, 40), (32, 32), (48, 20)))
    labelled16 = [("t%d" % i, t) for i, t in enumerate(temps)]
    labelledf = [(n, t.astype(np.float32)) for n, t in labelled16]
    imgf = img.astype(np.float32)
    for kw in (dict(score_threshold=0.5, maxOverlap=0.25), dict(N_object=1)):
        want = mtm_port.match_templates(labelledf, imgf, **kw)
        assert len(want) > 0
        got = mtm.matchTemplates(labelledf, imgf, context=ct, **kw)
        assert_hits_equal(got, want)
        assert_hits_equal(got, mtm.matchTemplates(labelled16, img, context=ct, **kw), tol=0.0)
    batch = mtm.matchTemplatesBatch(labelledf, [imgf, imgf[::-1].copy()], context=ct)
    assert_hits_equal(batch[0], mtm.matchTemplates(labelledf, imgf, context=ct), tol=0.0)
    assert_hits_equal(batch[1], mtm.matchTemplates(labelledf, imgf[::-1].copy(), context=ct), tol=0.0)
