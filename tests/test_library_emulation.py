"""CPU tests of the WHOLE library: csrc/*.cu -- C ABI, host-side sequencing (mtm_api.cu) and every kernel -- compiled for the host by
tests/emu_library.py and loaded through the product's own ctypes binding (mtm_b200._native), so that the BODIES of the -m gpu parity
tests run here against the golden vectors of the unmodified reference and against the oracle.  The tensor-core route runs on the
functional tcgen05 model of tests/emu/tcgen05_model.h; the emulated device has 4 SMs.

The same build serves `MTM_B200_EMULATE=1 python -m pytest tests -m gpu -k ...` (tests/conftest.py): any -m gpu test on the CPU, slowly.
This file is the curated, fast selection that belongs to the default CPU suite.

TEST INFRASTRUCTURE: the host build lives in a temporary directory and is reachable only through the monkeypatched library path
of these tests; the product still fails loudly without a GPU (tests/test_abi_and_host.py)."""
import shutil

import numpy as np
import pytest

import emu_library
import test_gpu_parity as gp


@pytest.fixture(scope="session")
def emu_lib_path(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    return emu_library.build(str(tmp_path_factory.mktemp("emu_lib")))


@pytest.fixture()
def lib_mtm(emu_lib_path, mtm, monkeypatch):
    """The product API with libmtm_b200.so replaced by its host build (fresh contexts; everything is restored afterwards)."""
    from mtm_b200 import _native
    monkeypatch.setattr(_native, "LIB_PATH", emu_lib_path)
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "_default", {})
    monkeypatch.setattr(_native, "_helpers", {})
    yield mtm
    for ctx in list(_native._default.values()) + [c for cs in _native._helpers.values() for c in cs]:
        ctx.close()


@pytest.mark.parametrize("name", ["t3_downscaled", "c1_fish256_n1", "c1_fish256_inf", "synth_rot8", "synth_mixed_n5", "synth_searchbox",
                                  "synth_exact_fit"])
def test_match_templates_golden(lib_mtm, golden, name):
    """MTM.matchTemplates -> api.py -> ctypes -> mtm_match_templates (host build: tensor-core route on the model, candidate list or
    arg-best, one-launch sort + NMS, mapped result mirror) == outputs of the UNMODIFIED reference."""
    gp.test_match_templates_golden(lib_mtm, golden, name)
    from mtm_b200 import _native
    assert _native.default_context().counters()["kernel_launches"] > 0


@pytest.mark.parametrize("name", ["c1_fish256_find", "synth_row_map", "synth_col_map"])
def test_find_matches_golden(lib_mtm, golden, name):
    gp.test_find_matches_golden(lib_mtm, golden, name)


def test_score_maps(lib_mtm, golden):
    """computeScoreMap: the reference's own Fish map, all six methods, RGB, flat windows and constant templates."""
    gp.test_score_map_c1_fish(lib_mtm, golden)
    for method in range(6):
        gp.test_score_map_all_methods(lib_mtm, method)
    gp.test_score_map_rgb(lib_mtm)
    gp.test_flat_and_constant_inputs(lib_mtm)


def test_nms_validation_and_other_methods(lib_mtm, golden):
    """mtm_nms (standalone NMS entry point), the reference's error behaviour, and matchTemplates with TM_SQDIFF_NORMED /
    TM_CCORR_NORMED against the port."""
    gp.test_nms_demo_and_random(lib_mtm, golden)
    gp.test_validation_errors_match_reference(lib_mtm)
    for method, thr in ((1, 0.35), (3, 0.92)):
        gp.test_match_templates_other_methods_vs_port(lib_mtm, method, thr)


def test_direct_and_tensor_routes_agree(lib_mtm):
    """MTM_OPT_PATH: the dp4a route and the tcgen05 route through the same C ABI give the same hit lists and maps within 2e-6."""
    from mtm_b200 import _native
    from oracle import synth
    rng = np.random.default_rng(8)
    temps = [("a", synth.make_template(rng, 20, 24)), ("b", synth.make_template(rng, 20, 24)), ("c", synth.make_template(rng, 14, 31))]
    img, _ = synth.make_scene(120, 160, [t[1] for t in temps], 3, seed=8)
    ct, cd = _native.Context(0), _native.Context(0)
    try:
        ct.set_path(_native.PATH_TENSOR)
        cd.set_path(_native.PATH_DIRECT)
        for name, t in temps:
            a, b = lib_mtm.computeScoreMap(t, img, context=ct), lib_mtm.computeScoreMap(t, img, context=cd)
            assert np.max(np.abs(a - b)) <= 2e-6
        ht = lib_mtm.matchTemplates(temps, img, score_threshold=0.5, context=ct)
        hd = lib_mtm.matchTemplates(temps, img, score_threshold=0.5, context=cd)
        assert [(h[0], h[1]) for h in ht] == [(h[0], h[1]) for h in hd] and len(ht) >= 6
        assert ct.counters()["ncc_launches"] >= 0
    finally:
        ct.close()
        cd.close()


def test_device_resident_images_through_the_real_host_code(lib_mtm):
    """The body of tests/test_gpu_zz_device_inputs.py with host pointers dressed as CUDA arrays (the emulated device's memory IS host
    memory): mtm_set_image_device for uint8 / float32 / uint16 / RGB, searchBox crops as pointer arithmetic, no upload of the image,
    same results as the numpy route; the batch entry point with mixed device / host images and a strided view."""
    from mtm_b200 import _native
    from oracle import synth
    from test_host_api_mock import _FakeDeviceArray
    rng = np.random.default_rng(43)
    temps = [("a", synth.make_template(rng, 16, 24)), ("b", synth.make_template(rng, 20, 20))]
    img, _ = synth.make_scene(110, 150, [t[1] for t in temps], 3, seed=43)
    cases = [(img, temps),
             (img.astype(np.float32), [(n, t.astype(np.float32)) for n, t in temps]),
             (img.astype(np.uint16) * 100, [(n, t.astype(np.uint16) * 100) for n, t in temps]),
             (np.stack([img, 255 - img, img[::-1]], axis=2), [(n, np.ascontiguousarray(np.stack([t, 255 - t, t[::-1]], axis=2))) for n, t in temps])]
    ctx = _native.default_context()
    for image, ts in cases:
        image = np.ascontiguousarray(image)
        dev = _FakeDeviceArray(image)
        all_kw = (dict(score_threshold=0.5, maxOverlap=0.25), dict(N_object=1), dict(score_threshold=0.5, searchBox=(13, 7, 120, 90)))
        for kw in (all_kw if image is cases[0][0] or image.dtype == np.uint8 and image.ndim == 2 else all_kw[2:]):
            before = ctx.counters()["h2d_bytes"]
            got = lib_mtm.matchTemplates(ts, dev, **kw)
            moved = ctx.counters()["h2d_bytes"] - before
            assert moved < image.nbytes // 4, "the image must not be uploaded again (%d bytes moved)" % moved
            want = lib_mtm.matchTemplates(ts, image, **kw)
            assert [(h[0], h[1], float(h[2])) for h in got] == [(h[0], h[1], float(h[2])) for h in want] and len(want) >= 1
        assert np.array_equal(lib_mtm.computeScoreMap(ts[0][1], dev), lib_mtm.computeScoreMap(ts[0][1], image))
    view = np.ascontiguousarray(img)[10:100, 20:140]                         # a strided view, as torch slicing gives
    batch = lib_mtm.matchTemplatesBatch(temps, [_FakeDeviceArray(img), img, _FakeDeviceArray(view)], score_threshold=0.5)
    want = [lib_mtm.matchTemplates(temps, im, score_threshold=0.5) for im in (img, img, np.ascontiguousarray(view))]
    assert [[(h[0], h[1], float(h[2])) for h in hits] for hits in batch] == [[(h[0], h[1], float(h[2])) for h in hits] for hits in want]
