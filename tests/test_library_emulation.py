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


def test_many_template_sizes_are_grouped_into_full_launches(lib_mtm):
    """A C5-like template set (20 distinct square sizes) through the real host code: plan_tensor_path partitions the (h, w)-sorted
    list by dynamic programming into mode-A launches of up to eight templates zero padded to the largest member (the first planner
    left launches with four or five templates at the price of eight); the hit list equals the port's."""
    from mtm_b200 import _native
    from oracle import mtm_port, synth
    rng = np.random.default_rng(21)
    sides = np.linspace(16, 54, 20).round().astype(int)
    temps = [synth.make_template(rng, int(s), int(s)) for s in sides]
    image, _ = synth.make_scene(150, 190, temps[::3], 1, seed=21)
    labelled = [("t%02d" % i, t) for i, t in enumerate(temps)]
    ctx = _native.default_context()
    ctx.reset_counters()
    got = lib_mtm.matchTemplates(labelled, image, N_object=30, score_threshold=0.5, maxOverlap=0.25)
    want = mtm_port.match_templates(labelled, image, N_object=30, score_threshold=0.5, maxOverlap=0.25)
    assert len(want) >= 5
    gp.assert_hits_equal(got, want)
    # one tensor map per numerator launch: 20 templates on a small image -> three launches (the per-launch term of the cost model
    # outweighs what single-template mode-B launches of the smallest templates would save)
    assert ctx.counters()["tma_launches"] == 3, ctx.counters()


def test_banded_moment_ring_through_the_host_code(lib_mtm, monkeypatch):
    """The window moments of a template group are produced band by band into a ring (mtm_api.cu: compute_maps), each band
    followed by its numerator launch.  A 64 KB ring forces several bands on a small scene: hit lists (hits-only search), the
    N_object = 1 search and the score maps must equal the single-band results bit for bit, for one size, mixed sizes and RGB."""
    from mtm_b200 import _native
    from oracle import synth
    rng = np.random.default_rng(12)
    # (distinct sizes: a size shared by two launch groups keeps the resident placement)
    temps = [("a", synth.make_template(rng, 20, 24)), ("b", synth.make_template(rng, 22, 26)), ("c", synth.make_template(rng, 14, 31))]
    img, _ = synth.make_scene(150, 200, [t[1] for t in temps], 3, seed=12)
    rgb = np.stack([img, np.roll(img, 3, axis=0), 255 - img], axis=2)
    t_rgb = [("r", np.ascontiguousarray(rgb[30:52, 40:70])), ("s", np.ascontiguousarray(rgb[90:110, 100:124]))]
    results = []
    for ring_kb in (None, "64"):
        if ring_kb:
            monkeypatch.setenv("MTM_B200_RING_KB", ring_kb)
        ctx = _native.Context(0)
        try:
            hits = lib_mtm.matchTemplates(temps, img, score_threshold=0.5, maxOverlap=0.3, context=ctx)
            launches = ctx.counters()["kernel_launches"]
            one = lib_mtm.matchTemplates(temps, img, N_object=1, context=ctx)
            maps = [lib_mtm.computeScoreMap(t, img, context=ctx) for _, t in temps]
            hits_rgb = lib_mtm.matchTemplates(t_rgb, rgb, score_threshold=0.6, maxOverlap=0.3, context=ctx)
            results.append((hits, one, maps, hits_rgb, launches))
        finally:
            ctx.close()
    a, b = results
    assert b[4] > a[4] + 4, "the small ring did not split the search into bands"
    assert a[0] == b[0] and len(a[0]) >= 6 and a[1] == b[1] and a[3] == b[3] and len(a[3]) >= 2
    for ma, mb in zip(a[2], b[2]):
        assert np.array_equal(ma.view(np.uint32), mb.view(np.uint32))


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


@pytest.mark.parametrize("seed", [203, 231])
def test_call_sequences_keep_no_stale_state(lib_mtm, monkeypatch, seed):
    """One context, a random sequence of calls that change image size, dtype, channel count, template set, method, threshold, N_object
    and search box between them: everything the library keeps between calls (map geometry of equal-sized images, the hashed template
    set, window moments, the candidate list, the result mirror) must never leak into the next answer.  Expected results: the port
    on EXACT maps (live cv2's own fp32 noise moves near-threshold peaks of methods 1 / 3)."""
    import warnings
    from oracle import mtm_port, ncc_exact, synth

    def exact_map(template, image, method=5, mask=None):
        if not (template.dtype == np.uint8 and image.dtype == np.uint8):
            template, image = np.float32(template), np.float32(image)
        return ncc_exact.match_template_exact(image, template, method)

    monkeypatch.setattr(mtm_port, "compute_score_map", exact_map)
    rng = np.random.default_rng(seed)
    sets, scenes = {}, {}
    for step in range(8):
        C = int(rng.choice([1, 1, 3]))
        kind = int(rng.integers(0, 2))
        if (kind, C) not in sets:
            k = int(rng.integers(1, 5))
            shapes = [(int(rng.integers(8, 25)), int(rng.integers(8, 30)))] * k if kind == 0 else \
                [(int(rng.integers(6, 25)), int(rng.integers(6, 30))) for _ in range(k)]
            ts = [synth.make_template(rng, h, w) for h, w in shapes]
            if C == 3:
                ts = [np.ascontiguousarray(np.stack([t, 255 - t, t[::-1, ::-1]], axis=2)) for t in ts]
            sets[kind, C] = [("t%d" % i, t) for i, t in enumerate(ts)]
        temps = sets[kind, C]
        idx = int(rng.integers(0, 3))
        if (idx, kind, C) not in scenes:
            H, W = [(90, 130), (90, 130), (70, 101)][idx]
            img, _ = synth.make_scene(H, W, [t[1] if C == 1 else t[1][:, :, 0] for t in temps], 2, seed=seed * 10 + idx)
            scenes[idx, kind, C] = img if C == 1 else np.ascontiguousarray(np.stack([img, 255 - img, img[::-1, ::-1]], axis=2))
        img = scenes[idx, kind, C]
        dt = rng.choice(["u8", "u8", "u8", "f32", "u16"]) if C == 1 else rng.choice(["u8", "u8", "f32"])
        if dt == "f32":
            im, ts = img.astype(np.float32), [(n, t.astype(np.float32)) for n, t in temps]
        elif dt == "u16":
            im, ts = img.astype(np.uint16) * 150, [(n, t.astype(np.uint16) * 150) for n, t in temps]
        else:
            im, ts = img, temps
        method = int(rng.choice([5, 5, 5, 1, 3]))
        thr = 0.9 if method == 3 else float(rng.choice([0.3, 0.4])) if method == 1 else float(rng.choice([0.4, 0.5, 0.7]))
        n_object = [float("inf"), 1, int(rng.integers(2, 6))][int(rng.integers(0, 3))]
        box = (int(rng.integers(0, 10)), int(rng.integers(0, 10)), im.shape[1] - 12, im.shape[0] - 12) if rng.random() < 0.3 else None
        kw = dict(method=method, N_object=n_object, score_threshold=thr, maxOverlap=float(rng.choice([0.0, 0.25, 0.5])), searchBox=box)
        op = int(rng.integers(0, 3))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if op == 0:
                got, want = lib_mtm.matchTemplates(ts, im, **kw), mtm_port.match_templates(ts, im, **kw)
                assert [(h[0], h[1]) for h in got] == [(h[0], h[1]) for h in want], (step, kw)
                assert all(abs(float(a[2]) - float(b[2])) <= 1e-4 for a, b in zip(got, want))
            elif op == 1:
                kw.pop("maxOverlap")
                got, want = lib_mtm.findMatches(ts, im, **kw), mtm_port.find_matches(ts, im, **kw)
                assert sorted((h[0], h[1]) for h in got) == sorted((h[0], h[1]) for h in want), (step, kw)
            else:
                t = ts[int(rng.integers(0, len(ts)))][1]
                got, want = lib_mtm.computeScoreMap(t, im, method=method), exact_map(t, im, method)
                assert got.shape == want.shape and float(np.max(np.abs(got - want))) <= 1e-4 * max(1.0, float(np.abs(want).max())), (step, method)


def test_baseline_config_c2_at_full_size(lib_mtm):
    """BASELINE configs[1] -- 1920 x 1080, 8 templates 64 x 64 (two bases x four rotations), threshold 0.5 -- through the host build at
    FULL size: 585 tiles and 112 320 tcgen05.mma on the functional model, candidate list, one-launch sort + NMS; the hit list equals the
    port's (the body of tests/test_gpu_parity.py::test_baseline_configs_full_size[C2])."""
    gp.test_baseline_configs_full_size(lib_mtm, "C2")
