"""-m gpu parity tests of SURVEY §8 f3: on-device template augmentation (transform.cu), INTER_AREA downscale,
device-side search regions and the coarse-to-fine front end, against the numpy / live-cv2 restatements of the
tutorials' user code (oracle/augment_port.py) and the notebook answers of Tutorial3."""
import numpy as np
import pytest

from helpers import assert_hits_equal

pytestmark = pytest.mark.gpu

ALL = ("identity", "rot90", "rot180", "rot270", "fliplr", "flipud", "transpose", "antitranspose")


def _read_back_image(ctx, shape, dtype):
    """Pixels of the context's current image, read through TM_CCORR with a one-hot 1x1 template (exact)."""
    C = 1 if len(shape) == 2 else shape[2]
    planes = []
    for c in range(C):
        one = np.zeros((1, 1) if C == 1 else (1, 1, C), dtype)
        one[(0, 0) if C == 1 else (0, 0, c)] = 1
        ctx.set_templates([one])
        planes.append(ctx.score_map(0, 2, shape[:2]))
    return planes[0] if C == 1 else np.stack(planes, axis=2)


@pytest.mark.parametrize("dtype,channels", [(np.uint8, 1), (np.uint8, 3), (np.uint8, 4), (np.uint16, 1), (np.float32, 1), (np.float32, 3)])
def test_device_downscale_equals_inter_area(mtm, dtype, channels):
    """mtm_set_image_scaled: bit-identical to cv2.resize(INTER_AREA) for integer pixels (ties, cropped remainders)."""
    from mtm_b200 import _native
    from oracle import augment_port as ap
    rng = np.random.default_rng(21)
    ctx = _native.Context(0)
    for f in (1, 2, 3, 4, 5, 8, 16):
        shape = (37 * f + (f - 1), 53 * f + 1) + ((channels,) if channels > 1 else ())
        if dtype == np.float32:
            img = (rng.random(shape) * 255).astype(np.float32)
        else:
            img = rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
        want = ap.area_downscale(img, f)
        with ctx.lock:
            ctx.set_image_scaled(img, f)
            got = _read_back_image(ctx, want.shape, dtype)
        if dtype == np.float32:
            assert np.max(np.abs(got - want)) <= 1e-6 * 255, f
            assert np.max(np.abs(got - ap.cv_area_downscale(img, f))) <= 1e-5 * 255, f
        else:
            assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), "factor %d" % f
            assert np.array_equal(want, ap.cv_area_downscale(img, f))
    ctx.close()


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_device_transforms_equal_numpy(mtm, dtype):
    """mtm_set_templates_transformed: every symmetry (and the reduced templates) bit-identical to numpy / the
    INTER_AREA rule.  Read-back: TM_CCORR of a unit impulse returns the template mirrored."""
    from mtm_b200 import _native
    from oracle import augment_port as ap
    rng = np.random.default_rng(22)
    bases = [rng.integers(0, 256, (13, 22)).astype(dtype), rng.integers(0, 256, (30, 9)).astype(dtype)]
    ctx = _native.Context(0)
    for f in (1, 2, 3):
        want = [np.ascontiguousarray(ap.HOST_TRANSFORMS[t](ap.area_downscale(b, f))) for b in bases for t in ALL]
        H = 2 * max(w.shape[0] for w in want) - 1
        W = 2 * max(w.shape[1] for w in want) - 1
        with ctx.lock:
            ctx.set_templates_transformed(bases, [mtm.TRANSFORMS[t] for t in ALL], f)
            for k, w in enumerate(want):
                h_, w_ = w.shape
                impulse = np.zeros((2 * h_ - 1, 2 * w_ - 1), dtype)
                impulse[h_ - 1, w_ - 1] = 1
                canvas = np.zeros((H, W), dtype)
                canvas[:2 * h_ - 1, :2 * w_ - 1] = impulse
                ctx.set_image(canvas)
                m = ctx.score_map(k, 2, (H - h_ + 1, W - w_ + 1))[:h_, :w_]
                got = m[::-1, ::-1]
                assert np.array_equal(got, w.astype(np.float32)), (f, k)
    ctx.close()


@pytest.mark.parametrize("case", ["gray", "rgb", "float32", "flips_n3"])
def test_match_templates_augmented(mtm, case):
    from oracle import augment_port as ap, synth
    rng = np.random.default_rng(31)
    base = [synth.make_template(rng, 24, 40), synth.make_template(rng, 32, 32)]
    transforms = ("identity", "rot90", "rot180", "rot270")
    kw = dict(score_threshold=0.5, maxOverlap=0.25)
    planted = [np.ascontiguousarray(np.rot90(base[0], 1)), base[1], np.ascontiguousarray(np.rot90(base[1], 2)), base[0]]
    img, _ = synth.make_scene(300, 420, planted, 2, seed=31)
    if case == "rgb":
        img = np.stack([img, img[::-1], 255 - img], axis=2)
        base = [np.ascontiguousarray(img[40:64, 50:90]), np.ascontiguousarray(img[100:132, 200:232])]
    elif case == "float32":
        img = img.astype(np.float32) * 0.5
        base = [b.astype(np.float32) * 0.5 for b in base]
    elif case == "flips_n3":
        transforms = ("fliplr", "identity", "flipud", "transpose", "antitranspose")
        kw = dict(score_threshold=0.4, maxOverlap=0.1, N_object=3)
    labelled = [("a", base[0]), ("b", base[1])]
    got = mtm.matchTemplatesAugmented(labelled, img, transforms, **kw)
    same_kernels = mtm.matchTemplates(mtm.expandTemplates(labelled, transforms), img, **kw)
    assert [(h[0], h[1], float(h[2])) for h in got] == [(h[0], h[1], float(h[2])) for h in same_kernels]
    want = ap.match_templates_augmented(labelled, img, transforms, **kw)
    assert len(want) >= 3
    assert_hits_equal(got, want)
    # searchBox + a second call with the same template set (content-hash hit) + a different set afterwards
    box = (20, 10, 380, 260)
    got = mtm.matchTemplatesAugmented(labelled, img, transforms, searchBox=box, **kw)
    assert_hits_equal(got, ap.match_templates_augmented(labelled, img, transforms, searchBox=box, **kw))
    got = mtm.matchTemplatesAugmented(labelled[:1], img, ("rot180",), **kw)
    assert_hits_equal(got, ap.match_templates_augmented(labelled[:1], img, ("rot180",), **kw))
    assert_hits_equal(mtm.matchTemplates(labelled, img, **kw), ap.match_templates_augmented(labelled, img, ("identity",), **kw))


def test_device_search_region_equals_host_crop(mtm):
    """mtm_set_image_roi == the searchBox crop of MTM/__init__.py:140-144 done on the host."""
    from mtm_b200 import _native
    from oracle import synth
    rng = np.random.default_rng(33)
    temps = [synth.make_template(rng, 20, 28), synth.make_template(rng, 31, 17)]
    img, _ = synth.make_scene(260, 333, temps, 3, seed=33)
    ctx = _native.Context(0)
    with ctx.lock:
        ctx.set_image_scaled(img, 1)
        ctx.set_templates(temps)
        whole = ctx.find_matches(5, -1, 0.4).copy()
        for (x, y, w, h) in [(0, 0, 333, 260), (37, 21, 201, 150), (101, 3, 232, 257), (5, 200, 64, 60)]:
            ctx.set_image_roi(x, y, w, h)
            got = ctx.find_matches(5, -1, 0.4).copy()
            ctx.set_image(img[y:y + h, x:x + w])
            want = ctx.find_matches(5, -1, 0.4).copy()
            assert got.tolist() == want.tolist(), (x, y, w, h)
        assert len(whole) >= 4
        with pytest.raises(_native.NativeError, match="outside"):
            ctx.set_image_roi(300, 0, 64, 64)
    fresh = _native.Context(0)
    with pytest.raises(_native.NativeError, match="no full-resolution image"):
        fresh.set_image_roi(0, 0, 8, 8)
    fresh.close()
    ctx.close()


def test_pyramid_reproduces_notebook_answers(mtm):
    """Tutorial3: the refined coarse-to-fine search gives cell 10's answer; the device-reduced Fish image searched
    with the notebook's small template gives cell 21's."""
    from mtm_b200 import _native
    from oracle import golden_cases as gc
    fish = gc.fish()
    head = [("head", fish[842:842 + 184, 528:528 + 196])]
    for f in (2, 4, 8):
        got = mtm.matchTemplatesPyramid(head, fish, downscale=f, N_object=1)
        assert_hits_equal(got, gc.NOTEBOOK_ANSWERS["t3_full"], tol=1e-5)
    kind, temps, small, kw = gc.build("t3_downscaled")
    ctx = _native.Context(0)
    with ctx.lock:
        ctx.set_image_scaled(fish, 4)                       # cv2.resize(image, (512, 512), INTER_AREA) on the device
        ctx.set_templates([np.ascontiguousarray(temps[0][1])])
        raw = ctx.match_templates(5, 1, 0.5, 0.25)
    want = gc.NOTEBOOK_ANSWERS["t3_downscaled"][0]
    assert (int(raw[0]["x"]), int(raw[0]["y"]), int(raw[0]["w"]), int(raw[0]["h"])) == want[1]
    assert abs(float(raw[0]["score"]) - want[2]) <= 1e-5
    ctx.close()


@pytest.mark.parametrize("f,refine,kw", [
    (2, True, dict(score_threshold=0.5, maxOverlap=0.25)),
    (4, True, dict(score_threshold=0.5, maxOverlap=0.25)),
    (4, False, dict(score_threshold=0.4, maxOverlap=0.25)),
    (3, True, dict(score_threshold=0.5, maxOverlap=0.1, N_object=5)),
    (4, True, dict(score_threshold=0.5, maxOverlap=0.25, searchBox=(40, 30, 700, 520), coarse_threshold=0.35)),
    (2, True, dict(score_threshold=0.3, maxOverlap=0.25, method=1, N_object=1)),
])
def test_pyramid_equals_specification(mtm, f, refine, kw):
    from oracle import augment_port as ap, mtm_port, synth
    rng = np.random.default_rng(3)
    temps = [synth.make_template(rng, 64, 64), synth.make_template(rng, 48, 80)]
    img, _ = synth.make_scene(600, 800, temps, 4, seed=3)
    labelled = [("a", temps[0]), ("b", temps[1])]
    got = mtm.matchTemplatesPyramid(labelled, img, downscale=f, refine=refine, **kw)
    want = ap.match_templates_pyramid(labelled, img, downscale=f, refine=refine, **kw)
    assert len(want) >= 1
    assert_hits_equal(got, want)
    if refine and kw.get("method", 5) == 5 and "N_object" not in kw:
        full = mtm_port.match_templates(labelled, img, **{k: v for k, v in kw.items() if k != "coarse_threshold"})
        assert_hits_equal(got, full)                        # and the full-resolution search finds the same objects


def _f3_golden():
    import json
    import os
    from oracle import golden_cases as gc
    with open(os.path.join(gc.GOLDEN_DIR, "ref_outputs_f3.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["aug_rot4", "aug_flips_n3", "pyr_f4_refined", "pyr_f4_coarse", "pyr_f3_n5",
                                  "pyr_f2_sqdiff_n1", "pyr_fish_f4"])
def test_f3_front_ends_against_the_unmodified_reference(mtm, name):
    """Golden vectors made by running the tutorials' user code on the UNMODIFIED reference (oracle/make_golden.py --f3)."""
    from oracle import golden_cases as gc
    case = gc.build_f3(name)
    if case[0] == "aug":
        _, temps, transforms, img, kw = case
        got = mtm.matchTemplatesAugmented(temps, img, transforms, **kw)
    else:
        _, temps, img, f, refine, kw = case
        got = mtm.matchTemplatesPyramid(temps, img, downscale=f, refine=refine, **kw)
    want = [(w[0], tuple(w[1]), w[2]) for w in _f3_golden()[name]]
    assert_hits_equal(got, want)
