"""CPU tests of the host layer above the C ABI for SURVEY §8 f3 (augment.py) with the oracle-backed test double of
tests/mock_device.py in place of the GPU context.  They run the BODIES of the -m gpu tests of test_gpu_xform.py, so
what those tests expect from the device is checked against the restated semantics before it reaches the GPU box."""
import numpy as np
import pytest

import test_gpu_xform as gx
from mock_device import MockContext


@pytest.fixture()
def mock_mtm(mtm, monkeypatch):
    from mtm_b200 import _native
    shared = MockContext()
    monkeypatch.setattr(_native, "Context", MockContext)
    monkeypatch.setattr(_native, "default_context", lambda device=None: shared)
    helpers = [MockContext() for _ in range(4)]
    monkeypatch.setattr(_native, "helper_contexts", lambda device, n, owner=None: helpers[:n])
    mtm._mock_helpers = helpers
    mtm._mock = shared
    return mtm


@pytest.mark.parametrize("dtype,channels", [(np.uint8, 1), (np.uint8, 3), (np.uint16, 1), (np.float32, 3)])
def test_downscale_read_back(mock_mtm, dtype, channels):
    gx.test_device_downscale_equals_inter_area(mock_mtm, dtype, channels)


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_transform_read_back(mock_mtm, dtype):
    gx.test_device_transforms_equal_numpy(mock_mtm, dtype)


@pytest.mark.parametrize("case", ["gray", "rgb", "float32", "flips_n3"])
def test_augmented_front_end(mock_mtm, case):
    gx.test_match_templates_augmented(mock_mtm, case)
    assert "set_templates_transformed" in mock_mtm._mock.calls


def test_search_region(mock_mtm):
    gx.test_device_search_region_equals_host_crop(mock_mtm)


def test_pyramid_notebook_answers(mock_mtm):
    gx.test_pyramid_reproduces_notebook_answers(mock_mtm)
    calls = mock_mtm._mock.calls
    assert "set_image_scaled" in calls and "set_image_roi" in calls and "set_image" not in calls   # one upload per search


@pytest.mark.parametrize("f,refine,kw", [
    (2, True, dict(score_threshold=0.5, maxOverlap=0.25)),
    (4, False, dict(score_threshold=0.4, maxOverlap=0.25)),
    (3, True, dict(score_threshold=0.5, maxOverlap=0.1, N_object=5)),
    (4, True, dict(score_threshold=0.5, maxOverlap=0.25, searchBox=(40, 30, 700, 520), coarse_threshold=0.35)),
    (2, True, dict(score_threshold=0.3, maxOverlap=0.25, method=1, N_object=1)),
])
def test_pyramid_front_end(mock_mtm, f, refine, kw):
    gx.test_pyramid_equals_specification(mock_mtm, f, refine, kw)


def test_batch_front_end_equals_loop(mock_mtm):
    """matchTemplatesBatch over 1..3 streams == the per-image calls (offsets, labels, order; mixed image dtypes)."""
    from oracle import synth
    rng = np.random.default_rng(77)
    temps = [("t%d" % i, synth.make_template(rng, 24 + 4 * i, 30)) for i in range(3)]
    images = [synth.make_scene(160, 200, [t[1] for t in temps], 2, seed=100 + k)[0] for k in range(7)]
    images[3] = images[3].astype(np.float32)                   # dtype policy differs for this image: templates are re-routed
    for streams in (1, 2, 3):
        for kw in (dict(score_threshold=0.5, maxOverlap=0.25), dict(N_object=1), dict(score_threshold=0.5, searchBox=(10, 20, 150, 120))):
            want = [mock_mtm.matchTemplates(temps, im, **kw) for im in images]
            got = mock_mtm.matchTemplatesBatch(temps, images, streams=streams, **kw)
            assert len(got) == len(want) == 7
            for g, w in zip(got, want):
                assert [(a[0], a[1], float(a[2])) for a in g] == [(b[0], b[1], float(b[2])) for b in w]
    assert "match_templates_async" in mock_mtm._mock_helpers[0].calls and not mock_mtm._mock_helpers[3].calls
    with pytest.raises(ValueError, match="larger than image"):     # a bad image in the middle: slots are drained, error propagates
        mock_mtm.matchTemplatesBatch(temps, images[:2] + [images[0][:10, :10]] + images[2:], streams=2)
    assert not mock_mtm._mock.slots and not mock_mtm._mock_helpers[0].slots


@pytest.mark.parametrize("name", ["aug_rot4", "aug_flips_n3", "pyr_f4_refined", "pyr_f4_coarse", "pyr_f3_n5",
                                  "pyr_f2_sqdiff_n1", "pyr_fish_f4"])
def test_f3_front_ends_against_reference_goldens(mock_mtm, name):
    gx.test_f3_front_ends_against_the_unmodified_reference(mock_mtm, name)
