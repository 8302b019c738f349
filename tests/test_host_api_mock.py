"""CPU tests of the host layer (api.py: validation, dtype / mask routing, offsets, label mapping, the rare-corner
splits, NMS plumbing) with the oracle-backed test double of tests/mock_device.py in place of the GPU context.  They run
the BODIES of the -m gpu parity tests that exercise host logic, against the golden vectors of the unmodified reference;
the kernels themselves are only tested on the GPU."""
import numpy as np
import pytest

import test_gpu_parity as gp
from mock_device import MockComm, MockContext


@pytest.fixture()
def mock_mtm(mtm, monkeypatch):
    from mtm_b200 import _native
    shared = MockContext()
    helpers = [MockContext() for _ in range(4)]
    monkeypatch.setattr(_native, "Context", MockContext)
    monkeypatch.setattr(_native, "Comm", MockComm)
    monkeypatch.setattr(_native, "default_context", lambda device=None: shared)
    monkeypatch.setattr(_native, "helper_contexts", lambda device, n, owner=None: helpers[:n])
    return mtm


@pytest.mark.parametrize("name", gp.GOLDEN_MATCH)
def test_match_templates_golden(mock_mtm, golden, name):
    gp.test_match_templates_golden(mock_mtm, golden, name)


@pytest.mark.parametrize("name", ["t3_full", "t3_searchbox"])
def test_tutorial3_answers(mock_mtm, golden, name):
    gp.test_match_templates_tutorial3_fullres(mock_mtm, golden, name)


@pytest.mark.parametrize("name", ["c1_fish256_find", "synth_row_map", "synth_col_map"])
def test_find_matches_golden(mock_mtm, golden, name):
    gp.test_find_matches_golden(mock_mtm, golden, name)


def test_nms_plumbing(mock_mtm, golden):
    gp.test_nms_demo_and_random(mock_mtm, golden)


def test_validation(mock_mtm):
    gp.test_validation_errors_match_reference(mock_mtm)


@pytest.mark.parametrize("method,thr", [(1, 0.35), (3, 0.92), (2, 0.0), (4, 0.0)])
def test_other_methods(mock_mtm, method, thr):
    gp.test_match_templates_other_methods_vs_port(mock_mtm, method, thr)


def test_rare_corners_split_find_and_nms(mock_mtm):
    """N_object = 0, a negative NMS threshold (cv2.error from NMSBoxes) and method 0 take the reference's two-stage
    route (MTM/__init__.py:289-296, MTM/NMS.py:53-82)."""
    import cv2
    from oracle import mtm_port, synth
    rng = np.random.default_rng(9)
    temps = [("a", synth.make_template(rng, 20, 24)), ("b", synth.make_template(rng, 16, 30))]
    img, _ = synth.make_scene(150, 200, [t[1] for t in temps], 3, seed=9)
    assert mock_mtm.matchTemplates(temps, img, N_object=0) == mtm_port.match_templates(temps, img, N_object=0) == []
    with pytest.raises(cv2.error):
        mtm_port.match_templates(temps, img, score_threshold=-0.5)
    with pytest.raises(cv2.error):
        mock_mtm.matchTemplates(temps, img, score_threshold=-0.5)
    one = mock_mtm.matchTemplates(temps[:1], img[:40, :60], score_threshold=-0.5, N_object=1)     # single hit: no threshold check
    assert [(h[0], h[1]) for h in one] == [(h[0], h[1]) for h in mtm_port.match_templates(temps[:1], img[:40, :60], score_threshold=-0.5, N_object=1)]
    with pytest.raises(ValueError, match="TM_SQDIFF is not supported"):
        mock_mtm.matchTemplates(temps, img, method=0)
    # uint16 / float32 / mixed inputs are routed like the reference casts them
    for im, ts in ((img.astype(np.uint16) * 200, [(n, t.astype(np.uint16) * 200) for n, t in temps]),
                   (img.astype(np.float32), temps)):
        got = mock_mtm.matchTemplates(ts, im, score_threshold=0.5)
        want = mtm_port.match_templates(ts, im, score_threshold=0.5)
        assert [(h[0], h[1]) for h in got] == [(h[0], h[1]) for h in want] and len(want) >= 4


def test_sharded_entry_points_without_a_process_group(mock_mtm):
    """world size 1: the sharded wrappers reduce to the plain calls (bodies of tests/test_gpu_sharded.py)."""
    import test_gpu_sharded as gs
    gs.test_sharded_world1_equals_match_templates(mock_mtm)


class _FakeDeviceArray:
    """A numpy array dressed as a CUDA array (host pointer in ``__cuda_array_interface__``): lets the CPU tests drive the
    device-image route of the host layer; the test double reads the pixels back through the pointer."""

    def __init__(self, arr):
        self._arr = arr
        self.__cuda_array_interface__ = {"shape": arr.shape, "typestr": arr.dtype.str, "data": (arr.ctypes.data, False),
                                         "strides": None if arr.flags.c_contiguous else arr.strides, "version": 3}


def test_device_resident_images_take_the_device_route(mock_mtm):
    """Objects exposing __cuda_array_interface__ (PyTorch / CuPy arrays) are searched where they are: same results as
    the numpy image, searchBox crops are pointer arithmetic, and nothing is cast on the host."""
    from mtm_b200 import _native
    from oracle import synth
    rng = np.random.default_rng(41)
    temps = [("a", synth.make_template(rng, 20, 24)), ("b", synth.make_template(rng, 16, 30))]
    img, _ = synth.make_scene(150, 200, [t[1] for t in temps], 3, seed=41)
    shared = _native.default_context()
    for image, ts in ((img, temps), (img.astype(np.float32), [(n, t.astype(np.float32)) for n, t in temps]),
                      (img.astype(np.uint16) * 100, [(n, t.astype(np.uint16) * 100) for n, t in temps]),
                      (np.stack([img, 255 - img, img[::-1]], axis=2), [(n, np.stack([t, 255 - t, t[::-1]], axis=2)) for n, t in temps])):
        image = np.ascontiguousarray(image)
        dev = _FakeDeviceArray(image)
        for kw in (dict(score_threshold=0.5), dict(N_object=1), dict(score_threshold=0.5, searchBox=(13, 7, 150, 120))):
            shared.calls.clear()
            got = mock_mtm.matchTemplates(ts, dev, **kw)
            assert "set_image_device" in shared.calls and "set_image" not in shared.calls
            want = mock_mtm.matchTemplates(ts, image, **kw)
            assert [(h[0], h[1], float(h[2])) for h in got] == [(h[0], h[1], float(h[2])) for h in want] and len(want) >= 1
        assert np.array_equal(mock_mtm.computeScoreMap(ts[0][1], dev), mock_mtm.computeScoreMap(ts[0][1], image))
        assert [(h[0], h[1]) for h in mock_mtm.findMatches(ts, dev)] == [(h[0], h[1]) for h in mock_mtm.findMatches(ts, image)]
    batch = mock_mtm.matchTemplatesBatch(temps, [_FakeDeviceArray(img), img, _FakeDeviceArray(np.ascontiguousarray(img[::-1]))])
    assert [[(h[0], h[1]) for h in hits] for hits in batch] == \
        [[(h[0], h[1]) for h in mock_mtm.matchTemplates(temps, im)] for im in (img, img, np.ascontiguousarray(img[::-1]))]
    aug = mock_mtm.matchTemplatesAugmented(temps, _FakeDeviceArray(img), ("identity", "rot180"))
    assert [(h[0], h[1]) for h in aug] == [(h[0], h[1]) for h in mock_mtm.matchTemplatesAugmented(temps, img, ("identity", "rot180"))]
    # combinations that would need the reference's float32 cast of the IMAGE are refused, not cast on the host
    with pytest.raises(NotImplementedError, match="device-resident"):
        mock_mtm.matchTemplates([("a", temps[0][1].astype(np.float32))], _FakeDeviceArray(img))
    with pytest.raises(NotImplementedError, match="host images"):
        mock_mtm.matchTemplatesPyramid(temps, _FakeDeviceArray(img), downscale=2)
    with pytest.raises(ValueError, match="64-bit"):
        mock_mtm.matchTemplates(temps, _FakeDeviceArray(img.astype(np.float64)))
    with pytest.raises(ValueError, match="larger than searchBox"):
        mock_mtm.matchTemplates(temps, _FakeDeviceArray(img), searchBox=(0, 0, 10, 10))
    view = _native.as_image(_FakeDeviceArray(img))
    assert view[5:25, 8:40].shape == (20, 32) and view[5:25, 8:40].ptr == img[5:25, 8:40].ctypes.data and view.strides == img.strides
    with pytest.raises(TypeError):
        view[::2, :]
