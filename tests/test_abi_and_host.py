"""CPU tests: the C-ABI library loads and exports every declared symbol; the host
layer validates like the reference and fails loudly (no fallback) without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "mtm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mtm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(mtm):
    import ctypes
    from mtm_b200 import _native
    lib = _native.load()
    declared = _declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libmtm_b200.so lacks %s" % name
    assert sorted(_native.exported_symbols()) == declared            # the binding covers the whole header
    assert lib.mtm_abi_version() == 2
    assert ctypes.sizeof(ctypes.c_int32) * 5 + 4 == _native.HIT_DTYPE.itemsize


def test_no_gpu_means_loud_failure_not_fallback(mtm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mtm_b200 import _native
    with pytest.raises(_native.NativeError, match="no CUDA device|no CPU fallback"):
        _native.Context(0)
    img = np.zeros((32, 32), np.uint8)
    with pytest.raises(RuntimeError):
        mtm.matchTemplates([("a", img[:8, :8])], img)


def test_validation_precedes_device_work(mtm):
    img = np.zeros((50, 60), np.uint8)
    t = np.zeros((10, 10), np.uint8)
    with pytest.raises(ValueError, match="Maximal overlap"):
        mtm.matchTemplates([("a", t)], img, maxOverlap=-0.1)
    with pytest.raises(TypeError, match="N_object must be an integer"):
        mtm.findMatches([("a", t)], img, N_object=2.0)
    with pytest.raises(ValueError, match="height of 0"):
        mtm.findMatches([("a", t)], np.zeros((0, 5), np.uint8))
    with pytest.raises(ValueError, match="width of 0"):
        mtm.findMatches([("a", np.zeros((3, 0), np.uint8))], img)
    with pytest.raises(ValueError, match="'big' at index 1 in the list of templates is larger than image"):
        mtm.findMatches([("a", t), ("big", np.zeros((51, 5), np.uint8))], img)
    with pytest.raises(ValueError, match="larger than searchBox"):
        mtm.findMatches([("a", t)], img, searchBox=(0, 0, 9, 9))
    with pytest.raises(ValueError, match="list of tuples"):
        mtm.findMatches([("a",)], img)
    with pytest.raises(ValueError, match="64-bit images not supported"):
        mtm.computeScoreMap(t, img.astype(np.float64))
    assert mtm.NMS([]) == [] and mtm.NMS([("a", (0, 0, 1, 1), 0.1)]) == [("a", (0, 0, 1, 1), 0.1)]


def test_api_surface_matches_reference(mtm):
    import inspect
    assert mtm.__version__ == "2.0.1" and mtm.__all__ == ["NMS"]
    sig = inspect.signature(mtm.matchTemplates)
    assert list(sig.parameters)[:7] == ["listTemplates", "image", "method", "N_object", "score_threshold", "maxOverlap", "searchBox"]
    assert sig.parameters["maxOverlap"].default == 0.25 and sig.parameters["method"].default == 5
    assert sig.parameters["N_object"].default == float("inf") and sig.parameters["score_threshold"].default == 0.5
    sig = inspect.signature(mtm.NMS)
    assert list(sig.parameters)[:5] == ["listHit", "scoreThreshold", "sortAscending", "N_object", "maxOverlap"]
    assert sig.parameters["maxOverlap"].default == 0.5
    assert list(inspect.signature(mtm.computeScoreMap).parameters)[:4] == ["template", "image", "method", "mask"]
    assert list(inspect.signature(mtm.findMatches).parameters)[:6] == ["listTemplates", "image", "method", "N_object", "score_threshold", "searchBox"]
    from MTM.NMS import NMS
    assert NMS is mtm.NMS
    assert callable(mtm.drawBoxesOnRGB) and callable(mtm.drawBoxesOnGray)


def test_dtype_routing_of_the_host_layer(mtm):
    """The dtype / mask policy (MTM/__init__.py:67-88, 207-222) without a device: which arrays reach the C ABI."""
    import warnings
    from mtm_b200 import _native, api
    img8, t8 = np.zeros((20, 30), np.uint8), np.zeros((5, 6), np.uint8)
    img16, t16 = img8.astype(np.uint16), t8.astype(np.uint16)
    # uint8 stays uint8; anything else becomes float32, the image cast once and shared
    _, arrs, img, _ = api._prepare([("a", t8), ("b", t8)], img8, 5)
    assert img is img8 and all(a.dtype == np.uint8 for a in arrs)
    _, arrs, img, _ = api._prepare([("a", t8.astype(np.float32)), ("b", t16)], img16, 5)
    assert img.dtype == np.float32 and all(a.dtype == np.float32 for a in arrs)
    # all-uint16 grayscale: handed over untouched (MTM_U16), unless a usable mask forces the float32 route
    _, arrs, img, masks = api._prepare([("a", t16), ("b", t16)], img16, 5)
    assert img is img16 and arrs[0] is t16 and masks == [None, None]
    assert _native._dtype_code(img16) == _native.MTM_U16 == 2
    _, arrs, img, masks = api._prepare([("a", t16, np.ones_like(t16))], img16, 3)
    assert img.dtype == np.float32 and arrs[0].dtype == np.float32 and masks[0].dtype == np.float32
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        _, arrs, img, masks = api._prepare([("a", t16, np.ones_like(t16))], img16, 5)     # mask unusable with method 5
    assert img is img16 and masks == [None] and len(w) == 1
    rgb16 = np.zeros((20, 30, 3), np.uint16)
    _, arrs, img, _ = api._prepare([("a", rgb16[:5, :6])], rgb16, 5)                         # 16-bit RGB: float32 route
    assert img.dtype == np.float32
    with pytest.raises(ValueError, match="64-bit"):
        api._prepare([("a", t8.astype(np.float64))], img8, 5)
    with pytest.raises(TypeError):
        _native._dtype_code(np.zeros(3, np.int32))


def test_augmented_and_pyramid_validate_before_device_work(mtm):
    """SURVEY §8 f3 front ends: the reference's checks (on the EXPANDED list) and the extra ones, without a GPU."""
    img = np.zeros((40, 100), np.uint8)
    wide = np.zeros((10, 60), np.uint8)                      # fits, but its 90-degree rotations (60 x 10) do not
    with pytest.raises(ValueError, match="'w_rot90' at index 1 in the list of templates is larger than image"):
        mtm.matchTemplatesAugmented([("w", wide)], img)
    with pytest.raises(ValueError, match="unknown transform"):
        mtm.matchTemplatesAugmented([("w", wide)], img, transforms=("identity", "rot45"))
    with pytest.raises(ValueError, match="Maximal overlap"):
        mtm.matchTemplatesAugmented([("w", wide)], img, transforms=("identity",), maxOverlap=2)
    with pytest.raises(NotImplementedError, match="masks"):
        mtm.matchTemplatesAugmented([("w", wide, np.ones_like(wide))], img, transforms=("identity", "fliplr"))
    with pytest.raises(NotImplementedError, match="all be uint8 or all be float32"):
        mtm.matchTemplatesAugmented([("w", wide.astype(np.uint16))], img, transforms=("identity", "fliplr"))
    with pytest.raises(ValueError, match="64-bit images not supported"):
        mtm.matchTemplatesPyramid([("w", wide.astype(np.float64))], img, downscale=2)
    with pytest.raises(ValueError, match="downscale must be an integer"):
        mtm.matchTemplatesPyramid([("w", wide)], img, downscale=17)
    with pytest.raises(ValueError, match="vanishes at downscale 16"):
        mtm.matchTemplatesPyramid([("w", wide)], img, downscale=16)
    with pytest.raises(ValueError, match="TM_SQDIFF is not supported"):
        mtm.matchTemplatesPyramid([("w", wide)], img, downscale=2, method=0)
    with pytest.raises(TypeError, match="N_object must be an integer"):
        mtm.matchTemplatesPyramid([("w", wide)], img, downscale=2, N_object=1.5)
    assert mtm.matchTemplatesAugmented([], img) == [] and mtm.matchTemplatesPyramid([], img) == []
    assert set(mtm.TRANSFORMS) == {"identity", "rot90", "rot180", "rot270", "fliplr", "flipud", "transpose", "antitranspose"}


def test_header_is_plain_c(tmp_path):
    """include/mtm_b200.h must compile as C (the boundary is a C ABI, not a C++ one) and link against the library."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "mtm_b200.h"\n'
                   'int main(void) {\n'
                   '    mtm_hit h = {0, 1, 2, 3, 4, 0.5f};\n'
                   '    mtm_counters c = {0};\n'
                   '    mtm_ctx* ctx = 0;\n'
                   '    int ok = sizeof(mtm_hit) == 24 && MTM_MAX_INFLIGHT == 8 && MTM_XF_ANTITRANSPOSE == 7 && MTM_U16 == 2;\n'
                   '    (void)h; (void)c;\n'
                   '    /* without a GPU mtm_create must fail loudly; with one it must succeed */\n'
                   '    int rc = mtm_create(0, &ctx);\n'
                   '    if (rc == MTM_OK) mtm_destroy(ctx); else if (!mtm_last_error(0)[0]) return 3;\n'
                   '    return ok && mtm_abi_version() == MTM_ABI_VERSION ? 0 : 2;\n'
                   '}\n')
    lib_dir = os.path.join(ROOT, "multitemplatematching-python_b200")
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-lmtm_b200", "-Wl,-rpath," + lib_dir], check=True, capture_output=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_draw_helpers_equal_the_reference(mtm):
    """drawBoxesOnRGB / drawBoxesOnGray (MTM/__init__.py:299-391): pixel-identical to the unmodified reference when it is
    present (build container), and to the cv2 primitives it calls otherwise."""
    import cv2
    from oracle import ref_loader
    rng = np.random.default_rng(2)
    gray = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    rgb = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    hits = [("head", (10, 20, 40, 30), np.float32(0.9)), ("tail", (90, 50, 50, 60), np.float32(0.7))]
    ref = ref_loader.load() if ref_loader.available() else None
    for image in (gray, rgb):
        for kw in (dict(), dict(boxThickness=3, showLabel=True, labelScale=0.7)):
            ours_rgb, ours_gray = mtm.drawBoxesOnRGB(image, hits, **kw), mtm.drawBoxesOnGray(image, hits, **kw)
            assert ours_rgb.shape == (120, 160, 3) and ours_gray.shape == (120, 160)
            assert not np.shares_memory(ours_rgb, image) and not np.shares_memory(ours_gray, image)
            if ref is not None:
                assert np.array_equal(ours_rgb, ref.drawBoxesOnRGB(image, hits, **kw))
                assert np.array_equal(ours_gray, ref.drawBoxesOnGray(image, hits, **kw))
            else:
                want = cv2.cvtColor(image, cv2.COLOR_GRAY2RGB) if image.ndim == 2 else image.copy()
                for label, (x, y, w, h), _ in hits:
                    cv2.rectangle(want, (x, y), (x + w, y + h), color=(255, 255, 0), thickness=kw.get("boxThickness", 2))
                    if kw.get("showLabel"):
                        cv2.putText(want, text=label, org=(x, y), fontFace=cv2.FONT_HERSHEY_SIMPLEX, fontScale=kw["labelScale"],
                                    color=(255, 255, 0), lineType=cv2.LINE_AA)
                assert np.array_equal(ours_rgb, want)


def test_hit_buffer_growth_and_strided_inputs_without_a_device(mtm):
    """Host plumbing of the ctypes layer with a stub library: MTM_ERR_CAPACITY grows the reusable hit buffer and the call
    is repeated; the returned hits are a private copy; strided image views are passed without a host copy when their
    rows are contiguous (the searchBox crop of MTM/__init__.py:140-144) and copied otherwise."""
    import ctypes
    import threading
    from mtm_b200 import _native

    class Stub:
        def __init__(self):
            self.calls = []

        def mtm_find_matches(self, h, method, n_object, thr, buf, cap, n_ptr):
            self.calls.append(cap)
            need = 10000
            n_ptr._obj.value = need
            if cap < need:
                return _native.MTM_ERR_CAPACITY
            arr = np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ctypes.c_uint8)), shape=(cap * 24,)).view(_native.HIT_DTYPE)
            arr[:need]["x"] = np.arange(need)
            return _native.MTM_OK

        def mtm_set_image(self, h, ptr, H, W, C, code, stride):
            self.calls.append(("image", ptr.value, H, W, C, code, stride))
            return _native.MTM_OK

        def mtm_last_error(self, h):
            return b"stub"

    ctx = _native.Context.__new__(_native.Context)
    ctx._lib, ctx._h, ctx.device, ctx.lock = Stub(), ctypes.c_void_p(1), 0, threading.RLock()
    ctx.close = lambda: None
    first = ctx.find_matches(5, -1, 0.5)
    assert ctx._lib.calls == [4096, 10000] and len(first) == 10000 and first["x"][-1] == 9999
    second = ctx.find_matches(5, -1, 0.5)
    assert ctx._lib.calls == [4096, 10000, 10000]                      # the grown buffer is kept
    second["x"][0] = -7
    assert first["x"][0] == 0 and not np.shares_memory(first, second)     # results are private copies
    ctx._lib.calls.clear()
    img = np.zeros((50, 64), np.uint8)
    ctx.set_image(img[5:25, 8:40])                                        # rows contiguous inside: no copy, row stride 64
    tag, ptr, H, W, C, code, stride = ctx._lib.calls[-1]
    assert (H, W, C, code, stride) == (20, 32, 1, _native.MTM_U8, 64) and ptr == img[5:25, 8:40].ctypes.data
    ctx.set_image(img[:, ::2])                                            # strided inside a row: compact copy
    tag, ptr, H, W, C, code, stride = ctx._lib.calls[-1]
    assert (H, W, stride) == (50, 32, 32)
    rgb = np.zeros((30, 40, 3), np.float32)
    ctx.set_image(rgb[2:12, 3:13])
    tag, ptr, H, W, C, code, stride = ctx._lib.calls[-1]
    assert (H, W, C, code, stride) == (10, 10, 3, _native.MTM_F32, 40 * 3 * 4)


def test_rendezvous_hands_the_id_of_rank0_to_every_rank(mtm, monkeypatch):
    """mtm_b200.rendezvous.exchange_id: the 128 id bytes made on rank 0 reach the other ranks over a socket on
    MASTER_PORT + 117 (no torch); world 1 needs no socket."""
    import socket
    import threading
    from mtm_b200 import _native, rendezvous
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    monkeypatch.setenv("MTM_B200_COMM_PORT", str(port))
    payload = bytes(range(128))
    assert len(payload) == _native.COMM_ID_BYTES
    got = [None] * 3

    def rank(r):
        got[r] = rendezvous.exchange_id(r, 3, (lambda: payload) if r == 0 else (lambda: b"never called"), timeout=30.0)

    threads = [threading.Thread(target=rank, args=(r,)) for r in (1, 2, 0)]     # the clients may come up before the listener
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert got == [payload] * 3
    assert rendezvous.exchange_id(0, 1, lambda: b"solo") == b"solo"
