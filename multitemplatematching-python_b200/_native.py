"""ctypes binding of libmtm_b200.so (include/mtm_b200.h).  No CPU fallback:
loading fails loudly when the library has not been built, and creating a
context fails loudly when there is no sm_100 GPU."""
import ctypes
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmtm_b200.so")

MTM_OK, MTM_ERR_INVALID, MTM_ERR_CUDA, MTM_ERR_CAPACITY, MTM_ERR_UNSUPPORTED, MTM_ERR_PEER = 0, -1, -2, -3, -4, -5
MTM_U8, MTM_F32, MTM_U16 = 0, 1, 2
PATH_AUTO, PATH_DIRECT, PATH_TENSOR = 0, 1, 2
OPT_PATH, OPT_TIME_NCC = 0, 1
# mtm_transform (include/mtm_b200.h)
XF_IDENTITY, XF_ROT90, XF_ROT180, XF_ROT270, XF_FLIPLR, XF_FLIPUD, XF_TRANSPOSE, XF_ANTITRANSPOSE = range(8)
MAX_DOWNSCALE = 16

HIT_DTYPE = np.dtype([("tmpl", "<i4"), ("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("score", "<f4")])
assert HIT_DTYPE.itemsize == 24


class Counters(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_int64), ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64),
                ("ncc_launches", ctypes.c_int64), ("ncc_ms", ctypes.c_double), ("tma_launches", ctypes.c_int64),
                ("hits_only_searches", ctypes.c_int64)]


# name -> (restype, argtypes); mirrors include/mtm_b200.h one to one
_P = ctypes.c_void_p
_SIGNATURES = {
    "mtm_abi_version": (ctypes.c_int, []),
    "mtm_device_count": (ctypes.c_int, []),
    "mtm_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_P)]),
    "mtm_destroy": (ctypes.c_int, [_P]),
    "mtm_last_error": (ctypes.c_char_p, [_P]),
    "mtm_set_stream": (ctypes.c_int, [_P, _P]),
    "mtm_synchronize": (ctypes.c_int, [_P]),
    "mtm_set_option": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int64]),
    "mtm_get_counters": (ctypes.c_int, [_P, ctypes.POINTER(Counters)]),
    "mtm_reset_counters": (ctypes.c_int, [_P]),
    "mtm_timer_begin": (ctypes.c_int, [_P]),
    "mtm_timer_end": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_float)]),
    "mtm_measure_i8_peak": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "mtm_set_image": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64]),
    "mtm_set_image_device": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64]),
    "mtm_set_templates": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32),
                                         ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int]),
    "mtm_set_templates_masked": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_P), ctypes.POINTER(_P),
                                                ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                                ctypes.c_int, ctypes.c_int]),
    "mtm_set_templates_transformed": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32),
                                                     ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                     ctypes.POINTER(ctypes.c_int32), ctypes.c_int]),
    "mtm_set_image_scaled": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                            ctypes.c_int]),
    "mtm_set_image_roi": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "mtm_score_map": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P, ctypes.c_int64]),
    "mtm_find_matches": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int64, ctypes.c_double, _P, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_int)]),
    "mtm_nms": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int64, ctypes.c_double,
                               _P, ctypes.POINTER(ctypes.c_int)]),
    "mtm_match_templates": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_double, _P,
                                           ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "mtm_match_templates_async": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                                 ctypes.c_int]),
    "mtm_match_templates_collect": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    # multi-GPU
    "mtm_comm_unique_id": (ctypes.c_int, [_P]),
    "mtm_comm_init_rank": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, ctypes.POINTER(_P)]),
    "mtm_comm_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_P)]),
    "mtm_comm_destroy": (ctypes.c_int, [_P]),
    "mtm_comm_info": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "mtm_comm_last_error": (ctypes.c_char_p, [_P]),
    "mtm_comm_allreduce_max": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double), ctypes.c_int]),
    "mtm_comm_barrier": (ctypes.c_int, [_P]),
    "mtm_match_templates_sharded": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_double,
                                                   ctypes.c_double, _P, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "mtm_gather_results": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                          ctypes.c_int, _P, _P]),
}
MAX_INFLIGHT = 8
COMM_ID_BYTES = 128
SLOT_HITS = 1024

_lib = None
_lib_lock = threading.Lock()


def exported_symbols():
    return sorted(_SIGNATURES)


def _prefer_bundled_nccl():
    """libmtm_b200.so resolves libnccl.so.2 with dlopen when a multi-GPU communicator is first made.  A process that also
    imports PyTorch must end up with ONE NCCL: torch links the copy bundled in the ``nvidia.nccl`` wheel, so when that wheel
    is installed (and MTM_B200_NCCL_LIB is not set) the library is pointed at the same file -- whichever of the two loads
    first, the other finds it.  Without the wheel the system libnccl.so.2 is used."""
    if "MTM_B200_NCCL_LIB" in os.environ:
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["MTM_B200_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def load():
    """dlopen libmtm_b200.so and declare every entry point (no GPU needed for this)."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "libmtm_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `python multitemplatematching-python_b200/build.py`. There is no CPU fallback." % LIB_PATH)
            _prefer_bundled_nccl()
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)      # AttributeError if the library lacks a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmtm_b200 error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


def _dtype_code(arr):
    if arr.dtype == np.uint8:
        return MTM_U8
    if arr.dtype == np.float32:
        return MTM_F32
    if arr.dtype == np.uint16:
        return MTM_U16
    raise TypeError("unsupported dtype %s" % arr.dtype)


class DeviceArray:
    """An image that already lives in HBM, seen through ``__cuda_array_interface__`` (PyTorch CUDA tensors, CuPy
    arrays, ...).  Exposes just what the host layer reads from an image (``shape``, ``dtype``, ``ndim``, ``strides``) and
    the 2-D slicing of the searchBox crop (MTM/__init__.py:140-144); the pixels are never touched on the host.
    The producer's work on the array must be complete before it is handed over (synchronise its stream)."""

    def __init__(self, ptr, shape, dtype, strides, owner):
        self.ptr, self.shape, self.dtype, self.owner = int(ptr), tuple(int(v) for v in shape), np.dtype(dtype), owner
        if strides is None:                          # C-contiguous
            strides, step = [], self.dtype.itemsize
            for extent in reversed(self.shape):
                strides.append(step)
                step *= extent
            strides = strides[::-1]
        self.strides = tuple(int(v) for v in strides)
        self.ndim = len(self.shape)

    @classmethod
    def wrap(cls, obj):
        cai = obj.__cuda_array_interface__
        if cai.get("mask") is not None:
            raise ValueError("masked device arrays are not supported")
        ptr, read_only = cai["data"]
        return cls(ptr, cai["shape"], np.dtype(cai["typestr"]), cai.get("strides"), obj)

    def __getitem__(self, key):
        if not (isinstance(key, tuple) and len(key) == 2 and all(isinstance(k, slice) for k in key)) or self.ndim < 2:
            raise TypeError("device images support only image[y0:y1, x0:x1] slicing")
        (y0, y1, ys), (x0, x1, xs) = key[0].indices(self.shape[0]), key[1].indices(self.shape[1])
        if ys != 1 or xs != 1:
            raise TypeError("device images support only unit-step slices")
        shape = (max(y1 - y0, 0), max(x1 - x0, 0)) + self.shape[2:]
        return DeviceArray(self.ptr + y0 * self.strides[0] + x0 * self.strides[1], shape, self.dtype, self.strides, self.owner)


def as_image(image):
    """numpy arrays pass through; objects exposing ``__cuda_array_interface__`` become a ``DeviceArray`` view."""
    if isinstance(image, (np.ndarray, DeviceArray)) or not hasattr(image, "__cuda_array_interface__"):
        return image
    return DeviceArray.wrap(image)


class Context:
    """One CUDA stream + device workspaces on one GPU.  Not re-entrant (guarded by a lock)."""

    def __init__(self, device=0):
        self._lib = load()
        handle = _P()
        rc = self._lib.mtm_create(int(device), ctypes.byref(handle))
        if rc != MTM_OK:
            raise NativeError(rc, (self._lib.mtm_last_error(None) or b"").decode())
        self._h = handle
        self.device = int(device)
        self.lock = threading.RLock()
        self._keep = None
        self._hit_buf = None

    _hit_buf = None                       # reusable host buffer of the hit-list calls (guarded by `lock`)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mtm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != MTM_OK:
            raise NativeError(rc, (self._lib.mtm_last_error(self._h) or b"").decode())

    # -- plumbing ---------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.mtm_set_stream(self._h, _P(cuda_stream_ptr or 0)))

    def synchronize(self):
        self._check(self._lib.mtm_synchronize(self._h))

    def set_path(self, path):
        self._check(self._lib.mtm_set_option(self._h, OPT_PATH, int(path)))

    def counters(self):
        c = Counters()
        self._check(self._lib.mtm_get_counters(self._h, ctypes.byref(c)))
        return {"kernel_launches": c.kernel_launches, "h2d_bytes": c.h2d_bytes, "d2h_bytes": c.d2h_bytes,
                "ncc_launches": c.ncc_launches, "ncc_ms": c.ncc_ms, "tma_launches": c.tma_launches,
                "hits_only_searches": c.hits_only_searches}

    def set_time_ncc(self, on):
        self._check(self._lib.mtm_set_option(self._h, OPT_TIME_NCC, 1 if on else 0))

    def reset_counters(self):
        self._check(self._lib.mtm_reset_counters(self._h))

    def timer_begin(self):
        self._check(self._lib.mtm_timer_begin(self._h))

    def timer_end(self):
        ms = ctypes.c_float()
        self._check(self._lib.mtm_timer_end(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def measure_i8_peak(self, n_cols=256, iters=4000):
        """Dense u8 x u8 MAC rate of the tensor pipe in TMAC/s (mtm_measure_i8_peak)."""
        out = ctypes.c_double()
        self._check(self._lib.mtm_measure_i8_peak(self._h, int(n_cols), int(iters), ctypes.byref(out)))
        return float(out.value)

    # -- inputs -----------------------------------------------------------------
    @staticmethod
    def _image_view(image):
        """Returns (array kept alive, H, W, C, row_stride) with rows internally contiguous."""
        if image.ndim not in (2, 3):
            raise ValueError("image must be 2-D (grayscale) or 3-D (H, W, C)")
        C = 1 if image.ndim == 2 else image.shape[2]
        item = image.dtype.itemsize
        inner_ok = (image.strides[1] == item * C) and (image.ndim == 2 or image.strides[2] == item)
        if not inner_ok or image.strides[0] < image.shape[1] * C * item:
            image = np.ascontiguousarray(image)
        return image, image.shape[0], image.shape[1], C, image.strides[0]

    def set_image(self, image):
        if isinstance(image, DeviceArray):           # pixels already in HBM: device-to-device copy into the tile layout, no PCIe traffic
            if image.ndim not in (2, 3):
                raise ValueError("image must be 2-D (grayscale) or 3-D (H, W, C)")
            C = 1 if image.ndim == 2 else image.shape[2]
            item = image.dtype.itemsize
            if image.strides[1] != item * C or (image.ndim == 3 and image.strides[2] != item) or image.strides[0] < image.shape[1] * C * item:
                raise ValueError("device images must have contiguous rows (strides %r)" % (image.strides,))
            self._keep = image
            self.set_image_device(image.ptr, image.shape[0], image.shape[1], C, image.strides[0], _dtype_code(image))
            return
        arr, H, W, C, stride = self._image_view(image)
        self._check(self._lib.mtm_set_image(self._h, _P(arr.ctypes.data), H, W, C, _dtype_code(arr), stride))

    def set_image_device(self, dev_ptr, H, W, C, row_stride, dtype=MTM_U8):
        self._check(self._lib.mtm_set_image_device(self._h, _P(dev_ptr), H, W, C, dtype, row_stride))

    def set_templates(self, templates):
        arrs = [np.ascontiguousarray(t) for t in templates]
        n = len(arrs)
        if n == 0:
            raise ValueError("empty template list")
        C = 1 if arrs[0].ndim == 2 else arrs[0].shape[2]
        code = _dtype_code(arrs[0])
        for a in arrs:
            if (1 if a.ndim == 2 else a.shape[2]) != C or _dtype_code(a) != code:
                raise ValueError("templates must share dtype and channel count")
        ptrs = (_P * n)(*[a.ctypes.data for a in arrs])
        hs = (ctypes.c_int32 * n)(*[a.shape[0] for a in arrs])
        ws = (ctypes.c_int32 * n)(*[a.shape[1] for a in arrs])
        self._check(self._lib.mtm_set_templates(self._h, n, ptrs, hs, ws, C, code))
        self._tmpl_shapes = [a.shape[:2] for a in arrs]

    def set_templates_masked(self, templates, masks):
        arrs = [np.ascontiguousarray(t) for t in templates]
        mks = [np.ascontiguousarray(m) for m in masks]
        n = len(arrs)
        C = 1 if arrs[0].ndim == 2 else arrs[0].shape[2]
        code = _dtype_code(arrs[0])
        for a, m in zip(arrs, mks):
            if (1 if a.ndim == 2 else a.shape[2]) != C or _dtype_code(a) != code or m.shape != a.shape or m.dtype != a.dtype:
                raise ValueError("templates and masks must share shape, dtype and channel count")
        tp = (_P * n)(*[a.ctypes.data for a in arrs])
        mp = (_P * n)(*[m.ctypes.data for m in mks])
        hs = (ctypes.c_int32 * n)(*[a.shape[0] for a in arrs])
        ws = (ctypes.c_int32 * n)(*[a.shape[1] for a in arrs])
        self._check(self._lib.mtm_set_templates_masked(self._h, n, tp, mp, hs, ws, C, code))

    # -- on-device augmentation / pyramid helpers ------------------
    def set_templates_transformed(self, templates, ops, downscale=1):
        """Template list := [transform op of the downscaled base for base in templates for op in ops]."""
        arrs = [np.ascontiguousarray(t) for t in templates]
        n = len(arrs)
        if n == 0:
            raise ValueError("empty template list")
        C = 1 if arrs[0].ndim == 2 else arrs[0].shape[2]
        code = _dtype_code(arrs[0])
        for a in arrs:
            if (1 if a.ndim == 2 else a.shape[2]) != C or _dtype_code(a) != code:
                raise ValueError("templates must share dtype and channel count")
        ops = [int(o) for o in ops]
        ptrs = (_P * n)(*[a.ctypes.data for a in arrs])
        hs = (ctypes.c_int32 * n)(*[a.shape[0] for a in arrs])
        ws = (ctypes.c_int32 * n)(*[a.shape[1] for a in arrs])
        op_arr = (ctypes.c_int32 * len(ops))(*ops)
        self._check(self._lib.mtm_set_templates_transformed(self._h, n, ptrs, hs, ws, C, code, len(ops), op_arr, int(downscale)))

    def set_image_scaled(self, image, downscale):
        """Upload the full-resolution image (kept resident) and search its INTER_AREA reduction."""
        arr, H, W, C, stride = self._image_view(image)
        self._check(self._lib.mtm_set_image_scaled(self._h, _P(arr.ctypes.data), H, W, C, _dtype_code(arr), stride, int(downscale)))

    def set_image_roi(self, x, y, w, h):
        """Current image := region of the resident full-resolution image (device-side searchBox)."""
        self._check(self._lib.mtm_set_image_roi(self._h, int(x), int(y), int(w), int(h)))

    # -- hot path ---------------------------------------------------------------
    def score_map(self, tmpl, method, map_shape):
        out = np.empty(map_shape, np.float32)
        self._check(self._lib.mtm_score_map(self._h, int(tmpl), int(method), _P(out.ctypes.data), out.size))
        return out

    def _hits_call(self, fn, args):
        """Calls a hit-list entry point with the context's reusable output buffer (grown on MTM_ERR_CAPACITY);
        returns a private copy of the hits."""
        while True:
            buf = self._hit_buf
            if buf is None:
                buf = self._hit_buf = np.empty(4096, HIT_DTYPE)
            n = ctypes.c_int(0)
            rc = fn(self._h, *args, _P(buf.ctypes.data), buf.shape[0], ctypes.byref(n))
            if rc == MTM_ERR_CAPACITY:
                self._hit_buf = np.empty(max(2 * buf.shape[0], int(n.value)), HIT_DTYPE)
                continue
            self._check(rc)
            return buf[: n.value].copy()

    def find_matches(self, method, n_object, score_threshold):
        return self._hits_call(self._lib.mtm_find_matches, (int(method), int(n_object), float(score_threshold)))

    def match_templates(self, method, n_object, score_threshold, max_overlap):
        return self._hits_call(self._lib.mtm_match_templates,
                               (int(method), int(n_object), float(score_threshold), float(max_overlap)))

    def match_templates_async(self, method, n_object, score_threshold, max_overlap, slot):
        """Enqueue search + NMS of the current image/templates; results via match_templates_collect(slot)."""
        self._check(self._lib.mtm_match_templates_async(self._h, int(method), int(n_object), float(score_threshold),
                                                        float(max_overlap), int(slot)))

    def match_templates_collect(self, slot):
        """Hits of an earlier match_templates_async(slot); None when that image needs the synchronous path."""
        buf = np.empty(1024, HIT_DTYPE)
        n = ctypes.c_int(0)
        rc = self._lib.mtm_match_templates_collect(self._h, int(slot), _P(buf.ctypes.data), 1024, ctypes.byref(n))
        if rc == MTM_ERR_CAPACITY:
            return None
        self._check(rc)
        return buf[: n.value]

    def match_templates_sharded(self, comm, tmpl_base, n_local, method, n_object, score_threshold, max_overlap):
        """Template cut (mtm_match_templates_sharded): this context holds the image and its slice of the template list;
        collective over ``comm``; every rank gets the same list, ``tmpl`` indexing the whole template list."""
        return self._hits_call(lambda h, *a: self._lib.mtm_match_templates_sharded(h, comm._h, *a),
                               (int(tmpl_base), int(n_local), int(method), int(n_object), float(score_threshold), float(max_overlap)))

    def nms(self, hits, score_threshold, sort_ascending, n_object, max_overlap):
        hits = np.ascontiguousarray(hits, HIT_DTYPE)
        n = hits.shape[0]
        keep = np.empty(max(n, 1), np.int32)
        nk = ctypes.c_int(0)
        self._check(self._lib.mtm_nms(self._h, _P(hits.ctypes.data), n, float(score_threshold), int(bool(sort_ascending)),
                                      int(n_object), float(max_overlap), _P(keep.ctypes.data), ctypes.byref(nk)))
        return keep[: nk.value]


class Comm:
    """One rank's endpoint of a communicator (mtm_comm).  ``Comm.init_rank`` = one process per GPU (the NCCL id made by
    ``Comm.unique_id()`` on rank 0 has to reach the other ranks: see ``rendezvous.py``); ``Comm.create(devices)`` = one
    process driving several endpoints, each from its own thread (all devices equal: in-process loop-back)."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib
        w, r, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        lib.mtm_comm_info(handle, ctypes.byref(w), ctypes.byref(r), ctypes.byref(d))
        self.world, self.rank, self.device = w.value, r.value, d.value

    @staticmethod
    def unique_id():
        lib = load()
        buf = ctypes.create_string_buffer(COMM_ID_BYTES)
        rc = lib.mtm_comm_unique_id(buf)
        if rc != MTM_OK:
            raise NativeError(rc, (lib.mtm_comm_last_error(None) or b"").decode())
        return buf.raw

    @classmethod
    def init_rank(cls, device, world, rank, unique_id=None):
        lib = load()
        handle = _P()
        idbuf = ctypes.create_string_buffer(unique_id, COMM_ID_BYTES) if unique_id is not None else None
        rc = lib.mtm_comm_init_rank(int(device), int(world), int(rank), idbuf, ctypes.byref(handle))
        if rc != MTM_OK:
            raise NativeError(rc, (lib.mtm_comm_last_error(None) or b"").decode())
        return cls(handle, lib)

    @classmethod
    def create(cls, devices):
        lib = load()
        n = len(devices)
        devs = (ctypes.c_int * n)(*[int(d) for d in devices])
        handles = (_P * n)()
        rc = lib.mtm_comm_create(n, devs, handles)
        if rc != MTM_OK:
            raise NativeError(rc, (lib.mtm_comm_last_error(None) or b"").decode())
        return [cls(_P(handles[i]), lib) for i in range(n)]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mtm_comm_destroy(self._h)
            self._h = None

    def _check(self, rc):
        if rc != MTM_OK:
            raise NativeError(rc, (self._lib.mtm_comm_last_error(self._h) or b"").decode())

    def allreduce_max(self, values):
        arr = (ctypes.c_double * len(values))(*[float(v) for v in values])
        self._check(self._lib.mtm_comm_allreduce_max(self._h, arr, len(values)))
        return list(arr)

    def barrier(self):
        self._check(self._lib.mtm_comm_barrier(self._h))

    def gather_results(self, entries, images_per_rank, hits_per_image):
        """``entries`` = this rank's submissions in image order, as (Context, slot) pairs (mtm_match_templates_async).  Returns
        (hits, counts): a (world * images_per_rank, hits_per_image) HIT_DTYPE array and the per-entry counts (-1 unused,
        -2 does not fit / outside the fused fast path), identical on every rank."""
        n = len(entries)
        ctxs = (_P * max(n, 1))(*[c._h for c, _ in entries])
        slots = (ctypes.c_int * max(n, 1))(*[int(s) for _, s in entries])
        total = self.world * int(images_per_rank)
        hits = np.zeros((total, int(hits_per_image)), HIT_DTYPE)
        counts = np.zeros(total, np.int32)
        self._check(self._lib.mtm_gather_results(self._h, n, ctxs, slots, int(images_per_rank), int(hits_per_image),
                                                 _P(hits.ctypes.data), _P(counts.ctypes.data)))
        return hits, counts


_default = {}
_helpers = {}
_default_lock = threading.Lock()


def local_device():
    """CUDA ordinal of this process: MTM_B200_DEVICE when set, else LOCAL_RANK folded into the VISIBLE device count (a
    launcher that gives every rank its own CUDA_VISIBLE_DEVICES leaves one visible device per rank: ordinal 0)."""
    if "MTM_B200_DEVICE" in os.environ:
        return int(os.environ["MTM_B200_DEVICE"])
    n = load().mtm_device_count()
    return int(os.environ.get("LOCAL_RANK", "0")) % max(n, 1)


def default_context(device=None):
    """Lazily created per-device context used by the module-level API."""
    if device is None:
        device = local_device()
    with _default_lock:
        ctx = _default.get(device)
        if ctx is None:
            ctx = _default[device] = Context(device)
    return ctx


def helper_contexts(device, n, owner=None):
    """``n`` additional lazily created contexts (= CUDA streams with their own workspaces) on ``device``, used next
    to the caller's context by the batch entry points so that uploads overlap the searches of another stream.
    ``owner``: the caller's own context when it is not the default one -- helpers are per owner, so that two host
    threads driving two contexts of one device never share (and wait for) a helper."""
    with _default_lock:
        have = _helpers.setdefault((int(device), id(owner) if owner is not None else None), [])
        while len(have) < n:
            have.append(Context(device))
        return have[:n]
