"""TEST DOUBLE of ``mtm_b200._native.Context`` (CPU, oracle-backed) for the ``-m "not gpu"`` host-logic tests.

It states, in numpy and through ``oracle/``, what every C-ABI entry point of ``include/mtm_b200.h`` promises, so that
the Python host layer above the ABI (``api.py``, ``augment.py``: label mapping, offsets, search boxes, the
coarse-to-fine sequencing) can be exercised without a GPU, and so that the read-back tricks of the GPU tests are
checked before they reach the GPU box.  It is never importable from the product: it lives in ``tests/`` and is only
installed by ``monkeypatch`` inside CPU tests.
"""
import threading

import numpy as np

from oracle import augment_port as ap, mtm_port, ncc_exact

HIT_DTYPE = np.dtype([("tmpl", "<i4"), ("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("score", "<f4")])
_XF_NAMES = ["identity", "rot90", "rot180", "rot270", "fliplr", "flipud", "transpose", "antitranspose"]


class MockContext:
    def __init__(self, device=0):
        self.device = device
        self.lock = threading.RLock()
        self.image = None
        self.full = None
        self.templates = []
        self.calls = []                      # names of the entry points used, in order
        self.slots = {}

    def close(self):
        pass

    def synchronize(self):
        pass

    # -- inputs ---------------------------------------------------------------------
    def set_image(self, image):
        from mtm_b200._native import DeviceArray
        if isinstance(image, DeviceArray):           # CPU tests hand over HOST pointers dressed as device arrays
            self.calls.append("set_image_device")
            import ctypes
            C = 1 if image.ndim == 2 else image.shape[2]
            item = image.dtype.itemsize
            assert image.strides[1] == item * C and image.strides[0] >= image.shape[1] * C * item
            rows = image.shape[0]
            raw = np.ctypeslib.as_array(ctypes.cast(image.ptr, ctypes.POINTER(ctypes.c_uint8)), shape=((rows - 1) * image.strides[0] + image.shape[1] * C * item,))
            self.image = np.stack([raw[r * image.strides[0]: r * image.strides[0] + image.shape[1] * C * item].view(image.dtype).reshape(image.shape[1:])
                                   for r in range(rows)]).copy()
            return
        self.calls.append("set_image")
        self.image = np.ascontiguousarray(image)

    def set_templates(self, templates):
        self.calls.append("set_templates")
        self.templates = [np.ascontiguousarray(t) for t in templates]

    def set_templates_transformed(self, templates, ops, downscale=1):
        self.calls.append("set_templates_transformed")
        self.templates = [np.ascontiguousarray(ap.HOST_TRANSFORMS[_XF_NAMES[int(op)]](ap.area_downscale(t, downscale)))
                          for t in templates for op in ops]

    def set_image_scaled(self, image, downscale):
        self.calls.append("set_image_scaled")
        self.full = np.ascontiguousarray(image)
        self.image = ap.area_downscale(self.full, downscale)

    def set_image_roi(self, x, y, w, h):
        self.calls.append("set_image_roi")
        from mtm_b200._native import NativeError
        if self.full is None:
            raise NativeError(-1, "mtm_set_image_roi: no full-resolution image resident")
        H, W = self.full.shape[:2]
        if not (0 <= x and 0 <= y and w > 0 and h > 0 and x + w <= W and y + h <= H):
            raise NativeError(-1, "mtm_set_image_roi: region outside the image")
        self.image = np.ascontiguousarray(self.full[y:y + h, x:x + w])

    # -- hot path -------------------------------------------------------------------
    def score_map(self, tmpl, method, map_shape):
        self.calls.append("score_map")
        image, t = self.image, self.templates[tmpl]
        if image.dtype == np.uint16:                     # the reference's uint16 -> float32 cast
            image, t = image.astype(np.float32), t.astype(np.float32)
        out = ncc_exact.match_template_exact(image, t, method=method, use_fft=False)
        assert out.shape == tuple(map_shape)
        return out

    def _raw(self, hits):
        raw = np.zeros(len(hits), HIT_DTYPE)
        for k, (t, (x, y, w, h), s) in enumerate(hits):
            raw[k] = (t, x, y, w, h, s)
        return raw

    def find_matches(self, method, n_object, score_threshold):
        self.calls.append("find_matches")
        labelled = [(i, t) for i, t in enumerate(self.templates)]
        n = 1 if n_object == 1 else mtm_port.INF
        return self._raw(mtm_port.find_matches(labelled, self.image, method, n, score_threshold, workers=1))

    def match_templates(self, method, n_object, score_threshold, max_overlap):
        self.calls.append("match_templates")
        labelled = [(i, t) for i, t in enumerate(self.templates)]
        n = mtm_port.INF if n_object < 0 else int(n_object)
        return self._raw(mtm_port.match_templates(labelled, self.image, method, n, score_threshold, max_overlap, workers=1))

    def match_templates_async(self, method, n_object, score_threshold, max_overlap, slot):
        self.calls.append("match_templates_async")
        assert slot not in self.slots, "slot %d not collected yet" % slot
        self.slots[slot] = self.match_templates(method, n_object, score_threshold, max_overlap)

    def match_templates_collect(self, slot):
        self.calls.append("match_templates_collect")
        raw = self.slots.pop(slot)
        return None if len(raw) > 1024 else raw

    def match_templates_sharded(self, comm, tmpl_base, n_local, method, n_object, score_threshold, max_overlap):
        """mtm_match_templates_sharded for a one-rank communicator: the slice is the whole list."""
        self.calls.append("match_templates_sharded")
        assert comm.world == 1 and tmpl_base == 0 and n_local == len(self.templates)
        return self.match_templates(method, n_object, score_threshold, max_overlap)

    def nms(self, hits, score_threshold, sort_ascending, n_object, max_overlap):
        self.calls.append("nms")
        listed = [(k, (int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"])), r["score"]) for k, r in enumerate(hits)]
        n = mtm_port.INF if n_object < 0 else int(n_object)
        if n == 1:
            pick = min if sort_ascending else max
            return np.asarray([pick(listed, key=lambda h: h[2])[0]], np.int32)
        import cv2
        boxes = [h[1] for h in listed]
        scores = [float(h[2]) for h in listed]
        keep = cv2.dnn.NMSBoxes(boxes, scores, float(score_threshold), float(max_overlap))
        keep = [int(k) for k in keep]
        if n != mtm_port.INF:
            keep = keep[:n]
        return np.asarray(keep, np.int32)


class MockComm:
    """TEST DOUBLE of ``mtm_b200._native.Comm`` for a single rank (world 1): mtm_gather_results reduces to handing the
    local slots over in the gather layout."""

    def __init__(self, world=1, rank=0, device=0):
        self.world, self.rank, self.device = world, rank, device

    @classmethod
    def init_rank(cls, device, world, rank, unique_id=None):
        assert world == 1
        return cls(1, 0, device)

    def close(self):
        pass

    def gather_results(self, entries, images_per_rank, hits_per_image):
        hits = np.zeros((images_per_rank, hits_per_image), HIT_DTYPE)
        counts = np.full(images_per_rank, -1, np.int32)
        for i, (ctx, slot) in enumerate(entries):
            raw = ctx.slots.pop(slot)
            if len(raw) > hits_per_image:
                counts[i] = -2
            else:
                counts[i] = len(raw)
                hits[i, :len(raw)] = raw
        return hits, counts
