"""Host-side mirror of the MTM 2.0.1 API for the sliding-window NCC hot path.

Same names, argument order, defaults, return types and exception types as the
reference (``MTM/__init__.py:56,95,247`` and ``MTM/NMS.py:20``); every array
operation is executed by libmtm_b200.so on a B200 (see include/mtm_b200.h).
There is no CPU fallback: input combinations whose GPU kernel is not written yet
raise ``NotImplementedError``.
"""
import contextlib
import warnings

import numpy as np

from . import _native

__all__ = ["computeScoreMap", "findMatches", "matchTemplates", "matchTemplatesBatch", "NMS"]

TM_SQDIFF, TM_SQDIFF_NORMED, TM_CCORR, TM_CCORR_NORMED, TM_CCOEFF, TM_CCOEFF_NORMED = range(6)
_INF = float("inf")
_U8 = np.dtype(np.uint8)


def _cv_error(msg):
    try:
        import cv2
        return cv2.error(msg)
    except Exception:                      # cv2 is optional for the compute path
        return ValueError(msg)


def _dtype_policy(template, image, mask):
    """MTM/__init__.py:67-74."""
    if template.dtype == "float64" or image.dtype == "float64":
        raise ValueError("64-bit images not supported, max 32-bit")
    if not (template.dtype == "uint8" and image.dtype == "uint8"):
        template = np.float32(template)
        if not isinstance(image, _native.DeviceArray):            # device images: checked by _device_image_dtypes
            image = np.float32(image)
        if mask is not None:
            mask = np.float32(mask)
    return template, image, mask


def _device_image_dtypes(image, templates, usable_mask):
    """An image that is already in HBM (``_native.DeviceArray``) is never cast on the host: the combinations the
    reference would cast (MTM/__init__.py:71-74) or mask are only accepted when no cast of the IMAGE is needed."""
    if not isinstance(image, _native.DeviceArray):
        return
    if image.dtype == "float64":
        raise ValueError("64-bit images not supported, max 32-bit")
    all_u8 = image.dtype == np.uint8 and all(t.dtype == np.uint8 for t in templates)
    all_u16 = image.dtype == np.uint16 and image.ndim == 2 and not usable_mask and \
        all(t.dtype == np.uint16 and t.ndim == 2 for t in templates)
    if not (all_u8 or all_u16 or image.dtype == np.float32):
        raise NotImplementedError("device-resident %s image with %s templates: the reference's float32 cast of the image is not done "
                                  "on the host; pass a float32 (or matching uint8 / 2-D uint16) device image"
                                  % (image.dtype, sorted({str(t.dtype) for t in templates})))


def _mask_policy(template, mask, method):
    """MTM/__init__.py:76-88."""
    if mask is not None:
        if method not in (0, 3):
            mask = None
            warnings.warn("Template matching method not compatible with use of mask (only 0/TM_SQDIFF or 3/TM_CCORR_NORMED).\n-> Ignoring mask.")
        elif not (mask.shape == template.shape and mask.dtype == template.dtype):
            mask = None
            warnings.warn("Mask does not have the same dimension or bit depth than the template.\n-> Ignoring mask.")
    return mask


def _require_gpu_support(image, templates, mask=None):
    if image.dtype not in (np.uint8, np.float32) or any(t.dtype != image.dtype for t in templates):
        raise NotImplementedError("unsupported dtype combination %s / %s" % (image.dtype, [t.dtype for t in templates]))


def computeScoreMap(template, image, method=TM_CCOEFF_NORMED, mask=None, *, context=None):
    """Score map of ``template`` over ``image`` -- ``MTM.computeScoreMap`` (MTM/__init__.py:56-92).

    Note the argument order (template, image), the reverse of ``cv2.matchTemplate``.
    Returns a float32 array of shape (H-h+1, W-w+1).
    """
    image = _native.as_image(image)
    _device_image_dtypes(image, [template], mask is not None and method in (0, 3))
    template16, image16 = template, image
    template, image, mask = _dtype_policy(template, image, mask)
    mask = _mask_policy(template, mask, method)
    if template.ndim != image.ndim or any(t > i for t, i in zip(template.shape, image.shape)) \
            or template.shape[2:] != image.shape[2:]:
        raise _cv_error("matchTemplate: template must not be larger than the image and must have the same channels")
    if mask is None and _all_uint16(image16, [template16]):
        image, template = image16, template16              # MTM_U16: see _prepare
    else:
        _require_gpu_support(image, [template], mask)
    ctx = context or _native.default_context()
    with ctx.lock:
        ctx.set_image(image)
        _upload_templates(ctx, [template], [mask])
        return ctx.score_map(0, method, (image.shape[0] - template.shape[0] + 1, image.shape[1] - template.shape[1] + 1))


def _upload_templates(ctx, arrays, masks):
    """Plain or masked template upload.  When at least one template carries a (valid) mask every
    template goes through the masked kernels; templates without one get an all-ones mask, for which
    OpenCV's masked and plain formulas of TM_SQDIFF / TM_CCORR_NORMED coincide."""
    if all(m is None for m in masks):
        ctx.set_templates(arrays)
        return
    one = 255 if arrays[0].dtype == np.uint8 else 1
    full = [m if m is not None else np.full(a.shape, one, a.dtype) for a, m in zip(arrays, masks)]
    ctx.set_templates_masked(arrays, full)


def _validate_search(listTemplates, image, N_object, searchBox):
    """Argument checks of MTM.findMatches (MTM/__init__.py:129-167).  Returns the
    (possibly cropped) image view and the offsets."""
    if N_object != _INF and not isinstance(N_object, int):
        raise TypeError("N_object must be an integer")
    if image.shape[0] == 0:
        raise ValueError("Image has a height of 0.")
    if image.shape[1] == 0:
        raise ValueError("Image has a width of 0.")
    xOffset = yOffset = 0
    if searchBox is not None:
        xOffset, yOffset, searchWidth, searchHeight = searchBox
        image = image[yOffset:yOffset + searchHeight, xOffset:xOffset + searchWidth]
    ishape = image.shape
    for index, tempTuple in enumerate(listTemplates):
        if not isinstance(tempTuple, tuple) or len(tempTuple) < 2:
            raise ValueError("listTemplates should be a list of tuples as ('name','array') or ('name', 'array', 'mask')")
        tshape = tempTuple[1].shape
        if tshape[0] == 0:
            raise ValueError(f"Template '{tempTuple[0]}' has a height of 0.")
        if tshape[1] == 0:
            raise ValueError(f"Template '{tempTuple[0]}' has a width of 0.")
        # `all(t <= i for t, i in zip(template.shape, image.shape))` of the reference, unrolled for the first two axes
        if tshape[0] > ishape[0] or tshape[1] > ishape[1] or \
                (len(tshape) > 2 and len(ishape) > 2 and not all(t <= i for t, i in zip(tshape[2:], ishape[2:]))):
            fitIn = "searchBox" if (searchBox is not None) else "image"
            raise ValueError("Template '{}' at index {} in the list of templates is larger than {}.".format(tempTuple[0], index, fitIn))
    return image, xOffset, yOffset


def _prepare(listTemplates, image, method):
    """Per-template mask / dtype policy of _multi_compute + computeScoreMap
    (MTM/__init__.py:207-222, 67-88) applied to the whole list."""
    if image.dtype == _U8:
        # the common case in one pass: uint8 image, uint8 templates, no mask entries -> nothing is cast or dropped
        names, arrays = [], []
        ndim, chans = image.ndim, image.shape[2:]
        for tempTuple in listTemplates:
            template = tempTuple[1]
            if len(tempTuple) != 2 or template.dtype != _U8:
                break
            if template.ndim != ndim or template.shape[2:] != chans:
                raise _cv_error("matchTemplate: image and template must have the same number of dimensions/channels")
            names.append(tempTuple[0])
            arrays.append(template)
        else:
            return names, arrays, image, [None] * len(arrays)
    _device_image_dtypes(image, [t[1] for t in listTemplates],
                         method in (0, 3) and any(len(t) >= 3 and t[2] is not None for t in listTemplates))
    names, arrays, masks = [], [], []
    img = image
    img_f32 = None                     # the image is cast at most once (the reference casts it once per template)
    # 16-bit grayscale without usable masks: hand the integers over as they are (MTM_U16).  The library keeps the reference's
    # float32 semantics for the statistics and computes the numerator exactly from the byte planes on the tensor cores.
    route16 = _all_uint16(image, [t[1] for t in listTemplates]) and \
        not any(len(t) >= 3 and t[2] is not None and method in (0, 3) for t in listTemplates)
    for tempTuple in listTemplates:
        name, template = tempTuple[:2]
        if route16:
            if len(tempTuple) >= 3 and method not in (0, 3):
                warnings.warn("Template matching method not supporting the use of Mask. Use 0/TM_SQDIFF or 3/TM_CCORR_NORMED.")
            names.append(name)
            arrays.append(template)
            masks.append(None)
            continue
        mask = None
        if len(tempTuple) >= 3:
            if method in (0, 3):
                mask = tempTuple[2]
            else:
                warnings.warn("Template matching method not supporting the use of Mask. Use 0/TM_SQDIFF or 3/TM_CCORR_NORMED.")
        if template.dtype == "float64" or image.dtype == "float64":
            raise ValueError("64-bit images not supported, max 32-bit")
        if template.dtype == "uint8" and image.dtype == "uint8":
            img_t = image
        else:                          # MTM/__init__.py:71-74
            template = np.float32(template)
            if img_f32 is None:
                img_f32 = image if image.dtype == np.float32 else np.float32(image)
            img_t = img_f32
            if mask is not None:
                mask = np.float32(mask)
        mask = _mask_policy(template, mask, method)
        if template.ndim != img_t.ndim or template.shape[2:] != img_t.shape[2:]:
            raise _cv_error("matchTemplate: image and template must have the same number of dimensions/channels")
        _require_gpu_support(img_t, [template], mask)
        img = img_t
        names.append(name)
        arrays.append(template)
        masks.append(mask)
    return names, arrays, img, masks


def _all_uint16(image, templates):
    return (isinstance(image, (np.ndarray, _native.DeviceArray)) and image.dtype == np.uint16 and image.ndim == 2 and
            all(isinstance(t, np.ndarray) and t.dtype == np.uint16 and t.ndim == 2 for t in templates))


def _native_n_object(N_object):
    if N_object == _INF:
        return -1
    return 1 if N_object == 1 else -1      # only "exactly one" changes the peak search (MTM/__init__.py:225)


def _to_hits(raw, names, xOffset, yOffset):
    """Raw device records -> the reference's hit tuples (label, (x, y, w, h), np.float32 score), MTM/__init__.py:238-241."""
    if len(raw) == 0:
        return []
    # column-wise conversion: per-record field access on a structured array costs microseconds per hit
    return [(names[t], (x + xOffset, y + yOffset, w, h), s)
            for t, x, y, w, h, s in zip(raw["tmpl"].tolist(), raw["x"].tolist(), raw["y"].tolist(), raw["w"].tolist(),
                                        raw["h"].tolist(), list(raw["score"]))]


def findMatches(listTemplates, image, method=TM_CCOEFF_NORMED, N_object=_INF, score_threshold=0.5,
                searchBox=None, *, context=None):
    """All candidate hits before NMS -- ``MTM.findMatches`` (MTM/__init__.py:95-177).

    Hits are returned in template-list order (the reference appends them in
    thread-completion order, so any order it can produce is a permutation of this
    one), each template's hits in the order its peak finder yields them.
    """
    image, xOffset, yOffset = _validate_search(listTemplates, _native.as_image(image), N_object, searchBox)
    if len(listTemplates) == 0:
        return []
    names, arrays, img, masks = _prepare(listTemplates, image, method)
    ctx = context or _native.default_context()
    with ctx.lock:
        ctx.set_image(img)
        _upload_templates(ctx, arrays, masks)
        raw = ctx.find_matches(method, _native_n_object(N_object), score_threshold)
    return _to_hits(raw, names, xOffset, yOffset)


def matchTemplates(listTemplates, image, method=TM_CCOEFF_NORMED, N_object=_INF, score_threshold=0.5,
                   maxOverlap=0.25, searchBox=None, *, context=None):
    """Best non-overlapping hits -- ``MTM.matchTemplates`` (MTM/__init__.py:247-296).

    Search, peak extraction, sorting and NMS all run on the device; one small
    device->host copy brings back the final hit list.
    """
    if maxOverlap < 0 or maxOverlap > 1:
        raise ValueError("Maximal overlap between bounding box is in range [0-1]")
    if method == 0:
        # the reference searches first and rejects TM_SQDIFF afterwards (MTM/__init__.py:289-292)
        findMatches(listTemplates, image, method, N_object, score_threshold, searchBox, context=context)
        raise ValueError("The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.")
    image = _native.as_image(image)
    crop, xOffset, yOffset = _validate_search(listTemplates, image, N_object, searchBox)
    if len(listTemplates) == 0:
        return []
    names, arrays, img, masks = _prepare(listTemplates, crop, method)
    finite = N_object != _INF
    nms_threshold = (1 - score_threshold) if method == 1 else score_threshold
    if (finite and N_object < 1) or nms_threshold < 0:
        # rare corners whose behaviour depends on the pre-NMS list length (nHits<=1 short-cut before
        # the [:N_object] cut, NMSBoxes' score_threshold>=0 assertion): run the two stages separately
        hits = findMatches(listTemplates, image, method, N_object, score_threshold, searchBox, context=context)
        return NMS(hits, score_threshold, method == 1, N_object, maxOverlap, context=context)
    n_dev = int(N_object) if finite else -1
    ctx = context or _native.default_context()
    with ctx.lock:
        ctx.set_image(img)
        _upload_templates(ctx, arrays, masks)
        raw = ctx.match_templates(method, n_dev, score_threshold, maxOverlap)
    return _to_hits(raw, names, xOffset, yOffset)


def matchTemplatesBatch(listTemplates, images, method=TM_CCOEFF_NORMED, N_object=_INF, score_threshold=0.5,
                        maxOverlap=0.25, searchBox=None, *, context=None, streams=2):
    """``[matchTemplates(listTemplates, im, ...) for im in images]`` as one pipelined submission.

    The reference has no batch entry point (users loop over images, e.g. the 16-image batch of
    BASELINE.json configs[4]); here the templates are uploaded once per stream and the images go through the
    GPU back to back (``mtm_match_templates_async`` / ``_collect``): no per-image host synchronisation.
    ``streams`` contexts of the device (the given / default one plus helpers, one CUDA stream each) take the
    images in turn, so the upload of one image overlaps the search of the previous one.
    Results are identical to the per-image calls.
    """
    images = list(images)
    if maxOverlap < 0 or maxOverlap > 1:
        raise ValueError("Maximal overlap between bounding box is in range [0-1]")
    finite = N_object != _INF
    nms_threshold = (1 - score_threshold) if method == 1 else score_threshold
    if method == 0 or len(listTemplates) == 0 or (finite and N_object < 1) or nms_threshold < 0:
        return [matchTemplates(listTemplates, im, method, N_object, score_threshold, maxOverlap, searchBox,
                               context=context) for im in images]
    ctx = context or _native.default_context()
    n_streams = max(1, min(int(streams), len(images)))
    ctxs = [ctx] + _native.helper_contexts(ctx.device, n_streams - 1)
    n_dev = int(N_object) if finite else -1
    depth = _native.MAX_INFLIGHT
    results = [None] * len(images)
    pending = {}                                   # (stream, slot) -> (image index, names, offsets, keep-alive)
    uploaded = [None] * n_streams                  # per stream: (signature, arrays kept alive) of its resident template set

    def collect(key):
        i, nm, xo, yo, _keep = pending.pop(key)
        raw = ctxs[key[0]].match_templates_collect(key[1])
        results[i] = None if raw is None else _to_hits(raw, nm, xo, yo)

    with contextlib.ExitStack() as stack:
        for c in ctxs:
            stack.enter_context(c.lock)
        try:
            for i, image in enumerate(images):
                crop, xOffset, yOffset = _validate_search(listTemplates, _native.as_image(image), N_object, searchBox)
                nm, arrays, img, masks = _prepare(listTemplates, crop, method)
                k = i % n_streams
                key = (k, (i // n_streams) % depth)
                if key in pending:
                    collect(key)
                c = ctxs[k]
                c.set_image(img)
                # the routed template arrays are the caller's own objects unless the dtype policy had to cast them
                sig = (img.dtype, img.ndim, tuple(map(id, arrays)), tuple(map(id, masks)))
                if uploaded[k] is None or uploaded[k][0] != sig:
                    _upload_templates(c, arrays, masks)
                    uploaded[k] = (sig, arrays, masks)
                c.match_templates_async(method, n_dev, score_threshold, maxOverlap, key[1])
                pending[key] = (i, nm, xOffset, yOffset, img)
        finally:
            # in submission order; also drains the slots when an image of the batch was rejected
            for key in sorted(pending, key=lambda q: pending[q][0]):
                collect(key)
    for i, r in enumerate(results):                # images that exceeded the fused fast path
        if r is None:
            results[i] = matchTemplates(listTemplates, images[i], method, N_object, score_threshold, maxOverlap,
                                        searchBox, context=context)
    return results


def NMS(listHit, scoreThreshold=0.5, sortAscending=False, N_object=_INF, maxOverlap=0.5, *, context=None):
    """Non-maxima suppression of a hit list -- ``MTM.NMS.NMS`` (MTM/NMS.py:20-84)."""
    nHits = len(listHit)
    if nHits <= 1:
        return listHit[:]
    boxes = [hit[1] for hit in listHit]
    scores = [hit[2] for hit in listHit]
    raw = np.zeros(nHits, _native.HIT_DTYPE)
    raw["tmpl"] = np.arange(nHits)
    b = np.asarray(boxes, dtype=np.int64).reshape(nHits, 4)
    raw["x"], raw["y"], raw["w"], raw["h"] = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    ctx = context or _native.default_context()
    if N_object == 1:
        raw["score"] = np.asarray(scores, dtype=np.float32)
        with ctx.lock:
            keep = ctx.nms(raw, 0.0, sortAscending, 1, maxOverlap)
        return [listHit[int(i)] for i in keep]
    if sortAscending:
        # list plumbing of MTM/NMS.py:73-75 (keeps Python's own float semantics for `1 - score`)
        scores = [1 - s for s in scores]
        scoreThreshold = 1 - scoreThreshold
    if scoreThreshold < 0 or maxOverlap < 0:
        raise _cv_error("NMSBoxes: score_threshold >= 0 and nms_threshold >= 0 are required")
    raw["score"] = np.asarray(scores, dtype=np.float32)
    finite = N_object != _INF
    n_dev = int(N_object) if (finite and N_object > 1) else -1
    with ctx.lock:
        keep = ctx.nms(raw, scoreThreshold, False, n_dev, maxOverlap)
    keep = [int(i) for i in keep]
    if finite:
        keep = keep[:N_object]
    return [listHit[i] for i in keep]
