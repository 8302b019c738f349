"""On-device template augmentation and coarse-to-fine search -- SURVEY.md §8 f3.

The reference leaves both to user code in its tutorials:

* ``Tutorial2-Template_Augmentation.ipynb`` cell 15 builds the template list with ``np.rot90``
  ("We could also do some flipping with np.fliplr, flipud") before calling ``matchTemplates``;
* ``Tutorial3-SpeedingUp.ipynb`` cells 17-25 shrink image and template with
  ``cv2.resize(..., interpolation=cv2.INTER_AREA)``, search the small pair and scale the boxes up,
  and cell 26 notes that combining this with a search region "starts to be a bit code-heavy".

Here the base pixels cross PCIe once and the rotated / flipped / reduced copies are made by
``transform.cu`` on the device; the searches themselves are the same kernels ``matchTemplates``
uses.  These functions are additive: nothing in the reference's own API changes.
"""
import numpy as np

from . import _native
from .api import (NMS, TM_CCOEFF_NORMED, _INF, _cv_error, _to_hits, _validate_search)

__all__ = ["TRANSFORMS", "expandTemplates", "matchTemplatesAugmented", "matchTemplatesPyramid"]

# name -> mtm_transform code (include/mtm_b200.h)
TRANSFORMS = {
    "identity": _native.XF_IDENTITY, "rot90": _native.XF_ROT90, "rot180": _native.XF_ROT180, "rot270": _native.XF_ROT270,
    "fliplr": _native.XF_FLIPLR, "flipud": _native.XF_FLIPUD, "transpose": _native.XF_TRANSPOSE,
    "antitranspose": _native.XF_ANTITRANSPOSE,
}
_SWAPS = {"rot90", "rot270", "transpose", "antitranspose"}
_HOST = {
    "identity": lambda a: a, "rot90": lambda a: np.rot90(a, 1), "rot180": lambda a: np.rot90(a, 2),
    "rot270": lambda a: np.rot90(a, 3), "fliplr": np.fliplr, "flipud": np.flipud,
    "transpose": lambda a: a.swapaxes(0, 1), "antitranspose": lambda a: np.rot90(a, 2).swapaxes(0, 1),
}


class _ShapeOnly:
    """Stand-in for a template array in the reference's size checks (only ``.shape`` is read)."""

    def __init__(self, shape):
        self.shape = tuple(shape)


def _label(name, transform):
    return name if transform == "identity" else "%s_%s" % (name, transform)


def _check_transforms(transforms):
    transforms = list(transforms)
    if not transforms:
        raise ValueError("transforms must name at least one of %s" % sorted(TRANSFORMS))
    for t in transforms:
        if t not in TRANSFORMS:
            raise ValueError("unknown transform %r (expected one of %s)" % (t, sorted(TRANSFORMS)))
    return transforms


def expandTemplates(listTemplates, transforms=("identity", "rot90", "rot180", "rot270")):
    """The tutorial's host-side loop, for reference: ``[(label, transformed array), ...]``, base-major.

    ``matchTemplatesAugmented(L, image, T, ...)`` returns what ``matchTemplates(expandTemplates(L, T), image, ...)``
    returns.  Labels: the template's own name for ``"identity"``, ``name_<transform>`` otherwise.
    """
    transforms = _check_transforms(transforms)
    return [(_label(entry[0], t), _HOST[t](entry[1])) for entry in listTemplates for t in transforms]


def _device_template_set(listTemplates, image, what):
    """Arrays of a template list the device transforms accept: no masks, one dtype (uint8 or float32)
    shared with the image (the reference's float32 casts are not applied silently here)."""
    arrays = []
    for entry in listTemplates:
        if not isinstance(entry, tuple) or len(entry) < 2:
            raise ValueError("listTemplates should be a list of tuples as ('name','array') or ('name', 'array', 'mask')")
        if len(entry) >= 3 and entry[2] is not None:
            raise NotImplementedError("%s: templates with masks are not transformed on the device" % what)
        arrays.append(entry[1])
    if image.dtype == "float64" or any(a.dtype == "float64" for a in arrays):
        raise ValueError("64-bit images not supported, max 32-bit")
    if image.dtype not in (np.uint8, np.float32) or any(a.dtype != image.dtype for a in arrays):
        raise NotImplementedError("%s: image and templates must all be uint8 or all be float32 (got %s / %s)"
                                  % (what, image.dtype, sorted({str(a.dtype) for a in arrays})))
    for a in arrays:
        if a.ndim != image.ndim or a.shape[2:] != image.shape[2:]:
            raise _cv_error("matchTemplate: image and template must have the same number of dimensions/channels")
    return arrays


def _fused_or_split(ctx, method, N_object, score_threshold, maxOverlap, names, xOffset, yOffset):
    """Search + NMS of the context's current image / template list with MTM.matchTemplates' semantics
    (MTM/__init__.py:289-296)."""
    finite = N_object != _INF
    nms_threshold = (1 - score_threshold) if method == 1 else score_threshold
    if (finite and N_object < 1) or nms_threshold < 0:
        # same rare corners as api.matchTemplates: the two stages run separately
        raw = ctx.find_matches(method, 1 if N_object == 1 else -1, score_threshold)
        return NMS(_to_hits(raw, names, xOffset, yOffset), score_threshold, method == 1, N_object, maxOverlap, context=ctx)
    raw = ctx.match_templates(method, int(N_object) if finite else -1, score_threshold, maxOverlap)
    return _to_hits(raw, names, xOffset, yOffset)


def matchTemplatesAugmented(listTemplates, image, transforms=("identity", "rot90", "rot180", "rot270"),
                            method=TM_CCOEFF_NORMED, N_object=_INF, score_threshold=0.5, maxOverlap=0.25,
                            searchBox=None, *, context=None):
    """``matchTemplates(expandTemplates(listTemplates, transforms), image, ...)`` with the augmentation on the device.

    Only the base templates are uploaded; ``mtm_set_templates_transformed`` writes the rotated / flipped copies
    straight into the device template arena (bit-identical to ``np.rot90`` / ``np.fliplr`` / ``np.flipud``).
    """
    transforms = _check_transforms(transforms)
    if maxOverlap < 0 or maxOverlap > 1:
        raise ValueError("Maximal overlap between bounding box is in range [0-1]")
    expanded_shapes = []
    for entry in listTemplates:
        if isinstance(entry, tuple) and len(entry) >= 2:
            for t in transforms:
                shp = tuple(entry[1].shape)
                expanded_shapes.append((_label(entry[0], t), _ShapeOnly((shp[1], shp[0]) + shp[2:] if t in _SWAPS else shp)))
        else:
            expanded_shapes.append(entry)              # rejected by the reference's check below
    crop, xOffset, yOffset = _validate_search(expanded_shapes, _native.as_image(image), N_object, searchBox)
    if len(listTemplates) == 0:
        return []
    arrays = _device_template_set(listTemplates, crop, "matchTemplatesAugmented")
    names = [lbl for lbl, _ in expanded_shapes]
    ops = [TRANSFORMS[t] for t in transforms]
    ctx = context or _native.default_context()
    with ctx.lock:
        ctx.set_image(crop)
        ctx.set_templates_transformed(arrays, ops, 1)
        if method == 0:
            # the reference searches first and rejects TM_SQDIFF afterwards (MTM/__init__.py:289-292)
            ctx.find_matches(method, 1 if N_object == 1 else -1, score_threshold)
            raise ValueError("The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.")
        return _fused_or_split(ctx, method, N_object, score_threshold, maxOverlap, names, xOffset, yOffset)


def matchTemplatesPyramid(listTemplates, image, downscale=4, method=TM_CCOEFF_NORMED, N_object=_INF,
                          score_threshold=0.5, maxOverlap=0.25, searchBox=None, refine=True,
                          coarse_threshold=None, *, context=None):
    """Coarse-to-fine ``matchTemplates``: Tutorial3-SpeedingUp.ipynb sections II + III combined, on the device.

    1. Image and templates are reduced ``downscale`` times with OpenCV's ``INTER_AREA`` rule (sizes are first cropped
       to multiples of ``downscale``) and searched like ``matchTemplates(small_templates, small_image, method,
       N_object, coarse_threshold, maxOverlap)``; ``coarse_threshold`` defaults to ``score_threshold``.
    2. ``refine=False``: the coarse hits are returned with their boxes multiplied by ``downscale`` (the tutorial's
       cell 25) and their coarse scores.
    3. ``refine=True``: every coarse hit ``(x, y)`` of template ``T`` (``th`` x ``tw``) is re-localised at full
       resolution with ``N_object=1`` inside the search box ``[x*f - f, x*f + tw + f) x [y*f - f, y*f + th + f)``
       clipped to the image (the full-resolution pixels are already resident: ``mtm_set_image_roi``), and the
       refined hits go through ``NMS(hits, score_threshold, method == 1, N_object, maxOverlap)``.

    Returns hits in full-resolution image coordinates, like ``matchTemplates``.
    """
    f = int(downscale)
    if f < 1 or f > _native.MAX_DOWNSCALE:
        raise ValueError("downscale must be an integer in [1, %d]" % _native.MAX_DOWNSCALE)
    if maxOverlap < 0 or maxOverlap > 1:
        raise ValueError("Maximal overlap between bounding box is in range [0-1]")
    if method == 0:
        raise ValueError("The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.")
    if not isinstance(image, np.ndarray) and hasattr(image, "__cuda_array_interface__"):
        raise NotImplementedError("matchTemplatesPyramid takes host images (the reduction starts from the uploaded full-resolution pixels)")
    crop, xOffset, yOffset = _validate_search(listTemplates, image, N_object, searchBox)
    if len(listTemplates) == 0:
        return []
    arrays = _device_template_set(listTemplates, crop, "matchTemplatesPyramid")
    names = [entry[0] for entry in listTemplates]
    for name, a in zip(names, arrays):
        if a.shape[0] // f < 1 or a.shape[1] // f < 1:
            raise ValueError("Template '%s' (%d x %d) vanishes at downscale %d" % (name, a.shape[0], a.shape[1], f))
    if coarse_threshold is None:
        coarse_threshold = score_threshold
    H, W = crop.shape[:2]
    ctx = context or _native.default_context()
    with ctx.lock:
        ctx.set_image_scaled(crop, f)
        ctx.set_templates_transformed(arrays, [_native.XF_IDENTITY], f)
        coarse = _fused_or_split(ctx, method, N_object, coarse_threshold, maxOverlap, list(range(len(names))), 0, 0)
        if not refine:
            return [(names[t], (x * f + xOffset, y * f + yOffset, w * f, h * f), score) for t, (x, y, w, h), score in coarse]
        # Re-localisation: one small search per coarse hit, pipelined through the asynchronous entry points (no host
        # synchronisation per hit); hits are visited template by template so that the template upload is reused.
        refined = [None] * len(coarse)
        depth = _native.MAX_INFLIGHT
        pending = {}                                   # slot -> (position in `coarse`, template, x0, y0)

        def region(k):
            t, (x, y, _w, _h), _score = coarse[k]
            th, tw = arrays[t].shape[:2]
            x0, y0 = max(0, x * f - f), max(0, y * f - f)
            return t, x0, y0, min(W, x * f + tw + f) - x0, min(H, y * f + th + f) - y0

        def collect(slot):
            k, t, x0, y0 = pending.pop(slot)
            raw = ctx.match_templates_collect(slot)
            refined[k] = None if raw is None else _to_hits(raw, [names[t]], x0 + xOffset, y0 + yOffset)

        resident = None                                # template whose full-resolution pixels are the context's current set
        try:
            for n, k in enumerate(sorted(range(len(coarse)), key=lambda q: coarse[q][0])):
                slot = n % depth
                if slot in pending:
                    collect(slot)
                t, x0, y0, bw, bh = region(k)
                ctx.set_image_roi(x0, y0, bw, bh)
                if t != resident:
                    ctx.set_templates([arrays[t]])
                    resident = t
                ctx.match_templates_async(method, 1, score_threshold, maxOverlap, slot)
                pending[slot] = (k, t, x0, y0)
        finally:
            for slot in sorted(pending, key=lambda q: pending[q][0]):
                collect(slot)
        for k, hits in enumerate(refined):             # a submission the fused fast path declined: synchronous call
            if hits is None:
                t, x0, y0, bw, bh = region(k)
                ctx.set_image_roi(x0, y0, bw, bh)
                ctx.set_templates([arrays[t]])
                refined[k] = _to_hits(ctx.match_templates(method, 1, score_threshold, maxOverlap), [names[t]],
                                      x0 + xOffset, y0 + yOffset)
    return NMS([hit for hits in refined for hit in hits], score_threshold, method == 1, N_object, maxOverlap, context=ctx)
