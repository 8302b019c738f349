"""CPU restatement of the tutorials' augmentation / downscaling user code (TEST INFRASTRUCTURE).

The reference has no function for these steps; its tutorials do them in user code before calling
``matchTemplates``:

* ``tutorials/Tutorial2-Template_Augmentation.ipynb`` cell 15 -- ``np.rot90(temp0, k=i+1)`` (and the remark about
  ``np.fliplr`` / ``np.flipud``);
* ``tutorials/Tutorial3-SpeedingUp.ipynb`` cells 17-25 -- ``cv2.resize(image, smallDim, interpolation=cv2.INTER_AREA)``,
  search the small pair, multiply the boxes by the scale; cell 14 -- ``searchBox``.

``area_downscale`` restates OpenCV's ``resizeAreaFast`` rounding for integer factors (third-party,
``opencv-python-headless``; pinned against the live ``cv2.resize`` in tests/test_oracle.py); everything else runs on
the live cv2 through ``oracle/mtm_port.py``.  ``match_templates_pyramid`` is the specification the product's
``matchTemplatesPyramid`` is tested against; its known answer owned by the reference is the notebook output of
Tutorial3 cell 10 (``[('head', (528, 842, 196, 184), 1.0)]``), which the refined search must reproduce.
Never imported by product code.
"""
import cv2
import numpy as np

from . import mtm_port

INF = float("inf")

HOST_TRANSFORMS = {
    "identity": lambda a: a,
    "rot90": lambda a: np.rot90(a, 1),
    "rot180": lambda a: np.rot90(a, 2),
    "rot270": lambda a: np.rot90(a, 3),
    "fliplr": np.fliplr,
    "flipud": np.flipud,
    "transpose": lambda a: a.swapaxes(0, 1),
    "antitranspose": lambda a: np.rot90(a, 2).swapaxes(0, 1),
}


def area_downscale(arr, f):
    """``cv2.resize(arr[:H//f*f, :W//f*f], (W//f, H//f), interpolation=cv2.INTER_AREA)``, restated.

    Integer pixels: f == 2 -> ``(sum + 2) >> 2``; f >= 3 -> round-half-even of ``float32(sum) * float32(1/f^2)``
    (OpenCV's ``saturate_cast<T>(sum * scale)`` with a float scale).  float32 pixels: mean in float64.
    """
    f = int(f)
    H, W = arr.shape[:2]
    h, w = H // f, W // f
    a = arr[:h * f, :w * f]
    if f == 1:
        return np.ascontiguousarray(a)
    blocks = a.reshape((h, f, w, f) + arr.shape[2:])
    if arr.dtype == np.float32:
        return (blocks.astype(np.float64).sum(axis=(1, 3)) / float(f * f)).astype(np.float32)
    s = blocks.astype(np.int64).sum(axis=(1, 3))
    if f == 2:
        return ((s + 2) >> 2).astype(arr.dtype)
    return np.rint(s.astype(np.float32) * np.float32(1.0 / (f * f))).astype(arr.dtype)


def cv_area_downscale(arr, f):
    """The live third-party call the tutorial makes (sizes cropped to multiples of f first)."""
    H, W = arr.shape[:2]
    h, w = H // f, W // f
    if f == 1:
        return np.ascontiguousarray(arr)
    out = cv2.resize(np.ascontiguousarray(arr[:h * f, :w * f]), (w, h), interpolation=cv2.INTER_AREA)
    return out.reshape((h, w) + arr.shape[2:])


def expand_templates(templates, transforms):
    """Tutorial2 cell 15 generalised: base-major list of (label, transformed array)."""
    out = []
    for entry in templates:
        for t in transforms:
            label = entry[0] if t == "identity" else "%s_%s" % (entry[0], t)
            out.append((label, np.ascontiguousarray(HOST_TRANSFORMS[t](entry[1]))))
    return out


def match_templates_augmented(templates, image, transforms, **kw):
    return mtm_port.match_templates(expand_templates(templates, transforms), image, **kw)


def reference_augmented(ref, templates, image, transforms, **kw):
    """Tutorial2 cell 15 + cell 17 with the unmodified reference: numpy builds the list, ``MTM.matchTemplates`` searches it."""
    return ref.matchTemplates(expand_templates(templates, transforms), image, **kw)


class _PortImpl:
    """The two reference entry points the specification is written in, served by the CPU port."""

    @staticmethod
    def matchTemplates(templates, image, method, N_object, score_threshold, maxOverlap, searchBox=None, workers=None):
        return mtm_port.match_templates(templates, image, method=method, N_object=N_object, score_threshold=score_threshold,
                                        maxOverlap=maxOverlap, searchBox=searchBox, workers=workers)

    NMS = staticmethod(mtm_port.nms)


class ReferenceImpl:
    """Adapter over the UNMODIFIED reference package (``oracle/ref_loader.py``; build container only):
    ``MTM.matchTemplates`` (MTM/__init__.py:247) and ``MTM.NMS`` (MTM/NMS.py:20)."""

    def __init__(self, ref):
        self.ref = ref

    def matchTemplates(self, templates, image, method, N_object, score_threshold, maxOverlap, searchBox=None, workers=None):
        return self.ref.matchTemplates(templates, image, method=method, N_object=N_object, score_threshold=score_threshold,
                                       maxOverlap=maxOverlap, searchBox=searchBox)

    def NMS(self, hits, scoreThreshold, sortAscending, N_object, maxOverlap):
        return self.ref.NMS(hits, scoreThreshold, sortAscending, N_object, maxOverlap)


def match_templates_pyramid(templates, image, downscale=4, method=cv2.TM_CCOEFF_NORMED, N_object=INF,
                            score_threshold=0.5, maxOverlap=0.25, searchBox=None, refine=True,
                            coarse_threshold=None, workers=None, impl=_PortImpl):
    """Coarse search on the INTER_AREA-reduced pair, then full-resolution re-localisation of every coarse hit
    inside a search box of +-downscale pixels (``N_object=1``), then ``NMS``.  See the product docstring.
    ``impl`` supplies ``matchTemplates`` / ``NMS``: the CPU port (default) or the unmodified reference."""
    f = int(downscale)
    if method == 0:
        raise ValueError("The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.")
    x_off = y_off = 0
    if searchBox is not None:
        x_off, y_off, bw, bh = searchBox
        image = image[y_off:y_off + bh, x_off:x_off + bw]
    if coarse_threshold is None:
        coarse_threshold = score_threshold
    H, W = image.shape[:2]
    small_image = cv_area_downscale(image, f)
    small_templates = [(i, cv_area_downscale(entry[1], f)) for i, entry in enumerate(templates)]
    coarse = impl.matchTemplates(small_templates, small_image, method, N_object, coarse_threshold, maxOverlap, workers=workers)
    if not refine:
        return [(templates[t][0], (x * f + x_off, y * f + y_off, w * f, h * f), s) for t, (x, y, w, h), s in coarse]
    refined = []
    for t, (x, y, _w, _h), _s in coarse:
        name, full = templates[t][:2]
        th, tw = full.shape[:2]
        x0, y0 = max(0, x * f - f), max(0, y * f - f)
        x1, y1 = min(W, x * f + tw + f), min(H, y * f + th + f)
        hit = impl.matchTemplates([(name, full)], image, method, 1, score_threshold, maxOverlap,
                                  searchBox=(x0, y0, x1 - x0, y1 - y0), workers=1)
        refined.extend((lbl, (bx + x_off, by + y_off, bw_, bh_), sc) for lbl, (bx, by, bw_, bh_), sc in hit)
    return impl.NMS(refined, score_threshold, method == 1, N_object, maxOverlap)
