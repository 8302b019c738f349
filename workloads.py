"""Seeded synthetic inputs for the BASELINE.json configs: the workload definition of bench.py / bench_sharded.py / bench_xform.py and
of the tests (which reach it as ``oracle.synth``, a re-export).  It is an input generator, not a checker: nothing of the oracle's
restated algorithms lives here, so the timed arm of the benches imports no checker code.

Generator of SURVEY.md section 8(d): a smooth background with planted, noisy
copies of every template, so that every window has sigma >~ 8 grey levels (cv2's
fp32 numerator noise stays < ~3e-5) and every template has a handful of true
hits scoring 0.5-0.99.  Pure numpy/scipy; no product code depends on it.
"""
import numpy as np
from scipy.ndimage import gaussian_filter


def _rescale(a, lo, hi):
    a = a - a.min()
    return lo + a * ((hi - lo) / max(float(a.max()), 1e-12))


def make_template(rng, h, w):
    return np.clip(np.rint(_rescale(gaussian_filter(rng.random((h, w)), 2.0), 0, 255)), 0, 255).astype(np.uint8)


def make_scene(H, W, templates, n_plant, seed, noise_sigma=8.0):
    """uint8 HxW image containing ``n_plant`` noisy copies of each template."""
    rng = np.random.default_rng(seed + 1000003)
    img = _rescale(gaussian_filter(rng.random((H, W)), 8.0), 40, 160)
    truth = []
    for k, t in enumerate(templates):
        h, w = t.shape[:2]
        for _ in range(n_plant):
            y = int(rng.integers(0, H - h + 1))
            x = int(rng.integers(0, W - w + 1))
            img[y:y + h, x:x + w] = t
            truth.append((k, x, y))
    img = img + rng.normal(0.0, noise_sigma, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8), truth


def config(name, seed=0, image_index=0):
    """Returns ``(image, [(label, template), ...], params)`` for C1..C5 shapes.

    C1 is Fish-derived and lives in tests/golden; the synthetic stand-in here is
    only for shape-compatible smoke runs.
    """
    rng = np.random.default_rng(seed)
    if name == "C1":
        H = W = 256
        temps = [make_template(rng, 64, 64)]
        n_plant, params = 1, dict(N_object=1, score_threshold=0.5, maxOverlap=0.25)
    elif name == "C2":
        H, W = 1080, 1920
        bases = [make_template(rng, 64, 64) for _ in range(2)]
        temps = [np.rot90(b, k) for b in bases for k in range(4)]
        n_plant, params = 4, dict(N_object=float("inf"), score_threshold=0.5, maxOverlap=0.25)
    elif name == "C3":
        H = W = 4096
        temps = [make_template(rng, 256, 256)]
        n_plant, params = 3, dict(N_object=float("inf"), score_threshold=0.5, maxOverlap=0.25)
    elif name == "C4":
        H = W = 2048
        temps = [make_template(rng, 48, 48) for _ in range(32)]
        n_plant, params = 4, dict(N_object=float("inf"), score_threshold=0.5, maxOverlap=0.25)
    elif name == "C5":
        H, W = 2160, 3840
        sides = np.linspace(32, 128, 64).round().astype(int)
        temps = [make_template(rng, int(s), int(s)) for s in sides]
        n_plant, params = 2, dict(N_object=50, score_threshold=0.5, maxOverlap=0.25)
    else:
        raise KeyError(name)
    image, _ = make_scene(H, W, temps, n_plant, seed + image_index)
    labelled = [("t%02d" % i, np.ascontiguousarray(t)) for i, t in enumerate(temps)]
    return image, labelled, params


def macs(image_shape, templates):
    """Direct-correlation MACs: sum_t C*h*w*(H-h+1)*(W-w+1)."""
    H, W = image_shape[:2]
    C = image_shape[2] if len(image_shape) == 3 else 1
    total = 0
    for t in templates:
        arr = t[1] if isinstance(t, tuple) else t
        h, w = arr.shape[:2]
        total += C * h * w * (H - h + 1) * (W - w + 1)
    return total
