"""-m gpu: parity of the experiment knobs of the library (every alternative kernel of the shipped .so runs on the GPU).

Every knob (environment variable read once per process, see TcEnv in csrc/ncc_tc.cu, ncc_points.cu and box_moments.cu) selects an
alternative kernel or launch plan that must produce the same results as the default; they exist so that the
measurements under profiles/ can be repeated and so that a variant can be validated before it becomes the default.
Each case runs a small parity script in a subprocess with the variable set (about 3 s per case).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r)
import numpy as np
import MTM
%(preamble)s
from oracle import mtm_port, ncc_exact, synth
rng = np.random.default_rng(5)
# grayscale, two sizes (window moments of several sizes), a map large enough for the candidate list
temps = [("a", synth.make_template(rng, 48, 48)), ("b", synth.make_template(rng, 40, 56)), ("c", synth.make_template(rng, 48, 48))]
img, _ = synth.make_scene(420, 520, [t[1] for t in temps], 3, seed=5)
for name, t in temps[:2]:
    got = MTM.computeScoreMap(t, img)
    exact = ncc_exact.match_template_exact(img, t)
    assert np.max(np.abs(got - exact)) <= 1e-4, (name, float(np.max(np.abs(got - exact))))
got = MTM.matchTemplates(temps, img, score_threshold=0.5, maxOverlap=0.25)
want = mtm_port.match_templates(temps, img, score_threshold=0.5, maxOverlap=0.25)
assert [(h[0], h[1]) for h in got] == [(h[0], h[1]) for h in want] and len(want) >= 6, (got, want)
assert max(abs(float(a[2]) - float(b[2])) for a, b in zip(got, want)) <= 1e-4
# a one-size template set (the box-sum moment route under MTM_B200_MOM_BOX), then another method on the same resident image
# (summed-area tables built on demand) and the default method again
same = [temps[0], temps[2]]
for method, thr in ((5, 0.5), (3, 0.97), (5, 0.6)):
    got = MTM.matchTemplates(same, img, method=method, score_threshold=thr, maxOverlap=0.25)
    want = mtm_port.match_templates(same, img, method=method, score_threshold=thr, maxOverlap=0.25)
    assert [(h[0], h[1]) for h in got] == [(h[0], h[1]) for h in want] and len(want) >= 2, (method, got, want)
# RGB (per-channel moments) and a small map of a large template (small-map kernel)
rgb = np.stack([img, img[::-1], 255 - img], axis=2)
t3 = np.ascontiguousarray(rgb[100:140, 200:260])
assert np.max(np.abs(MTM.computeScoreMap(t3, rgb) - ncc_exact.match_template_exact(rgb, t3))) <= 1e-4
big = synth.make_template(rng, 130, 140)
scene, _ = synth.make_scene(140, 160, [big], 1, seed=6)
assert np.max(np.abs(MTM.computeScoreMap(big, scene) - ncc_exact.match_template_exact(scene, big, use_fft=False))) <= 1e-4
print("knob parity ok")
"""

# MTM_B200_MOM_BOX=0: the summed-area moment route (the default until round 2; box sums are the default now)
KNOBS = [{"MTM_B200_MOM_BOX": "0"}, {"MTM_B200_MOM_BOX": "0", "MTM_B200_MOM_ROWS": "1"}, {"MTM_B200_MOM_BOX": "0", "MTM_B200_MOM_CS": "1"},
         {"MTM_B200_NO_POINTS": "1"}, {"MTM_B200_NO_CAND": "1"}, {"MTM_B200_PERSIST": "0"}, {"MTM_B200_EW": "12"}, {"MTM_B200_EW": "8"},
         # the other two output loops of the single-channel box kernel, small / one-row ring stages of the numerator kernel
         {"MTM_B200_BOX_OUT": "0"}, {"MTM_B200_BOX_OUT": "1"}, {"MTM_B200_STAGE_KB": "24"}, {"MTM_B200_DS": "1"},
         {"MTM_B200_PERSIST": "2"}, {"MTM_B200_CTRL_FIRST": "1"}, {}]


@pytest.mark.parametrize("knob", KNOBS, ids=lambda k: ",".join("%s=%s" % kv for kv in k.items()) or "default")
def test_knob_keeps_parity(knob):
    env = dict(os.environ)
    env.update(knob)
    # MTM_B200_EMULATE=1 (tests/conftest.py): the subprocess loads the host build of the library instead (CPU, slow)
    emulated = os.environ.get("MTM_B200_EMULATED_LIB")
    preamble = "from mtm_b200 import _native; _native.LIB_PATH = %r" % emulated if emulated else ""
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "preamble": preamble}], env=env, capture_output=True, text=True,
                       timeout=3600 if emulated else 300)
    assert r.returncode == 0 and "knob parity ok" in r.stdout, (knob, r.stdout[-2000:], r.stderr[-4000:])
