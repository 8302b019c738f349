"""Wall time of synchronous MTM.matchTemplates calls on a synthetic shape: T:h:w:count:H:W  (A/B of planning knobs on shapes outside
the BASELINE configs).   MTM_B200_PERSIST=0 python profiles/tools/time_shape.py T:200:200:8:2048:2048"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import MTM  # noqa: E402
import workloads  # noqa: E402

for name in sys.argv[1:]:
    h, w, n, H, W = (int(v) for v in name.split(":")[1:])
    rng = np.random.default_rng(0)
    temps = [workloads.make_template(rng, h, w) for _ in range(n)]
    image, _ = workloads.make_scene(H, W, temps, 2, 0)
    labelled = [("t%02d" % i, t) for i, t in enumerate(temps)]
    params = dict(N_object=float("inf"), score_threshold=0.5, maxOverlap=0.25)
    for _ in range(3):
        hits = MTM.matchTemplates(labelled, image, **params)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        hits = MTM.matchTemplates(labelled, image, **params)
    dt = (time.perf_counter() - t0) / reps
    print("%s persist=%s: %.3f ms per call, %d hits" % (name, os.environ.get("MTM_B200_PERSIST", "1"), dt * 1e3, len(hits)))
