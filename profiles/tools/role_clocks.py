"""Role clocks of the persistent numerator kernel (MTM_B200_PROF=1; needs MTM_B200_NO_HITS_ONLY=1: the profiled instantiation
is the map-writing one).  Prints the library's "[mtm prof]" lines of ONE synchronous call per workload (after a warm-up call).

    MTM_B200_PROF=1 MTM_B200_NO_HITS_ONLY=1 [MTM_B200_PDBG=8] python profiles/tools/role_clocks.py C2 C4 C5
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import MTM  # noqa: E402
import workloads  # noqa: E402

import numpy as np  # noqa: E402

for name in sys.argv[1:]:
    if name.startswith("T:"):                                     # T:h:w:count:H:W -- `count` templates of one size on an H x W scene
        h, w, n, H, W = (int(v) for v in name.split(":")[1:])
        rng = np.random.default_rng(0)
        temps = [workloads.make_template(rng, h, w) for _ in range(n)]
        image, _ = workloads.make_scene(H, W, temps, 2, 0)
        labelled = [("t%02d" % i, t) for i, t in enumerate(temps)]
        params = dict(N_object=float("inf"), score_threshold=0.5, maxOverlap=0.25)
    else:
        image, labelled, params = workloads.config(name)
    sys.stderr.write("==== %s warm-up\n" % name)
    MTM.matchTemplates(labelled, image, **params)
    sys.stderr.write("==== %s measured call\n" % name)
    hits = MTM.matchTemplates(labelled, image, **params)
    sys.stderr.write("==== %s: %d hits\n" % (name, len(hits)))
