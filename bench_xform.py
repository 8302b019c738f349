"""Timing of the SURVEY §8 f3 front ends on one B200 (not the driver's bench contract -- see bench.py for that).

    python bench_xform.py [--reps 30]

Prints one JSON line per measurement (host wall clock around the public calls, pageable host arrays, results read
back on the host every call; a synchronize on both sides of every timed loop):
  * C2-shaped augmentation: 1080x1920 image, 2 base templates 64x64, 4 rotations -- `matchTemplatesAugmented`
    against `matchTemplates(expandTemplates(...))`, with the template set resident ("warm": content-hash hit) and
    with a different set every call ("cold": upload + statistics + Toeplitz slabs every call);
  * C3-shaped pyramid: 4096x4096 image, one 256x256 template -- `matchTemplatesPyramid(downscale=2, 4, 8)` against
    the full-resolution `matchTemplates`, checking that the hit lists agree.
Nothing under oracle/ is imported; workloads.py generates the inputs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def timed(fn, reps, ctx):
    fn(0)
    fn(1)
    ctx.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        fn(2 + i)
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--once", action="store_true",
                    help="one cold-template call of each front end and exit (the launch list under ncu: profiles/README.md)")
    args = ap.parse_args()
    import MTM
    from mtm_b200 import _native
    import workloads as synth
    ctx = _native.default_context()
    rng = np.random.default_rng(0)

    # ---- augmentation, C2 shape -------------------------------------------------------------
    bases_a = [synth.make_template(rng, 64, 64) for _ in range(2)]
    bases_b = [synth.make_template(rng, 64, 64) for _ in range(2)]
    transforms = ("identity", "rot90", "rot180", "rot270")
    planted = [np.ascontiguousarray(np.rot90(b, k)) for b in bases_a for k in range(4)]
    image, _ = synth.make_scene(1080, 1920, planted, 4, seed=0)
    images = [np.ascontiguousarray(np.roll(image, (13 * s, 29 * s), axis=(0, 1))) for s in range(8)]
    sets = [[("a", bases_a[0]), ("b", bases_a[1])], [("a", bases_b[0]), ("b", bases_b[1])]]
    kw = dict(score_threshold=0.5, maxOverlap=0.25)
    if args.once:
        MTM.matchTemplatesAugmented(sets[0], images[0], transforms, **kw)
        big_t = synth.make_template(rng, 256, 256)
        big, _ = synth.make_scene(4096, 4096, [big_t], 3, seed=0)
        hits = MTM.matchTemplatesPyramid([("big", big_t)], big, downscale=4, **kw)
        print(json.dumps({"bench": "once", "pyramid_hits": len(hits)}), flush=True)
        return
    ref = MTM.matchTemplates(MTM.expandTemplates(sets[0], transforms), images[0], **kw)
    got = MTM.matchTemplatesAugmented(sets[0], images[0], transforms, **kw)
    same = [(h[0], h[1], float(h[2])) for h in ref] == [(h[0], h[1], float(h[2])) for h in got]
    res = {
        "host_expansion_warm_ms": timed(lambda i: MTM.matchTemplates(MTM.expandTemplates(sets[0], transforms), images[i % 8], **kw), args.reps, ctx),
        "device_augmentation_warm_ms": timed(lambda i: MTM.matchTemplatesAugmented(sets[0], images[i % 8], transforms, **kw), args.reps, ctx),
        "host_expansion_cold_ms": timed(lambda i: MTM.matchTemplates(MTM.expandTemplates(sets[i % 2], transforms), images[i % 8], **kw), args.reps, ctx),
        "device_augmentation_cold_ms": timed(lambda i: MTM.matchTemplatesAugmented(sets[i % 2], images[i % 8], transforms, **kw), args.reps, ctx),
    }
    print(json.dumps({"bench": "augmentation", "workload": "C2 shape: 1080x1920 u8, 2 base templates 64x64 x 4 rotations",
                      "identical_hits": same, "hits": len(got), **{k: round(v, 4) for k, v in res.items()}}), flush=True)

    # ---- pyramid, C3 shape ------------------------------------------------------------------
    big_t = synth.make_template(rng, 256, 256)
    big, _ = synth.make_scene(4096, 4096, [big_t], 3, seed=0)
    labelled = [("big", big_t)]
    reps = max(5, args.reps // 3)
    full = MTM.matchTemplates(labelled, big, **kw)
    out = {"bench": "pyramid", "workload": "C3 shape: 4096x4096 u8, 1 template 256x256", "hits": len(full),
           "full_resolution_ms": round(timed(lambda i: MTM.matchTemplates(labelled, big, **kw), reps, ctx), 4)}
    for f in (2, 4, 8):
        pyr = MTM.matchTemplatesPyramid(labelled, big, downscale=f, **kw)
        out["pyramid_f%d_ms" % f] = round(timed(lambda i: MTM.matchTemplatesPyramid(labelled, big, downscale=f, **kw), reps, ctx), 4)
        out["pyramid_f%d_identical_boxes" % f] = [(h[0], h[1]) for h in pyr] == [(h[0], h[1]) for h in full]
        out["pyramid_f%d_coarse_only_ms" % f] = round(timed(
            lambda i: MTM.matchTemplatesPyramid(labelled, big, downscale=f, refine=False, **kw), reps, ctx), 4)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
