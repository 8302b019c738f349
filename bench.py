#!/usr/bin/env python
"""bench.py -- headline benchmark of the MTM hot path on B200 (BASELINE.json metric: template-matches/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C5x16|C2|C3|C4|C5]

HEADLINE (default, every N so that the scaling curve is ONE metric): BASELINE.json configs[4] as a STRONG-scaling
batch -- 16 images 3840x2160 x 64 templates (32..128 px), maxOverlap 0.25, N_object 50.  One "step" = the whole batch:
the images are cut into blocks over the ranks (one process per GPU), every rank searches its block with the whole template
list (score maps -> peaks -> NMS on the device, NMS local to an image, MTM/__init__.py:296), then ONE all-gather of the final
hit blocks (ncclAllGather inside libmtm_b200.so, mtm_gather_results) hands every image's hit list to every rank.  The
exchange and the host read-back of all 16 lists are INSIDE the timed region; the result is checked against the
single-GPU per-image calls (`identical_to_single_gpu`).

* ``value``  : 16*64*K matches / time, images resident in HBM (per-rank device pool larger than L2, rotated every
               step), timed with CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks.
* ``e2e``    : the same batch through the public API ``mtm_b200.sharded.matchTemplatesBatchSharded`` with HOST images
               as a drop-in caller passes them (pageable numpy; every rank uploads its own block each step; ``e2e.pinned``
               = the same from page-locked arrays).
* ``per_config``: C2 / C3 / C4 (BASELINE configs[1..3]) each with value / e2e / roofline / cpu_baseline.  N > 1: C2 and C3
               run as independent replicas (weak), C4 is TEMPLATE-sharded (mtm_match_templates_sharded: one all-gather of
               hit blocks + replicated global NMS, strong).
* ``roofline``: the numerator kernels (tcgen05 kind::i8), bracketed by CUDA events on their stream in a single-stream
               synchronous region; ``peak`` = the dense u8 MAC rate measured in this run by mtm_measure_i8_peak
               (back-to-back tcgen05.mma from resident shared memory); ``traffic`` = dram bytes of those kernels from a
               child of this run under ``ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`` (nothing else of the
               child is used).
* ``cpu_baseline`` / ``--impl reference``: the CPU port of the reference on live OpenCV (oracle/mtm_port.py;
               /root/reference cannot travel to the GPU box), its own thread pool + cv2's threads, bounded sample.

``--workload C2|C3|C4|C5`` runs one single-image config alone (A/B experiments, profiling).
Only ``torch`` plumbing used: device memory for the resident pools.  Ranks come from the torchrun environment; the
communicator is the library's own (mtm_b200.rendezvous.comm_from_env).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "template-matches/sec"
L2_BYTES = 126 * 1024 * 1024
BATCH_IMAGES = 16
SINGLE_DEFAULT_STEPS = {"C2": 1000, "C3": 100, "C4": 300, "C5": 30}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def build_workload(name, n_images):
    """Seeded synthetic inputs of SURVEY.md 8(d); the image pool is made of circular
    shifts of a few generated scenes (distinct memory, same statistics)."""
    import workloads as synth
    base_imgs = []
    image0, templates, params = synth.config(name, seed=0)
    base_imgs.append(image0)
    for k in range(1, min(4, n_images)):
        img, _, _ = synth.config(name, seed=0, image_index=k)
        base_imgs.append(img)
    pool = []
    for i in range(n_images):
        b = base_imgs[i % len(base_imgs)]
        s = i // len(base_imgs)
        pool.append(np.ascontiguousarray(np.roll(b, (37 * s, 101 * s), axis=(0, 1))))
    return pool, templates, params


def batch_images(n_images):
    """The 16 images of the headline batch: seeds 0..n-1 of the C5 generator (same templates in every scene)."""
    import workloads as synth
    images, templates, params = [], None, None
    for k in range(n_images):
        im, templates, params = synth.config("C5", seed=0, image_index=k)
        images.append(im)
    return images, templates, params


def workload_config(name, world):
    """The `config` object of the JSON line: identical in both arms (ours / reference)."""
    if name == "C5x16":
        per_rank = -(-BATCH_IMAGES // world)
        versions = max(2, int(np.ceil(1.5 * L2_BYTES / (per_rank * 2160 * 3840))))
        return {"workload": "C5x16: BASELINE configs[4], batch of 16 images 3840x2160 uint8 x 64 templates 32..128 px mixed",
                "images": BATCH_IMAGES, "image": [2160, 3840], "templates": 64, "template_shape": "32..128 px square, 64 sizes",
                "score_threshold": 0.5, "maxOverlap": 0.25, "N_object": 50,
                "cut": "images in blocks over the ranks, whole template list on every rank, ONE all-gather of the final hit "
                       "lists inside the timed region (strong scaling: the batch is fixed)",
                "l2_policy": "inputs larger than L2: every rank rotates %d resident versions of its %d-image block (%.0f MB)"
                             % (versions, per_rank, versions * per_rank * 2160 * 3840 / 1e6)}
    shapes = {"C2": ([1080, 1920], 8, [64, 64]), "C3": ([4096, 4096], 1, [256, 256]), "C4": ([2048, 2048], 32, [48, 48]),
              "C5": ([2160, 3840], 64, "32..128 px square, 64 sizes")}
    image, n_t, tshape = shapes[name]
    n_pool = min(128, max(8, int(np.ceil(1.5 * L2_BYTES / (image[0] * image[1])))))
    return {"workload": name, "images": 1, "image": image, "templates": n_t, "template_shape": tshape,
            "score_threshold": 0.5, "maxOverlap": 0.25, "N_object": 50 if name == "C5" else "inf",
            "cut": "one image per step, whole template list (N > 1: independent replicas)",
            "l2_policy": "inputs larger than L2: %d-image device pool (%.0f MB) rotated every step" % (n_pool, n_pool * image[0] * image[1] / 1e6)}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions.  NVML is polled from a thread of this process every
    10 ms (three driver calls per sample, the GIL is released inside them); samples carry perf_counter time stamps and
    only those inside a marked window count.  Falls back to an `nvidia-smi -lms` child when pynvml is unusable."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.windows = []                 # [t_begin, t_end] of the timed regions
        self.samples = []                 # (t, sm_mhz, reason bitmask)
        self.sm_max = None
        self.thread = self.p = self.f = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = gpu_index
            if visible:
                try:
                    index = int(visible.split(",")[gpu_index])
                except (ValueError, IndexError):
                    index = gpu_index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                           "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                          stderr=subprocess.DEVNULL)
            except Exception:
                self.p = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))))
            except Exception:
                pass
            self._stop.wait(0.010)

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        self.windows[-1][1] = time.perf_counter()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            nv = self.nv
            masks = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            def within(slack):
                return [smp for smp in self.samples
                        if any(a - slack <= smp[0] <= (b if b is not None else smp[0]) + slack for a, b in self.windows)]
            inside, note = within(0.0), "inside the timed regions"
            if not inside:                     # timed regions shorter than the polling period (very small --steps)
                inside, note = within(0.03), "within 30 ms of the timed regions (the regions are shorter than the polling period)"
            if inside:
                bits = 0
                for smp in inside:
                    bits |= smp[2]
                out.update(sm_mhz=float(np.median([smp[1] for smp in inside])), samples=len(inside),
                           reasons=[nm for nm in names if bits & masks[nm]],
                           how="NVML polled every 10 ms; samples %s (%.0f ms in total)"
                               % (note, 1e3 * sum((b or a) - a for a, b in self.windows)))
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       how="nvidia-smi -lms 20 over the whole run")
        return out


class NullSampler:
    def begin(self):
        pass

    def end(self):
        pass

    def stop(self):
        return None


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_run(pool, templates, params, steps, warmup):
    """Times the CPU port (reference orchestration on live cv2).  Returns (sec/step, info)."""
    import cv2
    from oracle import mtm_port
    for i in range(warmup):
        mtm_port.match_templates(templates, pool[i % len(pool)], **params)
    t0 = time.perf_counter()
    for i in range(steps):
        mtm_port.match_templates(templates, pool[i % len(pool)], **params)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    info = {"cores": os.cpu_count(), "pool_workers": round(os.cpu_count() * .5), "cv2": cv2.__version__,
            "cv2_threads": cv2.getNumThreads(), "ipp": bool(cv2.ipp.useIPP())}
    return dt, info


def cpu_ncc_only_run(pool, templates, steps):
    """BASELINE.md section 3, baseline B: cv2.matchTemplate(TM_CCOEFF_NORMED) alone -- no peak search, no NMS -- one task per
    template under the reference's pool (MTM/__init__.py:172-175 with only line 92 inside): isolates the third-party kernel."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    arrays = [t[1] for t in templates]
    workers = round(os.cpu_count() * .5)
    t0 = time.perf_counter()
    for i in range(steps):
        image = pool[i % len(pool)]
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(lambda a: cv2.matchTemplate(image, a, cv2.TM_CCOEFF_NORMED), arrays))
    return (time.perf_counter() - t0) / max(steps, 1)


def cpu_baseline(pool, templates, params, name, budget_s=8.0, cpu_steps=None):
    """Bounded sample of the workload on the host cores: whole images (all templates), as many as fit the budget."""
    dt1, _ = cpu_port_run(pool, templates, params, 1, 1)
    steps = cpu_steps if cpu_steps is not None else int(min(100, max(1, round(budget_s / dt1))))
    dt, info = cpu_port_run(pool, templates, params, steps, 0)
    steps_b = max(1, min(steps, int(round(2.0 / max(dt, 1e-6)))))              # ~2 s of baseline B
    dt_b = cpu_ncc_only_run(pool, templates, steps_b)
    return {"value": len(templates) / dt, "unit": "matches/s", "cores": info["cores"], "kind": "port",
            "sample": "%d whole %s images (all %d templates each), %.1f s" % (steps, name, len(templates), dt * steps),
            "ms_per_image": dt * 1e3, **{k: info[k] for k in ("pool_workers", "cv2", "cv2_threads", "ipp")},
            "ncc_only": {"value": len(templates) / dt_b, "unit": "matches/s", "ms_per_image": dt_b * 1e3,
                         "what": "BASELINE.md baseline B: cv2.matchTemplate alone under the same pool, %d images" % steps_b}}


def reference_arm(args):
    """`--impl reference`: the CPU port on the host cores, same metric / config; a step = a bounded sample of the workload."""
    import workloads as synth
    steps = args.steps if args.steps is not None else 10
    warmup = args.warmup if args.warmup is not None else 1
    name = args.workload
    if name == "C5x16":
        pool, templates, params = batch_images(2)
        per_step_units = BATCH_IMAGES * len(templates)
    else:
        pool, templates, params = build_workload(name, 4)
        per_step_units = len(templates)
    # bounded sample: one image per step; a strided subset of the template list when the whole run would pass ~2.5 min
    t0 = time.perf_counter()
    from oracle import mtm_port
    mtm_port.match_templates(templates, pool[0], **params)
    t_image = time.perf_counter() - t0
    n_t = len(templates)
    m = n_t
    if (steps + warmup) * t_image > 150.0:
        m = max(1, int(n_t * 150.0 / ((steps + warmup) * t_image)))
    stride = n_t / m
    subset = [templates[int(i * stride)] for i in range(m)]
    dt, info = cpu_port_run(pool, subset, params, steps, warmup)
    value = m / dt                                              # matches/s of the sample == matches/s of the workload (same images, same list)
    macs = synth.macs(pool[0].shape, subset)
    sample = "each step = ONE image of the workload x %d of its %d templates%s" % (
        m, n_t, "" if m == n_t else " (every %.1f-th, so that %d steps end within minutes)" % (stride, steps))
    line = {"metric": METRIC, "value": value, "unit": "matches/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": per_step_units / value * 1e3, "higher_is_better": True,
            "scaling": "strong" if name == "C5x16" else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "impl": "reference", "config": workload_config(name, max(1, args.gpus)),
            "ms_per_sample_step": dt * 1e3, "gpix_corr_per_s": macs / dt / 1e9,
            "cpu_baseline": {"value": value, "unit": "matches/s", "cores": info["cores"], "kind": "port", "sample": sample,
                             **{k: info[k] for k in ("pool_workers", "cv2", "cv2_threads", "ipp")}},
            "e2e": {"value": value, "unit": "matches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm helpers
class Env:
    """Per-process state of the GPU arm: rank, communicator, contexts, clock sampler."""

    def __init__(self, args):
        import torch
        import MTM  # noqa: F401  (creates the mtm_b200 module alias)
        from mtm_b200 import _native, rendezvous
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.device = _native.local_device()
        torch.cuda.set_device(self.device)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        self.comm = rendezvous.comm_from_env(self.device)          # NCCL inside libmtm_b200.so
        self.n_ctx = max(1, args.contexts)
        self.ctxs = [_native.Context(self.device) for _ in range(self.n_ctx)]
        self.ctx = self.ctxs[0]
        if args.path != "auto":
            for c in self.ctxs:
                c.set_path({"direct": _native.PATH_DIRECT, "tensor": _native.PATH_TENSOR}[args.path])
        self.sampler = ClockSampler(self.device) if self.rank == 0 else NullSampler()
        self.peaks = load_peaks()
        self.i8_peak = None

    def barrier(self):
        self.comm.barrier()
        self.torch.cuda.synchronize()
        for c in self.ctxs:
            c.synchronize()

    def max_over_ranks(self, *vals):
        return self.comm.allreduce_max(list(vals))

    def sum_over_ranks(self, val):
        """all-gather through the MAX all-reduce (every rank fills its own slot of a zero vector)."""
        v = [0.0] * self.world
        v[self.rank] = float(val)
        return float(sum(self.comm.allreduce_max(v))) if self.world > 1 else float(val)

    def measure_i8_peak(self):
        if self.i8_peak is None:
            burst = self.ctx.measure_i8_peak(256, 4000)            # ~0.27 ms per launch
            sustained = self.ctx.measure_i8_peak(256, 3000000)     # ~0.2 s per launch, six launches back to back
            self.i8_peak = {"burst_tmacs": burst, "sustained_tmacs": sustained}
        return self.i8_peak


def roofline_block(env, macs_per_step, ncc_ms_per_step, launches_per_step, step_ms, timed_in, sustained):
    """achieved = ALGORITHMIC 2*MACs / summed duration of the numerator launches of a step; peak = measured kind::i8 rate."""
    pk = env.measure_i8_peak()
    peak_tops = 2.0 * (pk["sustained_tmacs"] if sustained else pk["burst_tmacs"])
    ach = 2.0 * macs_per_step / (ncc_ms_per_step * 1e-3) / 1e12 if ncc_ms_per_step > 0 else 0.0
    bf16 = env.peaks["bf16_tflops_sustained" if sustained else "bf16_tflops"]
    return {"bound": "tensor", "kernel": "ncc numerator, tcgen05 kind::i8 (+fused normalisation epilogue)",
            "achieved": ach, "peak": peak_tops, "unit": "TFLOP/s", "frac": ach / peak_tops if peak_tops else None, "traffic": None,
            "peak_source": "mtm_measure_i8_peak in this run: back-to-back tcgen05.mma kind::i8 M128xN256xK32 from resident shared "
                           "memory, %s (x2 ops per MAC)" % ("0.2 s launches back to back (sustained)" if sustained else "0.27 ms launches (burst)"),
            "peak_i8_measured": {"burst_tops": 2.0 * pk["burst_tmacs"], "sustained_tops": 2.0 * pk["sustained_tmacs"]},
            "frac_of_bf16_peak": ach / bf16, "bf16_peak": bf16, "bf16_peak_source": "%s MEASURED_PEAKS.json (%s)" % (env.peaks["source"], "sustained" if sustained else "burst"),
            "note": "`achieved` counts ALGORITHMIC flops (2*C*h*w per score pixel); the Toeplitz band executes 32*nk/(w*C) times more "
                    "(1.5x at 64- and 256-wide templates)",
            "algorithmic_flops_per_step": 2.0 * macs_per_step, "launches_per_step": launches_per_step,
            "kernel_ms_per_step": ncc_ms_per_step, "kernel_share_of_step": ncc_ms_per_step / step_ms if step_ms else None,
            "timed_in": timed_in}


def run_single(env, args, name, steps, warmup, want_cpu, cpu_steps=None):
    """One single-image config (C2 / C3 / C4 / C5): every rank its own image stream (N > 1: replicas, weak)."""
    import MTM
    import workloads as synth
    torch, ctx, ctxs, n_ctx = env.torch, env.ctx, env.ctxs, env.n_ctx
    probe, templates, params = build_workload(name, 1)
    img_bytes = probe[0].nbytes
    pool_n = min(128, max(8, int(np.ceil(1.5 * L2_BYTES / img_bytes))))          # image pool > L2
    pool, templates, params = build_workload(name, pool_n)
    pool = pool[env.rank % len(pool):] + pool[:env.rank % len(pool)]              # ranks do not process identical streams
    H, W = pool[0].shape[:2]
    n_t = len(templates)
    macs = synth.macs(pool[0].shape, templates)
    n_obj = -1 if params["N_object"] == float("inf") else int(params["N_object"])
    thr, ov = params["score_threshold"], params["maxOverlap"]
    d_pool = torch.empty((pool_n, H, W), dtype=torch.uint8, device="cuda")
    h_pinned = []
    for i, im in enumerate(pool):
        hp = torch.from_numpy(im).pin_memory()
        h_pinned.append(hp)
        d_pool[i].copy_(hp, non_blocking=True)
    torch.cuda.synchronize()
    h_pin = [hp.numpy() for hp in h_pinned]
    tmpl_arrays = [t[1] for t in templates]
    DEPTH = 4

    def resident_step(i):
        ctx.set_image_device(d_pool[i % pool_n].data_ptr(), H, W, 1, W)
        return ctx.match_templates(5, n_obj, thr, ov)

    def resident_stream(first, count):
        """`count` steps through the pipelined entry points, round-robin over the contexts; every result is read on the host."""
        n_hits = 0
        for k in range(count):
            c = ctxs[k % n_ctx]
            slot = (k // n_ctx) % DEPTH
            if k >= DEPTH * n_ctx:
                r = c.match_templates_collect(slot)
                n_hits += -1 if r is None else len(r)
            c.set_image_device(d_pool[(first + k) % pool_n].data_ptr(), H, W, 1, W)
            c.match_templates_async(5, n_obj, thr, ov, slot)
        for k in range(max(0, count - DEPTH * n_ctx), count):
            r = ctxs[k % n_ctx].match_templates_collect((k // n_ctx) % DEPTH)
            n_hits += -1 if r is None else len(r)
        return n_hits

    smp = env.sampler
    for c in ctxs:
        c.set_templates(tmpl_arrays)
    hits = None
    for i in range(warmup):
        hits = resident_step(i)
    resident_stream(0, max(warmup, 2 * n_ctx * DEPTH))
    env.barrier()
    for c in ctxs:
        c.reset_counters()
    smp.begin()
    ctx.timer_begin()
    resident_stream(warmup, steps)
    for c in ctxs:
        c.synchronize()
    ms = ctx.timer_end()
    smp.end()
    launches = sum(c.counters()["kernel_launches"] for c in ctxs)
    env.barrier()
    # single stream, one synchronous call per step: the per-call latency, and the region in which the numerator kernels
    # are bracketed by CUDA events on their own stream without another stream sharing the SMs
    lat_steps = max(10, steps // 3)
    ctx.reset_counters()
    ctx.set_time_ncc(True)
    smp.begin()
    ctx.timer_begin()
    for i in range(lat_steps):
        resident_step(warmup + steps + i)
    lat_ms = ctx.timer_end() / lat_steps
    smp.end()
    ctx.set_time_ncc(False)
    lat_ctr = ctx.counters()
    env.barrier()
    ms_max = env.max_over_ranks(ms)[0]
    value = env.world * n_t * steps / (ms_max * 1e-3)

    # ---- e2e: the public API with host arrays; pageable (what a drop-in caller passes) is the headline, pinned the sub-key
    e2e_steps = max(10, steps // 3)
    e2e = {}
    for kind, arrays in (("pageable", pool), ("pinned", h_pin)):
        for i in range(3):
            MTM.matchTemplates(templates, arrays[i % pool_n], score_threshold=thr, maxOverlap=ov, N_object=params["N_object"], context=ctx)
        env.barrier()
        ctx.reset_counters()
        smp.begin()
        ctx.timer_begin()
        for i in range(e2e_steps):
            MTM.matchTemplates(templates, arrays[(3 + i) % pool_n], score_threshold=thr, maxOverlap=ov, N_object=params["N_object"], context=ctx)
        e_ms = ctx.timer_end()
        smp.end()
        ectr = ctx.counters()
        env.barrier()
        e_max = env.max_over_ranks(e_ms)[0]
        e2e[kind] = {"value": env.world * n_t * e2e_steps / (e_max * 1e-3), "unit": "matches/s", "steps": e2e_steps,
                     "ms_per_step": e_max / e2e_steps, "h2d_bytes_per_step": env.world * ectr["h2d_bytes"] / e2e_steps,
                     "d2h_bytes_per_step": env.world * ectr["d2h_bytes"] / e2e_steps}
    # the same pinned images as one pipelined submission (matchTemplatesBatch: uploads overlap searches)
    batch_imgs = [h_pin[(3 + i) % pool_n] for i in range(e2e_steps)]
    MTM.matchTemplatesBatch(templates, batch_imgs[:8], score_threshold=thr, maxOverlap=ov, N_object=params["N_object"], context=ctx)
    env.barrier()
    smp.begin()
    ctx.timer_begin()
    MTM.matchTemplatesBatch(templates, batch_imgs, score_threshold=thr, maxOverlap=ov, N_object=params["N_object"], context=ctx)
    b_ms = ctx.timer_end()
    smp.end()
    env.barrier()
    b_max = env.max_over_ranks(b_ms)[0]
    out_e2e = dict(e2e["pageable"], api="MTM.matchTemplates(listTemplates, image) with pageable numpy arrays (a drop-in caller's)",
                   pinned=dict(e2e["pinned"], api="the same call with page-locked host arrays"),
                   batch={"value": env.world * n_t * e2e_steps / (b_max * 1e-3), "unit": "matches/s", "ms_per_step": b_max / e2e_steps,
                          "api": "MTM.matchTemplatesBatch(listTemplates, images): the same %d pinned host images as one pipelined submission" % e2e_steps})
    ncc_ms = lat_ctr["ncc_ms"] / lat_steps
    roof = roofline_block(env, macs, ncc_ms, lat_ctr["ncc_launches"] / lat_steps, lat_ms,
                          "single-stream synchronous region of this run (%d steps, %.3f ms/step)" % (lat_steps, lat_ms), sustained=False)
    cpu = None
    if want_cpu and env.world == 1:
        cpu = cpu_baseline(pool, templates, params, name, cpu_steps=cpu_steps)
    return {"metric": METRIC, "value": value, "unit": "matches/s", "ms_per_step": ms_max / steps, "steps": steps, "scaling": "weak",
            "config": dict(workload_config(name, env.world), hits_last_step=len(hits) if hits is not None else None,
                           pipelining="%d contexts (streams) x %d submissions in flight; every step's hit list is read back inside the timed region" % (n_ctx, DEPTH)),
            "sync_ms_per_step": lat_ms, "gpix_corr_per_s": env.world * macs * steps / (ms_max * 1e-3) / 1e9,
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / steps,
            "e2e": out_e2e, "roofline": roof, "cpu_baseline": cpu}


def run_template_sharded(env, args, name, steps, warmup):
    """BASELINE configs[3] at N > 1: the template list of ONE image cut over the ranks (mtm_match_templates_sharded)."""
    import MTM
    import workloads as synth
    from mtm_b200 import sharded
    torch, ctx, comm = env.torch, env.ctx, env.comm
    probe, templates, params = build_workload(name, 1)
    pool_n = min(128, max(8, int(np.ceil(1.5 * L2_BYTES / probe[0].nbytes))))
    pool, templates, params = build_workload(name, pool_n)           # every rank the SAME stream: they cooperate on each image
    H, W = pool[0].shape[:2]
    n_t = len(templates)
    macs = synth.macs(pool[0].shape, templates)
    n_obj = -1 if params["N_object"] == float("inf") else int(params["N_object"])
    thr, ov = params["score_threshold"], params["maxOverlap"]
    d_pool = torch.empty((pool_n, H, W), dtype=torch.uint8, device="cuda")
    for i, im in enumerate(pool):
        d_pool[i].copy_(torch.from_numpy(im))
    torch.cuda.synchronize()
    lo, hi = sharded.weighted_bounds(sharded.template_macs(templates, pool[0].shape), env.world)[env.rank]
    arrays = [t[1] for t in templates[lo:hi]]
    single = MTM.matchTemplates(templates, pool[0], context=env.ctxs[-1], **params)     # the single-GPU answer, on this rank's GPU
    if arrays:
        ctx.set_templates(arrays)

    def step(i):
        ctx.set_image_device(d_pool[i % pool_n].data_ptr(), H, W, 1, W)
        return ctx.match_templates_sharded(comm, lo, len(arrays), 5, n_obj, thr, ov)

    raw = step(0)
    got = [(templates[int(t)][0], (int(x), int(y), int(w), int(h)), float(s)) for t, x, y, w, h, s in
           zip(raw["tmpl"], raw["x"], raw["y"], raw["w"], raw["h"], raw["score"])]
    same = got == [(h[0], tuple(int(v) for v in h[1]), float(h[2])) for h in single]
    for i in range(warmup):
        step(i)
    env.barrier()
    ctx.reset_counters()
    env.sampler.begin()
    ctx.timer_begin()
    for i in range(steps):
        step(warmup + i)
    ms = ctx.timer_end()
    env.sampler.end()
    launches = ctx.counters()["kernel_launches"]
    env.barrier()
    # e2e: the public API with the pageable host image on every rank
    e2e_steps = max(10, steps // 3)
    for i in range(3):
        sharded.matchTemplatesSharded(templates, pool[i], comm=comm, context=ctx, **params)
    env.barrier()
    ctx.reset_counters()
    env.sampler.begin()
    ctx.timer_begin()
    for i in range(e2e_steps):
        sharded.matchTemplatesSharded(templates, pool[(3 + i) % pool_n], comm=comm, context=ctx, **params)
    e_ms = ctx.timer_end()
    env.sampler.end()
    ectr = ctx.counters()
    env.barrier()
    ms_max, e_max, bad = env.max_over_ranks(ms, e_ms, 0.0 if same else 1.0)
    return {"metric": METRIC, "value": n_t * steps / (ms_max * 1e-3), "unit": "matches/s", "ms_per_step": ms_max / steps, "steps": steps,
            "scaling": "strong", "identical_to_single_gpu": bad == 0.0,
            "config": dict(workload_config(name, env.world),
                           cut="template list cut into %d contiguous slices balanced by MACs; ONE ncclAllGather of hit blocks "
                               "(32-B header + 32-B records) on the search stream + replicated global NMS, inside the timed region" % env.world),
            "gpix_corr_per_s": macs * steps / (ms_max * 1e-3) / 1e9, "gpu_launches": int(launches), "gpu_launches_per_step": launches / steps,
            "e2e": {"value": n_t * e2e_steps / (e_max * 1e-3), "unit": "matches/s", "steps": e2e_steps, "ms_per_step": e_max / e2e_steps,
                    "h2d_bytes_per_step": env.world * ectr["h2d_bytes"] / e2e_steps, "d2h_bytes_per_step": env.world * ectr["d2h_bytes"] / e2e_steps,
                    "api": "mtm_b200.sharded.matchTemplatesSharded(listTemplates, image, comm=...) with the pageable host image on every rank"},
            "roofline": None, "cpu_baseline": None}


def run_batch_sharded(env, args, steps, warmup, want_cpu):
    """HEADLINE: the 16-image x 64-template batch, images in blocks over the ranks, one all-gather of the final lists."""
    import MTM
    import workloads as synth
    from mtm_b200 import _native, sharded
    torch, comm, ctxs, n_ctx = env.torch, env.comm, env.ctxs, env.n_ctx
    images, templates, params = batch_images(BATCH_IMAGES)
    n_img, n_t = len(images), len(templates)
    H, W = images[0].shape
    macs_image = synth.macs(images[0].shape, templates)
    n_obj = int(params["N_object"])
    thr, ov = params["score_threshold"], params["maxOverlap"]
    lo, hi, per = sharded.block_bounds(n_img, env.world, env.rank)
    n_local = hi - lo
    versions = max(2, int(np.ceil(1.5 * L2_BYTES / (per * H * W))))
    shifts = [(53 * v, 211 * v) for v in range(versions)]

    def version_of(im, v):
        return im if v == 0 else np.ascontiguousarray(np.roll(im, shifts[v], axis=(0, 1)))

    # resident pool of this rank: versions x its block
    d_pool = torch.empty((versions, max(n_local, 1), H, W), dtype=torch.uint8, device="cuda")
    for v in range(versions):
        for i in range(n_local):
            d_pool[v, i].copy_(torch.from_numpy(version_of(images[lo + i], v)))
    torch.cuda.synchronize()
    tmpl_arrays = [t[1] for t in templates]
    for c in ctxs:
        c.set_templates(tmpl_arrays)
    hits_per_image = min(max(n_obj, 1), _native.SLOT_HITS)
    depth = _native.MAX_INFLIGHT
    assert n_local <= depth * n_ctx

    def resident_batch(v):
        """One step: this rank's block (version v) through the pipelined entry points, then the all-gather; returns
        (hits, counts) of ALL images on the host."""
        entries = []
        for i in range(n_local):
            c = ctxs[i % n_ctx]
            slot = (i // n_ctx) % depth
            c.set_image_device(d_pool[v, i].data_ptr(), H, W, 1, W)
            c.match_templates_async(5, n_obj, thr, ov, slot)
            entries.append((c, slot))
        return comm.gather_results(entries, per, hits_per_image)

    def as_lists(hits, counts):
        out = []
        for g in range(n_img):
            r, i = divmod(g, per)
            k = int(counts[r * per + i])
            raw = hits[r * per + i, :max(k, 0)]
            out.append(None if k < 0 else [(templates[int(t)][0], (int(x), int(y), int(w), int(h)), float(s)) for t, x, y, w, h, s in
                                           zip(raw["tmpl"], raw["x"], raw["y"], raw["w"], raw["h"], raw["score"])])
        return out

    # ---- the single-GPU answer (every rank computes it on its own GPU for version 0) and the parity of the sharded batch
    single = [[(h[0], tuple(int(v) for v in h[1]), float(h[2])) for h in MTM.matchTemplates(templates, im, context=ctxs[0], **params)]
              for im in images]
    got = as_lists(*resident_batch(0))
    same = got == single
    hits_total = sum(len(s) for s in single)
    for w in range(warmup):
        resident_batch(w % versions)
    env.barrier()
    for c in ctxs:
        c.reset_counters()
    smp = env.sampler
    smp.begin()
    ctxs[0].timer_begin()
    for k in range(steps):
        resident_batch((warmup + k) % versions)
    for c in ctxs:
        c.synchronize()
    ms = ctxs[0].timer_end()
    smp.end()
    launches = sum(c.counters()["kernel_launches"] for c in ctxs) + (-(-per // 32)) * steps      # + the pack kernel of the gather
    env.barrier()

    # ---- numerator kernels of this rank's block, single stream, synchronous (roofline region)
    ctx = ctxs[0]
    lat_images = 0
    ctx.reset_counters()
    ctx.set_time_ncc(True)
    smp.begin()
    ctx.timer_begin()
    for rep in range(2 if n_local else 0):
        for i in range(n_local):
            ctx.set_image_device(d_pool[(rep + 1) % versions, i].data_ptr(), H, W, 1, W)
            ctx.match_templates(5, n_obj, thr, ov)
            lat_images += 1
    lat_ms = ctx.timer_end() / max(lat_images, 1)
    smp.end()
    ctx.set_time_ncc(False)
    lat_ctr = ctx.counters()
    env.barrier()

    # ---- e2e: the public API with host images; pageable headline, pinned sub-key.  Only this rank's block is touched.
    e2e_steps = max(3, steps // 3)
    host_versions = min(versions, 3)
    e2e = {}
    e_ctxs = [ctxs[0]] + _native.helper_contexts(ctxs[0].device, n_ctx - 1, owner=ctxs[0])      # the streams the public call uses
    for kind in ("pageable", "pinned"):
        keepalive, lists = [], []
        for v in range(host_versions):
            lst = list(images)                          # images of other ranks are only validated (shape), never read
            for g in range(lo, hi):
                arr = version_of(images[g], v)
                if kind == "pinned":
                    t = torch.from_numpy(arr).pin_memory()
                    keepalive.append(t)
                    arr = t.numpy()
                lst[g] = arr
            lists.append(lst)
        res = sharded.matchTemplatesBatchSharded(templates, lists[0], comm=comm, context=ctxs[0], streams=n_ctx, **params)
        same = same and [[(h[0], tuple(int(v) for v in h[1]), float(h[2])) for h in r] for r in res] == single
        env.barrier()
        for c in e_ctxs:
            c.synchronize()
            c.reset_counters()
        smp.begin()
        ctxs[0].timer_begin()
        for k in range(e2e_steps):
            sharded.matchTemplatesBatchSharded(templates, lists[(1 + k) % len(lists)], comm=comm, context=ctxs[0], streams=n_ctx, **params)
        e_ms = ctxs[0].timer_end()
        smp.end()
        h2d = sum(c.counters()["h2d_bytes"] for c in e_ctxs)
        env.barrier()
        e_max = env.max_over_ranks(e_ms)[0]
        h2d_all = env.sum_over_ranks(h2d)
        gathered_bytes = env.world * per * (32 + hits_per_image * 32)
        e2e[kind] = {"value": n_img * n_t * e2e_steps / (e_max * 1e-3), "unit": "matches/s", "steps": e2e_steps, "ms_per_step": e_max / e2e_steps,
                     "h2d_bytes_per_step": h2d_all / e2e_steps, "d2h_bytes_per_step": env.world * gathered_bytes}
        del keepalive, lists
    ms_max, bad = env.max_over_ranks(ms, 0.0 if same else 1.0)
    ncc_ms_image = lat_ctr["ncc_ms"] / max(lat_images, 1)
    roof = roofline_block(env, macs_image, ncc_ms_image, lat_ctr["ncc_launches"] / max(lat_images, 1), lat_ms,
                          "single-stream synchronous region of this run on rank 0: %d images of its block, %.3f ms/image "
                          "(per IMAGE: the batch step is %d such images per rank)" % (lat_images, lat_ms, per), sustained=True)
    cpu = None
    if want_cpu and env.world == 1:
        cpu = cpu_baseline(images[:2], templates, params, "C5x16", budget_s=10.0, cpu_steps=args.cpu_steps)
    value = n_img * n_t * steps / (ms_max * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": "matches/s", "n_gpus": env.world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config("C5x16", env.world),
            "identical_to_single_gpu": bad == 0.0, "hits_per_batch": hits_total,
            "images_per_rank": per, "ms_per_image_sync": lat_ms,
            "exchange": "mtm_gather_results: one ncclAllGather of %d x (32-B header + %d x 32-B records) per rank, then one D2H of all %d blocks on every rank"
                        % (per, hits_per_image, env.world * per),
            "pipelining": "%d contexts (streams) per rank, submissions in flight until the gather; every image's hit list is read on every rank inside the timed region" % n_ctx,
            "gpix_corr_per_s": n_img * macs_image * steps / (ms_max * 1e-3) / 1e9,
            "gpu_launches": int(env.sum_over_ranks(launches)), "gpu_launches_per_step": env.sum_over_ranks(launches) / steps,
            "e2e": dict(e2e["pageable"], api="mtm_b200.sharded.matchTemplatesBatchSharded(listTemplates, images, comm=...) with pageable numpy images",
                        pinned=dict(e2e["pinned"], api="the same call with page-locked host arrays")),
            "roofline": roof, "cpu_baseline": cpu}
    return line


# ------------------------------------------------------------------------------------------------ dram traffic (ncu child)
def traffic_child(names):
    """Runs under ncu (spawned by measure_traffic): one warm + one measured synchronous step of each workload on one context.
    Prints the number of numerator launches per step so that the parent can cut the launch list."""
    import MTM  # noqa: F401
    from mtm_b200 import _native
    ctx = _native.Context(_native.local_device())
    plan = []
    for name in names:
        wl = "C5" if name == "C5x16" else name
        pool, templates, params = build_workload(wl, 2)
        n_obj = -1 if params["N_object"] == float("inf") else int(params["N_object"])
        ctx.set_templates([t[1] for t in templates])
        ctx.synchronize()
        ctx.reset_counters()
        ctx.set_time_ncc(True)
        for k in range(2):
            ctx.set_image(pool[k])
            ctx.match_templates(5, n_obj, params["score_threshold"], params["maxOverlap"])
        plan.append((name, int(ctx.counters()["ncc_launches"]) // 2))
        ctx.set_time_ncc(False)
    print("TRAFFIC_PLAN " + json.dumps(plan))
    return 0


def measure_traffic(names, timeout=240):
    """dram__bytes_read.sum + dram__bytes_write.sum of the numerator kernels of one step of each workload, from a child of
    this run under ncu.  Returns {name: {...}}; empty when ncu is unavailable (recorded in the line)."""
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {"_error": "ncu not found"}
    with tempfile.TemporaryDirectory() as tmp:
        log = os.path.join(tmp, "traffic.csv")
        cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
               "-k", "regex:ncc_", "--csv", "--log-file", log, sys.executable, os.path.abspath(__file__), "--traffic-child", ",".join(names)]
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        except subprocess.TimeoutExpired:
            return {"_error": "ncu child timed out"}
        plan = None
        for ln in r.stdout.splitlines():
            if ln.startswith("TRAFFIC_PLAN "):
                plan = json.loads(ln[len("TRAFFIC_PLAN "):])
        if plan is None or not os.path.exists(log):
            return {"_error": "ncu child failed: %s" % (r.stderr or r.stdout)[-300:]}
        import csv
        rows = []
        with open(log) as f:
            lines = [ln for ln in f if not ln.startswith("==")]
        for rec in csv.DictReader(lines):
            rows.append(rec)
        # one row per (launch, metric): group by launch ID in order
        launches, order = {}, []
        for rec in rows:
            lid = rec.get("ID")
            if lid not in launches:
                launches[lid] = {"kernel": rec.get("Kernel Name", "")}
                order.append(lid)
            try:
                val = float(rec["Metric Value"].replace(",", ""))
            except (KeyError, ValueError):
                continue
            unit = rec.get("Metric Unit", "")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
            launches[lid][rec["Metric Name"]] = val * scale
        seq = [launches[i] for i in order]
        keep = os.environ.get("MTM_B200_TRAFFIC_CSV")          # keep the raw launch list (profiles/)
        if keep:
            shutil.copyfile(log, keep)
        out, pos = {"_captured": len(seq), "_plan": plan}, 0
        for name, per_step in plan:
            chunk = seq[pos + per_step: pos + 2 * per_step]            # the second (measured) step
            pos += 2 * per_step
            if len(chunk) != per_step or per_step == 0:
                out[name] = None
                continue
            rd = sum(c.get("dram__bytes_read.sum", 0.0) for c in chunk)
            wr = sum(c.get("dram__bytes_write.sum", 0.0) for c in chunk)
            out[name] = {"dram_bytes_per_launch": (rd + wr) / per_step, "dram_bytes_per_step": rd + wr, "dram_read_per_step": rd,
                         "dram_write_per_step": wr, "launches_per_step": per_step,
                         "kernel_ns_per_step_under_ncu": sum(c.get("gpu__time_duration.sum", 0.0) for c in chunk),
                         "kernels": sorted({c["kernel"].split("(")[0][-60:] for c in chunk})}
        return out


def attach_traffic(block, t):
    if block is None or block.get("roofline") is None:
        return
    if t:
        block["roofline"]["traffic"] = t["dram_bytes_per_launch"]
        block["roofline"]["traffic_detail"] = dict(t, how="child of this run under ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum "
                                                          "--clock-control none -k regex:ncc_ (one image, second step)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5x16", help="C5x16 (headline batch, default) or one single-image config C2|C3|C4|C5")
    ap.add_argument("--path", default="auto", choices=["auto", "direct", "tensor"])
    ap.add_argument("--cpu-steps", type=int, default=None, help="images of the cpu_baseline sample (default: ~8 s)")
    ap.add_argument("--contexts", type=int, default=None,
                    help="library contexts (CUDA streams) the throughput loops alternate between (default 4; 3 for the batch)")
    ap.add_argument("--no-per-config", action="store_true", help="headline only (skip the C2/C3/C4 blocks)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child that measures roofline.traffic")
    ap.add_argument("--traffic-child", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.traffic_child:
        return traffic_child(args.traffic_child.split(","))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        args.gpus = world
    if args.impl == "reference":
        return reference_arm(args) if rank == 0 else 0

    headline = args.workload == "C5x16"
    if args.contexts is None:
        args.contexts = 3 if headline else 4
    env = Env(args)
    warmup = max(args.warmup if args.warmup is not None else (3 if headline else 10), 3)
    if headline:
        steps = args.steps if args.steps is not None else 20
        line = run_batch_sharded(env, args, steps, warmup, want_cpu=True)
        per_config = {}
        if not args.no_per_config:
            for name in ("C2", "C3", "C4"):
                if name == "C4" and env.world > 1:
                    per_config[name] = run_template_sharded(env, args, name, SINGLE_DEFAULT_STEPS[name], 5)
                else:
                    per_config[name] = run_single(env, args, name, SINGLE_DEFAULT_STEPS[name], 10, want_cpu=True, cpu_steps=args.cpu_steps)
        line["per_config"] = per_config
    else:
        steps = args.steps if args.steps is not None else SINGLE_DEFAULT_STEPS.get(args.workload, 100)
        block = run_single(env, args, args.workload, steps, warmup, want_cpu=True, cpu_steps=args.cpu_steps)
        line = dict(block, n_gpus=env.world, warmup=warmup, higher_is_better=True, vs_baseline=None, dtype="u8", data="synthetic")
    line["clocks"] = env.sampler.stop()
    env.barrier()
    for c in env.ctxs:
        c.close()
    env.comm.close()
    if env.rank != 0:
        return 0
    if not args.no_traffic:
        names = (["C5x16"] + ([] if args.no_per_config else ["C2", "C3", "C4"])) if headline else [args.workload]
        t = measure_traffic(names)
        if "_error" in t:
            line["traffic_error"] = t["_error"]
        else:
            if any(t.get(n) is None for n in names):
                line["traffic_error"] = "launch list of the ncu child does not match its plan: %d launches captured, plan %r" % (t["_captured"], t["_plan"])
            if headline:
                attach_traffic(line, t.get("C5x16"))
                for name, block in line.get("per_config", {}).items():
                    attach_traffic(block, t.get(name))
            else:
                attach_traffic(line, t.get(args.workload))
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
