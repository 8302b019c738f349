"""Randomised differential runs on the CPU emulation (TEST INFRASTRUCTURE, a command-line tool -- not collected by pytest).

    python tests/fuzz_emulated.py seq    SEED0 N     random call sequences through the host build of the WHOLE library (one context)
                                                     against the port on exact maps; MTM_B200_* knobs apply (e.g. MTM_B200_MOM_BOX=1);
                                                     FUZZ_LIVE_CV2=1 compares with the port on live cv2 instead (its fp32 noise shows)
    python tests/fuzz_emulated.py tc     SEED0 N     tensor-core kernels on the tcgen05 model: random geometry / tiling / ring / warps
    python tests/fuzz_emulated.py post   SEED0 N     peak extraction + sort + NMS kernels on random maps full of ties
    python tests/fuzz_emulated.py direct SEED0 N     dp4a kernel, random shapes (templates as large as the image included)
    python tests/fuzz_emulated.py f32    SEED0 N     float32 kernels

Every mode prints one line per case and "bad: K" at the end.  The seeded cases of tests/test_kernel_emulation.py and
tests/test_library_emulation.py are samples of these generators."""
import os
import pathlib
import sys
import tempfile
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


class _Tmp:
    def mktemp(self, name):
        return pathlib.Path(tempfile.mkdtemp(prefix="mtm_fuzz_"))


def _kernel_lib():
    import test_kernel_emulation as tk
    return tk, tk.emu._get_wrapped_function()(_Tmp())


def fuzz_seq(seed0, n):
    import emu_library
    import MTM
    from mtm_b200 import _native
    from oracle import mtm_port, ncc_exact, synth
    _native.LIB_PATH = emu_library.build(tempfile.mkdtemp(prefix="mtm_emu_fz_"))
    _native._lib = None
    warnings.simplefilter("ignore")

    def _exact_map(template, image, method=5, mask=None):          # the port on EXACT maps: no fp32-DFT noise near thresholds / ties
        if not (template.dtype == np.uint8 and image.dtype == np.uint8):
            template, image = np.float32(template), np.float32(image)
            if mask is not None:
                mask = np.float32(mask)
        if mask is not None and method in (0, 3) and mask.shape == template.shape and mask.dtype == template.dtype:
            return ncc_exact.match_template_masked_exact(image, template, mask, method)
        return ncc_exact.match_template_exact(image, template, method)

    if os.environ.get("FUZZ_LIVE_CV2") != "1":
        mtm_port.compute_score_map = _exact_map
    bad = 0
    for seed in range(seed0, seed0 + n):
        rng = np.random.default_rng(seed)
        pool_t = {}
        def templates(kind, C):
            key = (kind, C)
            if key not in pool_t:
                k = int(rng.integers(1, 5))
                if kind == 0: shapes = [(int(rng.integers(8, 25)), int(rng.integers(8, 30)))] * k
                else: shapes = [(int(rng.integers(6, 25)), int(rng.integers(6, 30))) for _ in range(k)]
                ts = [synth.make_template(rng, h, w) for h, w in shapes]
                if C == 3: ts = [np.ascontiguousarray(np.stack([t, 255 - t, t[::-1, ::-1]], axis=2)) for t in ts]
                pool_t[key] = [("t%d" % i, t) for i, t in enumerate(ts)]
            return pool_t[key]
        images = {}
        def image(idx, C, temps):
            key = (idx, C, id(temps))
            if key not in images:
                H, W = [(90, 130), (90, 130), (70, 101)][idx]
                gray = [t[1] if t[1].ndim == 2 else t[1][:, :, 0] for t in temps]
                img, _ = synth.make_scene(H, W, gray, 2, seed=seed * 10 + idx)
                if C == 3: img = np.ascontiguousarray(np.stack([img, 255 - img, img[::-1, ::-1]], axis=2))
                images[key] = img
            return images[key]
        t0 = time.time()
        for step in range(10):
            C = int(rng.choice([1, 1, 3]))
            temps = templates(int(rng.integers(0, 2)), C)
            img = image(int(rng.integers(0, 3)), C, temps)
            dt = rng.choice(["u8", "u8", "u8", "f32", "u16"]) if C == 1 else rng.choice(["u8", "u8", "f32"])
            if dt == "f32": im, ts = img.astype(np.float32), [(n_, t.astype(np.float32)) for n_, t in temps]
            elif dt == "u16": im, ts = img.astype(np.uint16) * 150, [(n_, t.astype(np.uint16) * 150) for n_, t in temps]
            else: im, ts = img, temps
            method = int(rng.choice([5, 5, 5, 1, 3]))
            n_object = [float("inf"), 1, int(rng.integers(2, 6))][int(rng.integers(0, 3))]
            thr = float(rng.choice([0.4, 0.5, 0.7])) if method != 1 else float(rng.choice([0.3, 0.4]))
            if method == 3: thr = 0.9
            sb = None
            if rng.random() < 0.3:
                H, W = im.shape[:2]
                sb = (int(rng.integers(0, 10)), int(rng.integers(0, 10)), W - 12, H - 12)
            kw = dict(method=method, N_object=n_object, score_threshold=thr, maxOverlap=float(rng.choice([0.0, 0.25, 0.5])), searchBox=sb)
            op = int(rng.integers(0, 5))
            if op == 3 and dt == "u16": op = 0
            try:
                if op == 3:                                  # masked templates, method 3 (TM_CCORR_NORMED), through matchTemplates
                    tm = [(n_, t, (rng.random(t.shape) > 0.2).astype(t.dtype) if t.dtype == np.uint8 else rng.random(t.shape).astype(np.float32)) for n_, t in ts]
                    kw3 = dict(kw); kw3["method"] = 3; kw3["score_threshold"] = 0.9
                    got = MTM.matchTemplates(tm, im, **kw3); want = mtm_port.match_templates(tm, im, **kw3)
                    ok = [(h[0], h[1]) for h in got] == [(h[0], h[1]) for h in want] and all(abs(float(a[2]) - float(b[2])) <= 1e-4 for a, b in zip(got, want))
                elif op == 4:                                # the batch entry point on two images of the pool
                    ims = [im, im[::-1].copy()]
                    kwb = {k: v for k, v in kw.items() if k != "searchBox"}
                    got = MTM.matchTemplatesBatch(ts, ims, **kwb); want = [mtm_port.match_templates(ts, i2, **kwb) for i2 in ims]
                    ok = [[(h[0], h[1]) for h in g] for g in got] == [[(h[0], h[1]) for h in w] for w in want]
                elif op == 0:
                    got = MTM.matchTemplates(ts, im, **kw); want = mtm_port.match_templates(ts, im, **kw)
                    ok = [(h[0], h[1]) for h in got] == [(h[0], h[1]) for h in want] and all(abs(float(a[2]) - float(b[2])) <= 1e-4 for a, b in zip(got, want))
                elif op == 1:
                    kw2 = {k: v for k, v in kw.items() if k != "maxOverlap"}
                    got = MTM.findMatches(ts, im, **kw2); want = mtm_port.find_matches(ts, im, **kw2)
                    G = set((h[0], h[1]) for h in got); Wn = set((h[0], h[1]) for h in want)
                    near = lambda a, B: any(a[0] == b[0] and abs(a[1][0] - b[1][0]) <= 1 and abs(a[1][1] - b[1][1]) <= 1 for b in B)
                    ok = all(near(a, Wn - G) for a in G - Wn) and all(near(b, G - Wn) for b in Wn - G)      # cv2's own fp32 noise may move a peak by a pixel
                elif op == 2:
                    t = ts[int(rng.integers(0, len(ts)))][1]
                    got = MTM.computeScoreMap(t, im, method=method); want = mtm_port.compute_score_map(t, im, method)
                    sc = max(1.0, float(np.abs(want).max()))
                    ok = got.shape == want.shape and float(np.max(np.abs(got - want))) <= 2e-4 * sc
            except Exception as e:
                ok = False; print("   EXC", repr(e)[:300])
            if not ok:
                bad += 1
                print("seed", seed, "step", step, "MISMATCH", dict(op=op, C=C, dt=str(dt), n_t=len(ts), shapes=[t[1].shape for t in ts], img=im.shape, **kw), flush=True)
                if op != 2: print("   got ", got[:5], "\n   want", want[:5])
        print("seed", seed, "done %.0fs" % (time.time() - t0), flush=True)
    print("bad:", bad)

def fuzz_tc(seed0, n):
    from oracle import ncc_exact
    tk, emu = _kernel_lib()
    bad = 0
    for case in range(seed0, seed0 + n):
        rng = np.random.default_rng(case)
        C = int(rng.choice([1, 1, 3, 4]))
        mode = int(rng.integers(0, 2)) if C == 1 else 0
        H = int(rng.integers(12, 90)); W = int(rng.integers(20, 200 if mode == 0 else 400))
        count = 1 if mode == 1 else int(rng.integers(1, 9))
        shapes = []
        hmax = int(rng.integers(4, min(H, 40) + 1)); wmax = int(rng.integers(4, min(W, 60) + 1))
        for k in range(count):
            if k == 0 or rng.random() < 0.5:
                shapes.append((hmax, wmax))
            else:
                shapes.append((int(rng.integers(max(1, hmax // 2), hmax + 1)), int(rng.integers(max(1, wmax // 2), wmax + 1))))
        shapes = [(h, w) for h, w in shapes if h * w >= 16] or [(hmax, max(wmax, 4))]
        if shapes[0][0] * shapes[0][1] < 16: continue
        N = int(rng.choice([16, 32, 48, 64, 96]))
        opts = dict(stages=int(rng.integers(2, 6)), ds=int(rng.integers(1, 3)), EW=int(rng.choice([8, 12])), persist=bool(rng.random() < 0.8), ctas=int(rng.integers(1, 5)))
        image, tmpls = tk._planted(rng, H, W, C, shapes)
        method = int(rng.choice([2, 5, 5, 3]))
        t0 = time.time()
        try:
            got, n_mma, _ = tk._host_tensor_maps(emu, image, tmpls, method, mode, N, **opts)
        except AssertionError as e:
            print("case", case, "skipped/failed launch", e, dict(C=C, mode=mode, H=H, W=W, shapes=shapes, N=N, **opts)); continue
        ok = True
        for k, t in enumerate(tmpls):
            want = ncc_exact.match_template_exact(image, t, method)
            if method == 5:
                good = np.max(np.abs(got[k] - want)) <= 2e-6
            else:
                good = np.array_equal(got[k].view(np.uint32), want.view(np.uint32))
            ok = ok and good
        print("case", case, "ok" if ok else "MISMATCH", dict(C=C, mode=mode, H=H, W=W, shapes=shapes, N=N, method=method, **opts), "%.1fs" % (time.time() - t0), flush=True)
        bad += (not ok)
    print("bad:", bad)

def fuzz_post(seed0, n):
    from oracle import ncc_exact
    tk, emu = _kernel_lib()
    bad = 0
    for case in range(seed0, seed0 + n):
        rng = np.random.default_rng(case)
        nt = int(rng.integers(1, 7))
        levels = int(rng.choice([2, 4, 16, 1000]))
        maps, sizes = [], []
        for k in range(nt):
            kind = rng.integers(0, 6)
            if kind == 0: shp = (1, int(rng.integers(1, 50)))
            elif kind == 1: shp = (int(rng.integers(1, 50)), 1)
            else: shp = (int(rng.integers(2, 40)), int(rng.integers(2, 60)))
            m = (np.round(rng.random(shp) * levels) / levels).astype(np.float32)
            if rng.random() < 0.15: m[:] = m.flat[0]
            maps.append(m); sizes.append((int(rng.integers(1, 40)), int(rng.integers(1, 40))))
        method = int(rng.choice([1, 3, 5]))
        n_object = [float("inf"), 1, int(rng.integers(2, 30))][int(rng.integers(0, 3))]
        thr = float(rng.choice([0.0, 0.25, 0.5, 0.75, 0.9]))
        mo = float(rng.choice([0.0, 0.1, 0.25, 0.5, 1.0]))
        do_nms = bool(rng.random() < 0.7)
        fg = bool(rng.random() < 0.4)
        try:
            want = tk._port_postprocess(maps, sizes, method, n_object, thr, mo, do_nms)
            got, route = tk._host_postprocess(emu, maps, sizes, method, n_object, thr, mo, do_nms, fg)
        except Exception as e:
            print("case", case, "EXC", repr(e)[:200]); bad += 1; continue
        ok = got == want
        print("case", case, "ok" if ok else "MISMATCH", dict(nt=nt, levels=levels, shapes=[m.shape for m in maps], method=method, n_object=n_object, thr=thr, mo=mo, do_nms=do_nms, fg=fg, route=route, n=len(want)), flush=True)
        if not ok:
            bad += 1
            print("   want", want[:6]); print("   got ", got[:6])
    print("bad:", bad)

def fuzz_direct(seed0, n):
    from oracle import ncc_exact
    tk, emu = _kernel_lib()
    bad = 0
    for case in range(seed0, seed0 + n):
        rng = np.random.default_rng(case)
        C = int(rng.choice([1, 3, 4]))
        H = int(rng.integers(3, 80)); W = int(rng.integers(3, 150))
        h = int(rng.integers(1, H + 1)); w = int(rng.integers(1, W + 1))
        if rng.random() < 0.5: h = min(h, 12); w = min(w, 20)
        count = int(rng.integers(1, 10))
        image, tmpls = tk._planted(rng, H, W, C, [(h, w)] * count) if H > 20 and W > 50 else (None, None)
        if image is None:
            img = rng.integers(0, 256, (H, W, C)).astype(np.uint8)
            tmpls = [rng.integers(0, 256, (h, w, C)).astype(np.uint8) for _ in range(count)]
            image = img if C > 1 else img[:, :, 0]
            if C == 1: tmpls = [t[:, :, 0] for t in tmpls]
        methods = [int(m) for m in rng.choice(6, 2, replace=False)]
        got, tt = tk._host_direct_maps(emu, image, tmpls, methods)
        ok = True
        for m in methods:
            for k, t in enumerate(tmpls):
                want = ncc_exact.match_template_exact(image, t, m)
                ok = ok and np.array_equal(got[m][k].view(np.uint32), want.view(np.uint32))
        print("case", case, "ok" if ok else "MISMATCH", dict(C=C, H=H, W=W, h=h, w=w, count=count, methods=methods, tt=tt), flush=True)
        bad += (not ok)
    print("bad:", bad)

def fuzz_f32(seed0, n):
    from oracle import ncc_exact
    tk, emu = _kernel_lib()
    bad = 0
    for case in range(seed0, seed0 + n):
        rng = np.random.default_rng(case)
        C = int(rng.choice([1, 3, 4]))
        H = int(rng.integers(3, 70)); W = int(rng.integers(3, 120))
        h = int(rng.integers(1, min(H, 30) + 1)); w = int(rng.integers(1, min(W, 40) + 1))
        count = int(rng.integers(1, 7))
        kind = rng.integers(0, 3)
        img = rng.random((H, W, C)).astype(np.float32) * (1.0 if kind == 0 else 255.0 if kind == 1 else 65535.0)
        if kind: img = np.floor(img)
        tmpls = []
        for k in range(count):
            y0, x0 = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
            tmpls.append(np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w] * np.float32(0.9) + np.float32(rng.random()) * img.max() * np.float32(0.05)))
        image = img if C > 1 else img[:, :, 0]
        if C == 1: tmpls = [t[:, :, 0] for t in tmpls]
        methods = [int(m) for m in rng.choice(6, 3, replace=False)]
        got, tt, *_ = tk._host_f32_maps(emu, image, tmpls, methods)
        ok = True; worst = 0.0
        for m in methods:
            for k, t in enumerate(tmpls):
                want = ncc_exact.match_template_exact(image, t, m)
                sc = max(1.0, float(np.abs(want).max()))
                e = float(np.max(np.abs(got[m][k].astype(np.float64) - want))) / sc
                worst = max(worst, e)
                ok = ok and e <= 1e-4
        print("case", case, "ok" if ok else "MISMATCH", dict(C=C, H=H, W=W, h=h, w=w, count=count, kind=int(kind), methods=methods, worst=worst), flush=True)
        bad += (not ok)
    print("bad:", bad)


if __name__ == "__main__":
    mode, seed0, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    {"seq": fuzz_seq, "tc": fuzz_tc, "post": fuzz_post, "direct": fuzz_direct, "f32": fuzz_f32}[mode](seed0, n)
