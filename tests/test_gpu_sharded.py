"""-m gpu: the multi-GPU entry points of the C ABI (mtm_comm_*, mtm_match_templates_sharded, mtm_gather_results).

On a one-GPU box the whole exchange still runs: ``Comm.create([0, 0, ...])`` builds an in-process loop-back group (one
endpoint and one context per host thread, device-to-device copies ordered by CUDA events), so pack -> gather -> merge ->
replicated NMS is exercised through the same kernels and the same host code as with NCCL.  The NCCL cases (one
process per GPU through ``rendezvous.comm_from_env``, and one process driving two GPUs) need 2 GPUs and skip otherwise.
"""
import os
import pickle
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

from helpers import assert_hits_equal

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    from mtm_b200 import _native
    return _native.load().mtm_device_count()


def run_ranks(devices, fn):
    """fn(rank, comm, ctx) on one host thread per endpoint; returns the per-rank results (exceptions re-raised)."""
    from mtm_b200 import _native
    comms = _native.Comm.create(devices)
    ctxs = [_native.Context(d) for d in devices]
    out, err = [None] * len(devices), [None] * len(devices)

    def body(r):
        try:
            out[r] = fn(r, comms[r], ctxs[r])
        except BaseException as e:       # noqa: BLE001
            err[r] = e

    threads = [threading.Thread(target=body, args=(r,)) for r in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in the exchange"
    for c in comms:
        c.close()
    for c in ctxs:
        c.close()
    return out, err


def test_sharded_world1_equals_match_templates(mtm):
    from mtm_b200 import _native, sharded
    from oracle import golden_cases as gc
    comm = _native.Comm.init_rank(0, 1, 0)
    for name in ("synth_mixed", "synth_rot8", "synth_mixed_n5", "synth_searchbox"):
        kind, temps, img, kw = gc.build(name)
        assert_hits_equal(sharded.matchTemplatesSharded(temps, img, comm=comm, **kw), mtm.matchTemplates(temps, img, **kw), tol=0)
    kind, temps, img, kw = gc.build("synth_mixed")
    images = [img, np.ascontiguousarray(img[::-1]), np.ascontiguousarray(img[:, ::-1])]
    got = sharded.matchTemplatesBatchSharded(temps, images, comm=comm, **kw)
    for g, im in zip(got, images):
        assert_hits_equal(g, mtm.matchTemplates(temps, im, **kw), tol=0)
    comm.close()


@pytest.mark.parametrize("world", [2, 3])
def test_template_cut_loopback_equals_single_gpu(mtm, world):
    """Template slices on `world` endpoints of ONE GPU: every rank returns the single-GPU list, bit for bit."""
    from mtm_b200 import sharded
    from oracle import golden_cases as gc
    cases = []
    for name in ("synth_mixed", "synth_rot8", "synth_mixed_n5", "synth_searchbox", "c1_fish256_inf", "t3_full"):
        kind, temps, img, kw = gc.build(name)
        if kind == "match":
            cases.append((name, temps, img, kw))
    kind, temps, img, kw = gc.build("synth_rot8")
    cases.append(("rot8_n1", temps, img, dict(kw, N_object=1)))                       # minMaxLoc route: one hit per template, global best
    cases.append(("rot8_sqdiff", temps, img, dict(kw, method=1, score_threshold=0.4)))  # minimising method, ascending NMS keys
    cases.append(("rot8_lowthr", temps, img, dict(kw, score_threshold=0.05, maxOverlap=1.0)))   # thousands of raw peaks: block growth + general path
    want = {name: mtm.matchTemplates(t, im, **k) for name, t, im, k in cases}

    def fn(rank, comm, ctx):
        return {name: sharded.matchTemplatesSharded(t, im, comm=comm, context=ctx, **k) for name, t, im, k in cases}

    out, err = run_ranks([0] * world, fn)
    assert err == [None] * world, err
    for r in range(world):
        for name in want:
            assert_hits_equal(out[r][name], want[name], tol=0)
            assert [float(a[2]) for a in out[r][name]] == [float(b[2]) for b in want[name]], name


def test_image_cut_loopback_equals_per_image_calls(mtm):
    from mtm_b200 import sharded
    from oracle import golden_cases as gc
    kind, temps, img, kw = gc.build("synth_mixed")
    images = [img, np.ascontiguousarray(img[::-1]), np.ascontiguousarray(img[:, ::-1]), np.ascontiguousarray(img[::-1, ::-1]),
              np.ascontiguousarray(np.roll(img, 17, axis=1))]
    want = [mtm.matchTemplates(temps, im, **kw) for im in images]

    def fn(rank, comm, ctx):
        a = sharded.matchTemplatesBatchSharded(temps, images, comm=comm, context=ctx, **kw)            # 5 images / 2 ranks: blocks of 3
        b = sharded.matchTemplatesBatchSharded(temps, images[:1], comm=comm, context=ctx, **kw)        # rank 1 idle
        c = sharded.matchTemplatesBatchSharded(temps, images[:4], comm=comm, context=ctx, N_object=2, **{k: v for k, v in kw.items() if k != "N_object"})
        return a, b, c

    out, err = run_ranks([0, 0], fn)
    assert err == [None, None], err
    want_n2 = [mtm.matchTemplates(temps, im, **dict(kw, N_object=2)) for im in images[:4]]
    for r in range(2):
        a, b, c = out[r]
        assert len(a) == 5 and len(b) == 1 and len(c) == 4
        for g, w in zip(a, want):
            assert_hits_equal(g, w, tol=0)
        assert_hits_equal(b[0], want[0], tol=0)
        for g, w in zip(c, want_n2):
            assert_hits_equal(g, w, tol=0)


def test_failing_rank_does_not_hang_the_others(mtm):
    """A rank whose local stage fails (here: a slice that does not match its resident templates) still takes part in the
    exchange; it gets its own error, the others MTM_ERR_PEER."""
    from mtm_b200 import _native
    from oracle import golden_cases as gc
    kind, temps, img, kw = gc.build("synth_rot8")
    arrays = [t[1] for t in temps]

    def fn(rank, comm, ctx):
        ctx.set_image(img)
        ctx.set_templates(arrays[:4] if rank == 0 else arrays[4:])
        n_local = 4 if rank == 0 else 3                      # rank 1 lies about its slice
        try:
            ctx.match_templates_sharded(comm, 4 * rank, n_local, 5, -1, 0.5, 0.25)
        except _native.NativeError as e:
            first = e.code
        else:
            first = 0
        # and the pair still works afterwards
        second = ctx.match_templates_sharded(comm, 4 * rank, 4, 5, -1, 0.5, 0.25)
        return first, second

    out, err = run_ranks([0, 0], fn)
    assert err == [None, None], err
    assert out[0][0] == _native.MTM_ERR_PEER and out[1][0] == _native.MTM_ERR_INVALID
    want = mtm.matchTemplates(temps, img, **kw)
    for r in range(2):
        raw = out[r][1]
        got = [(temps[int(t)][0], (int(x), int(y), int(w), int(h)), s) for t, x, y, w, h, s in
               zip(raw["tmpl"], raw["x"], raw["y"], raw["w"], raw["h"], raw["score"])]
        assert_hits_equal(got, want, tol=0)


def test_allreduce_max_and_barrier_loopback(mtm):
    def fn(rank, comm, ctx):
        comm.barrier()
        return comm.allreduce_max([float(rank), 10.0 - rank, 3.5])

    out, err = run_ranks([0, 0, 0], fn)
    assert err == [None] * 3 and out == [[2.0, 10.0, 3.5]] * 3


# ---------------------------------------------------------------------------------------------- NCCL (2 GPUs)
_NCCL_WORKER = r"""
import os, pickle, sys
sys.path.insert(0, %(root)r)
import numpy as np
import MTM
from mtm_b200 import rendezvous, sharded
from oracle import synth
comm = rendezvous.comm_from_env()
image, temps, params = synth.config("C4")
got = sharded.matchTemplatesSharded(temps, image, comm=comm, **params)
images = [synth.config("C2", seed=0, image_index=k)[0] for k in range(3)]
_, temps2, params2 = synth.config("C2")
batch = sharded.matchTemplatesBatchSharded(temps2, images, comm=comm, **params2)
mx = comm.allreduce_max([float(comm.rank)])
with open(os.path.join(%(out)r, "rank%%d.pkl" %% comm.rank), "wb") as f:
    pickle.dump(([(h[0], tuple(h[1]), float(h[2])) for h in got],
                 [[(h[0], tuple(h[1]), float(h[2])) for h in hits] for hits in batch], mx), f)
comm.close()
"""


def test_nccl_one_process_per_gpu_world2(tmp_path, mtm):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", _NCCL_WORKER % {"root": ROOT, "out": str(tmp_path)}], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), logs
    res = [pickle.load(open(tmp_path / ("rank%d.pkl" % r), "rb")) for r in range(2)]
    assert res[0] == res[1] and res[0][2] == [1.0]
    image, temps, params = synth.config("C4")
    assert_hits_equal(res[0][0], mtm.matchTemplates(temps, image, **params), tol=0)
    _, temps2, params2 = synth.config("C2")
    for k in range(3):
        assert_hits_equal(res[0][1][k], mtm.matchTemplates(temps2, synth.config("C2", seed=0, image_index=k)[0], **params2), tol=0)


def test_nccl_one_process_two_gpus(mtm):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from mtm_b200 import sharded
    from oracle import golden_cases as gc
    kind, temps, img, kw = gc.build("synth_rot8")
    want = mtm.matchTemplates(temps, img, **kw)

    def fn(rank, comm, ctx):
        return sharded.matchTemplatesSharded(temps, img, comm=comm, context=ctx, **kw)

    out, err = run_ranks([0, 1], fn)
    assert err == [None, None], err
    for r in range(2):
        assert_hits_equal(out[r], want, tol=0)
