"""CPU test (gloo, world_size 2) of the multi-GPU host logic: template sharding, the
all-gather of hit rows and the replicated global NMS must reproduce the unsharded result.
The per-rank search and the exchange are injected here (oracle + gloo, no GPU); on the GPU box the same
host code runs on the library's own exchange (tests/test_gpu_sharded.py: NCCL / in-process loop-back)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _GlooComm:
    """Host-side stand-in for ``_native.Comm``: rank / world of the gloo group (no GPU here)."""

    def __init__(self, dist):
        self.world, self.rank, self.device = dist.get_world_size(), dist.get_rank(), 0


def _doubles(dist, comm):
    """Oracle-backed test doubles of the two device paths of mtm_b200.sharded: the per-rank search runs on the CPU port,
    the exchange is a gloo all_gather_object, everything after it follows the library's contract (rank-ordered
    concatenation, replicated NMS / fixed blocks of ``images_per_rank`` entries with counts -1 = unused)."""
    from mtm_b200 import _native
    from oracle import mtm_port

    def search_fn(arrays, masks, img, lo, method, n_dev, thr, overlap):
        n_object = float("inf") if n_dev < 0 else n_dev
        slice_list = [(k, a) if m is None else (k, a, m) for k, (a, m) in enumerate(zip(arrays, masks))]
        local = mtm_port.find_matches(slice_list, img, method, n_object, thr) if slice_list else []
        rows = [(lo + k, tuple(int(v) for v in box), float(score)) for k, box, score in local]
        gathered = [None] * comm.world
        dist.all_gather_object(gathered, rows)
        merged = [r for part in gathered for r in part]                 # rank order == template-list order
        kept = mtm_port.nms(merged, thr, method == 1, n_object, overlap)
        raw = np.zeros(len(kept), _native.HIT_DTYPE)
        for i, (t, box, score) in enumerate(kept):
            raw[i] = (t, box[0], box[1], box[2], box[3], score)
        return raw

    def batch_fn(prepared, method, n_dev, thr, overlap, per, hits_per_image):
        n_object = float("inf") if n_dev < 0 else n_dev
        mine = []
        for arrays, masks, img in prepared:
            temps = [(k, a) if m is None else (k, a, m) for k, (a, m) in enumerate(zip(arrays, masks))]
            mine.append([(t, tuple(int(v) for v in box), float(sc)) for t, box, sc in
                         mtm_port.match_templates(temps, img, method, n_object, thr, overlap)])
        gathered = [None] * comm.world
        dist.all_gather_object(gathered, mine)
        hits = np.zeros((comm.world * per, hits_per_image), _native.HIT_DTYPE)
        counts = np.full(comm.world * per, -1, np.int32)
        for r, part in enumerate(gathered):
            for i, lst in enumerate(part):
                counts[r * per + i] = len(lst)
                for k, (t, box, sc) in enumerate(lst):
                    hits[r * per + i, k] = (t, box[0], box[1], box[2], box[3], sc)
        return hits, counts

    return search_fn, batch_fn


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import MTM  # noqa: F401
    from mtm_b200 import sharded
    from oracle import golden_cases as gc
    import pickle
    comm = _GlooComm(dist)
    search_fn, batch_fn = _doubles(dist, comm)
    results = {}
    for name in ("synth_mixed", "synth_rot8", "synth_mixed_n5", "synth_searchbox"):
        kind, temps, img, kw = gc.build(name)
        got = sharded.matchTemplatesSharded(temps, img, comm=comm, search_fn=search_fn, **kw)
        results[name] = [(h[0], tuple(h[1]), float(h[2])) for h in got]
    # a rank with an empty shard (1 template, 2 ranks) and one with no hits at all
    kind, temps, img, kw = gc.build("c1_fish256_inf")
    got = sharded.matchTemplatesSharded(temps, img, comm=comm, search_fn=search_fn, **kw)
    results["c1_fish256_inf"] = [(h[0], tuple(h[1]), float(h[2])) for h in got]
    # validation errors are raised on every rank before any collective (a rank-local raise would hang the others)
    try:
        sharded.matchTemplatesSharded(temps, img[:10, :10], comm=comm, search_fn=search_fn, **kw)
        results["too_large"] = "no error"
    except ValueError as e:
        results["too_large"] = str(e)
    # the other cut (SURVEY 8e, configs[4]): images in blocks over the ranks, templates whole, one all-gather of the final lists
    kind, temps, img, kw = gc.build("synth_mixed")
    images = [img, np.ascontiguousarray(img[::-1]), np.ascontiguousarray(img[:, ::-1])]
    for name, ims in (("batch3", images), ("batch1", images[:1])):          # 3 images / 2 ranks; 1 image -> rank 1 idle
        got = sharded.matchTemplatesBatchSharded(temps, ims, comm=comm, batch_fn=batch_fn, **kw)
        results[name] = [[(h[0], tuple(h[1]), float(h[2])) for h in hits] for hits in got]
    with open(os.path.join(out_dir, "rank%d.pkl" % rank), "wb") as f:
        pickle.dump(results, f)
    dist.destroy_process_group()


def test_block_bounds_cover_everything():
    import MTM  # noqa: F401
    from mtm_b200.sharded import block_bounds
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            spans = [block_bounds(n, world, r) for r in range(world)]
            per = spans[0][2]
            assert all(sp[2] == per for sp in spans) and per * world >= n
            assert spans[0][0] == 0 and max(sp[1] for sp in spans) == n
            assert all(a[1] == b[0] or b[0] == b[1] == n for a, b in zip(spans, spans[1:]))
            assert all(sp[0] == min(n, r * per) for r, sp in enumerate(spans))


def test_shard_bounds_cover_everything():
    import MTM  # noqa: F401
    from mtm_b200.sharded import shard_bounds
    for n in (0, 1, 7, 8, 33, 64):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_mac_balanced_template_slices():
    """Mixed template sizes (configs[4]: 64 templates of 32..128 px): slices balanced by work, not by count."""
    import MTM  # noqa: F401
    from mtm_b200.sharded import template_macs, weighted_bounds
    sides = np.linspace(32, 128, 64).round().astype(int)
    temps = [("t%d" % i, np.zeros((s, s), np.uint8)) for i, s in enumerate(sides)]
    macs = template_macs(temps, (2160, 3840))
    assert macs[0] == 32 * 32 * (2160 - 31) * (3840 - 31)
    for world in (1, 2, 4, 8):
        spans = weighted_bounds(macs, world)
        assert spans[0][0] == 0 and spans[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        loads = [sum(macs[a:b]) for a, b in spans]
        assert max(loads) <= 1.1 * sum(macs) / world                      # the count split reaches 2.05x at world 8
    assert weighted_bounds([], 3) == [(0, 0)] * 3 and weighted_bounds([5.0], 2) in ([(0, 0), (0, 1)], [(0, 1), (1, 1)])
    assert template_macs([("a", np.zeros((10, 20, 3), np.uint8))], (100, 200, 3), searchBox=(5, 5, 50, 40)) == [3.0 * 200 * 31 * 31]


def test_weighted_bounds_is_optimal_on_small_lists():
    import itertools
    import random
    import MTM  # noqa: F401
    from mtm_b200.sharded import weighted_bounds
    random.seed(0)
    for _ in range(150):
        n, world = random.randint(0, 8), random.randint(1, 4)
        w = [random.randint(1, 50) for _ in range(n)]
        spans = weighted_bounds(w, world)
        assert len(spans) == world and spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(spans, spans[1:] + [(n, n)]))
        cost = max(sum(w[a:b]) for a, b in spans)
        best = min(max(sum(w[b[i]:b[i + 1]]) for i in range(world))
                   for cuts in itertools.combinations_with_replacement(range(n + 1), world - 1)
                   for b in [[0] + list(cuts) + [n]])
        assert cost == best, (w, world, spans)


def test_sharded_match_templates_world2(tmp_path, golden):
    import pickle
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [pickle.load(open(tmp_path / ("rank%d.pkl" % r), "rb")) for r in range(2)]
    assert res[0] == res[1]                                    # replicated NMS -> identical on every rank
    assert "larger than image" in res[0].pop("too_large")
    from oracle import golden_cases as gc, mtm_port
    kind, temps, img, kw = gc.build("synth_mixed")
    images = [img, np.ascontiguousarray(img[::-1]), np.ascontiguousarray(img[:, ::-1])]
    for name, ims in (("batch3", images), ("batch1", images[:1])):
        got = res[0].pop(name)
        want = [mtm_port.match_templates(temps, im, **kw) for im in ims]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert [(a[0], a[1], a[2]) for a in g] == [(b[0], tuple(b[1]), float(b[2])) for b in w]
    for name, got in res[0].items():
        want = [(r[0], tuple(r[1]), r[2]) for r in golden[name]]
        assert [(g[0], g[1]) for g in got] == [(w[0], w[1]) for w in want], name
        assert max(abs(g[2] - w[2]) for g, w in zip(got, want)) < 1e-6
