"""CPU test (gloo, world_size 2) of the multi-GPU host logic: template sharding, the
all-gather of hit rows and the replicated global NMS must reproduce the unsharded result.
The per-rank search/NMS are injected from the oracle here (no GPU); on the GPU box the same
code path runs with the CUDA functions (tests/test_gpu_sharded.py)."""
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
    from oracle import golden_cases as gc, mtm_port
    import pickle
    results = {}
    for name in ("synth_mixed", "synth_rot8", "synth_mixed_n5", "synth_searchbox"):
        kind, temps, img, kw = gc.build(name)
        got = sharded.matchTemplatesSharded(temps, img, find_fn=mtm_port.find_matches, nms_fn=mtm_port.nms, **kw)
        results[name] = [(h[0], tuple(h[1]), float(h[2])) for h in got]
    # a rank with an empty shard (1 template, 2 ranks) and one with no hits at all
    kind, temps, img, kw = gc.build("c1_fish256_inf")
    got = sharded.matchTemplatesSharded(temps, img, find_fn=mtm_port.find_matches, nms_fn=mtm_port.nms, **kw)
    results["c1_fish256_inf"] = [(h[0], tuple(h[1]), float(h[2])) for h in got]
    with open(os.path.join(out_dir, "rank%d.pkl" % rank), "wb") as f:
        pickle.dump(results, f)
    dist.destroy_process_group()


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


def test_sharded_match_templates_world2(tmp_path, golden):
    import pickle
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [pickle.load(open(tmp_path / ("rank%d.pkl" % r), "rb")) for r in range(2)]
    assert res[0] == res[1]                                    # replicated NMS -> identical on every rank
    for name, got in res[0].items():
        want = [(r[0], tuple(r[1]), r[2]) for r in golden[name]]
        assert [(g[0], g[1]) for g in got] == [(w[0], w[1]) for w in want], name
        assert max(abs(g[2] - w[2]) for g, w in zip(got, want)) < 1e-6
