"""-m gpu: the sharded entry point with the CUDA functions.  world_size 1 always; a 2-rank NCCL
run when the box has 2+ GPUs (gpurun --gpus 2)."""
import os
import socket

import pytest

from helpers import assert_hits_equal

pytestmark = pytest.mark.gpu


def test_sharded_world1_equals_match_templates(mtm):
    from mtm_b200 import sharded
    from oracle import golden_cases as gc
    for name in ("synth_mixed", "synth_rot8", "synth_mixed_n5", "synth_searchbox"):
        kind, temps, img, kw = gc.build(name)
        assert_hits_equal(sharded.matchTemplatesSharded(temps, img, **kw), mtm.matchTemplates(temps, img, **kw), tol=0)


def test_batch_sharded_world1_equals_per_image_calls(mtm):
    """The image-sharded cut (SURVEY 8e, configs[4]) without a process group: == the per-image calls."""
    import numpy as np
    from mtm_b200 import sharded
    from oracle import golden_cases as gc
    kind, temps, img, kw = gc.build("synth_mixed")
    images = [img, np.ascontiguousarray(img[::-1]), np.ascontiguousarray(img[:, ::-1])]
    got = sharded.matchTemplatesBatchSharded(temps, images, **kw)
    for g, im in zip(got, images):
        assert_hits_equal(g, mtm.matchTemplates(temps, im, **kw), tol=0)


def _worker(rank, world, port, out_dir):
    import pickle
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import MTM  # noqa: F401
    from mtm_b200 import sharded
    from oracle import synth
    image, temps, params = synth.config("C4")
    got = sharded.matchTemplatesSharded(temps, image, **params)
    with open(os.path.join(out_dir, "rank%d.pkl" % rank), "wb") as f:
        pickle.dump([(h[0], tuple(h[1]), float(h[2])) for h in got], f)
    dist.destroy_process_group()


def test_sharded_nccl_world2_c4(tmp_path, mtm):
    import pickle
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [pickle.load(open(tmp_path / ("rank%d.pkl" % r), "rb")) for r in range(2)]
    assert res[0] == res[1]
    image, temps, params = synth.config("C4")
    assert_hits_equal(res[0], mtm.matchTemplates(temps, image, **params), tol=0)
