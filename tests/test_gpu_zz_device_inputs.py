"""-m gpu: images that already live in HBM (objects exposing __cuda_array_interface__, here PyTorch CUDA tensors) go
through the public API without touching the host: device-to-device copy into the tile layout (mtm_set_image_device),
searchBox crops as pointer arithmetic.  Results must equal those of the same pixels passed as numpy arrays."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cuda_tensors_through_the_public_api(mtm):
    torch = pytest.importorskip("torch")
    from mtm_b200 import _native
    from oracle import synth
    rng = np.random.default_rng(43)
    temps = [("a", synth.make_template(rng, 24, 40)), ("b", synth.make_template(rng, 32, 32))]
    img, _ = synth.make_scene(300, 420, [t[1] for t in temps], 3, seed=43)
    cases = [(img, temps),
             (img.astype(np.float32), [(n, t.astype(np.float32)) for n, t in temps]),
             (img.astype(np.uint16) * 100, [(n, t.astype(np.uint16) * 100) for n, t in temps]),
             (np.stack([img, 255 - img, img[::-1]], axis=2), [(n, np.ascontiguousarray(np.stack([t, 255 - t, t[::-1]], axis=2))) for n, t in temps])]
    ctx = _native.default_context()
    for image, ts in cases:
        image = np.ascontiguousarray(image)
        dev = torch.from_numpy(image).cuda()               # uint16 included: only storage is needed, no torch arithmetic
        torch.cuda.synchronize()
        for kw in (dict(score_threshold=0.5, maxOverlap=0.25), dict(N_object=1), dict(score_threshold=0.5, searchBox=(13, 7, 350, 260))):
            before = ctx.counters()["h2d_bytes"]
            got = mtm.matchTemplates(ts, dev, **kw)
            moved = ctx.counters()["h2d_bytes"] - before
            assert moved < image.nbytes // 4, "the image must not cross PCIe again (%d bytes moved)" % moved
            want = mtm.matchTemplates(ts, image, **kw)
            assert [(h[0], h[1], float(h[2])) for h in got] == [(h[0], h[1], float(h[2])) for h in want] and len(want) >= 1
        assert np.array_equal(mtm.computeScoreMap(ts[0][1], dev), mtm.computeScoreMap(ts[0][1], image))
    dev = torch.from_numpy(img).cuda()
    torch.cuda.synchronize()
    batch = mtm.matchTemplatesBatch(temps, [dev, img, dev[10:250, 20:400]], score_threshold=0.5)
    want = [mtm.matchTemplates(temps, im, score_threshold=0.5) for im in (img, img, img[10:250, 20:400])]
    assert [[(h[0], h[1], float(h[2])) for h in hits] for hits in batch] == [[(h[0], h[1], float(h[2])) for h in hits] for hits in want]
    with pytest.raises(NotImplementedError, match="device-resident"):
        mtm.matchTemplates([("a", temps[0][1].astype(np.float32))], dev)
