"""``matchTemplates`` across GPUs: the two cuts of the reference's parallel axis.

The reference parallelises over templates with a thread pool (``MTM/__init__.py:172-175``) and couples them
again only in the NMS (``MTM/__init__.py:294-296``).  Across GPUs:

* ``matchTemplatesSharded`` -- TEMPLATE cut (BASELINE.json configs[3]): every rank holds the whole image and
  searches a CONTIGUOUS slice of the template list (slices balanced by multiply-accumulate count, so that a list
  of mixed template sizes loads the ranks evenly).  The library then exchanges the ranks' hit blocks with ONE
  ``ncclAllGather`` on the context's stream and runs the identical global NMS on every GPU
  (``mtm_match_templates_sharded``): no host round trip between the local search and the final list.
* ``matchTemplatesBatchSharded`` -- IMAGE cut (configs[4]: a batch of images): contiguous blocks of the image list
  per rank, the whole template list everywhere, NMS local to an image; ONE all-gather of the final per-image hit
  blocks (``mtm_gather_results``) returns every image's list to every rank.

No torch here: the communicator is ``_native.Comm`` (NCCL inside libmtm_b200.so; ``rendezvous.comm_from_env()``
builds it from the torchrun environment).  ``search_fn`` / ``batch_fn`` are injectable so that the host logic
(partition, labels, offsets, validation before the collective) runs on CPU with world_size 2
(tests/test_sharded_gloo.py drives it with the oracle and a gloo exchange).
"""
import numpy as np

from . import _native, api

_INF = float("inf")


def shard_bounds(n_items, world, rank):
    """Contiguous, balanced [start, stop) slice of ``n_items`` for ``rank``."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def block_bounds(n_items, world, rank):
    """Image cut: fixed blocks of ``ceil(n_items / world)`` images per rank (the gather layout), last ranks may idle.
    Returns (start, stop, images_per_rank)."""
    per = max(1, -(-n_items // world))
    start = min(n_items, rank * per)
    return start, min(n_items, start + per), per


def weighted_bounds(weights, world):
    """Contiguous partition of ``weights`` into ``world`` slices minimising the heaviest slice (linear partition,
    dynamic programming; lists are short).  Returns ``world`` [start, stop) pairs; slices may be empty."""
    n = len(weights)
    prefix = [0.0]
    for w in weights:
        prefix.append(prefix[-1] + float(w))
    inf = float("inf")
    # best[k][i]: smallest possible heaviest slice when the first i items go to k slices
    best = [[inf] * (n + 1) for _ in range(world + 1)]
    cut = [[0] * (n + 1) for _ in range(world + 1)]
    best[0][0] = 0.0
    for k in range(1, world + 1):
        for i in range(n + 1):
            for j in range(i + 1):
                if best[k - 1][j] == inf:
                    continue
                cost = max(best[k - 1][j], prefix[i] - prefix[j])
                if cost < best[k][i]:                         # strict: ties keep the earliest cut (deterministic on every rank)
                    best[k][i], cut[k][i] = cost, j
    bounds, i = [], n
    for k in range(world, 0, -1):
        j = cut[k][i]
        bounds.append((j, i))
        i = j
    return bounds[::-1]


def template_macs(listTemplates, image_shape, searchBox=None):
    """Multiply-accumulates of every template's score map: C*h*w*(H-h+1)*(W-w+1)."""
    H, W = image_shape[:2]
    if searchBox is not None:
        W, H = min(W, searchBox[2]), min(H, searchBox[3])
    out = []
    for entry in listTemplates:
        t = entry[1]
        h, w = t.shape[:2]
        c = t.shape[2] if t.ndim == 3 else 1
        out.append(float(c) * h * w * max(H - h + 1, 1) * max(W - w + 1, 1))
    return out


def _device_search(ctx, comm, arrays, masks, img, lo, method, n_dev, score_threshold, maxOverlap):
    """Product path of the template cut: this rank's slice on its GPU, exchange + global NMS inside the library."""
    with ctx.lock:
        ctx.set_image(img)
        if arrays:
            api._upload_templates(ctx, arrays, masks)
        return ctx.match_templates_sharded(comm, lo, len(arrays), method, n_dev, score_threshold, maxOverlap)


def matchTemplatesSharded(listTemplates, image, method=5, N_object=_INF, score_threshold=0.5, maxOverlap=0.25,
                          searchBox=None, *, comm, context=None, search_fn=None):
    """Same contract as ``MTM.matchTemplates``; collective over ``comm`` (every rank calls it with the same arguments
    and receives the same list)."""
    if maxOverlap < 0 or maxOverlap > 1:
        raise ValueError("Maximal overlap between bounding box is in range [0-1]")
    image = _native.as_image(image)
    # every rank raises the reference's errors (original labels) BEFORE any collective
    crop, xOffset, yOffset = api._validate_search(listTemplates, image, N_object, searchBox)
    if method == 0:
        raise ValueError("The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.")
    if len(listTemplates) == 0:
        return []
    finite = N_object != _INF
    nms_threshold = (1 - score_threshold) if method == 1 else score_threshold
    if (finite and N_object < 1) or nms_threshold < 0:
        # rare corners whose behaviour depends on the pre-NMS list length (api.matchTemplates): replicated, no exchange
        return api.matchTemplates(listTemplates, image, method, N_object, score_threshold, maxOverlap, searchBox, context=context)
    names, arrays, img, masks = api._prepare(listTemplates, crop, method)
    lo, hi = weighted_bounds(template_macs(listTemplates, image.shape, searchBox), comm.world)[comm.rank]
    n_dev = int(N_object) if finite else -1
    if search_fn is None:
        ctx = context or _native.default_context(comm.device)
        search_fn = lambda *a: _device_search(ctx, comm, *a)       # noqa: E731
    raw = search_fn(arrays[lo:hi], masks[lo:hi], img, lo, method, n_dev, score_threshold, maxOverlap)
    return api._to_hits(raw, names, xOffset, yOffset)


def _device_batch(ctxs, comm, prepared, method, n_dev, score_threshold, maxOverlap, images_per_rank, hits_per_image):
    """Product path of the image cut: this rank's images through the pipelined entry points of its contexts (streams),
    then ONE all-gather of the final hit blocks.  ``prepared`` = [(arrays, masks, img)] of the local images."""
    depth = _native.MAX_INFLIGHT
    n_streams = len(ctxs)
    assert len(prepared) <= depth * n_streams
    uploaded = [None] * n_streams
    entries = []
    for i, (arrays, masks, img) in enumerate(prepared):
        k = i % n_streams
        c = ctxs[k]
        c.set_image(img)
        sig = (img.dtype, img.ndim, tuple(map(id, arrays)), tuple(map(id, masks)))
        if uploaded[k] != sig:
            api._upload_templates(c, arrays, masks)
            uploaded[k] = sig
        slot = (i // n_streams) % depth
        c.match_templates_async(method, n_dev, score_threshold, maxOverlap, slot)
        entries.append((c, slot))
    return comm.gather_results(entries, images_per_rank, hits_per_image)


def matchTemplatesBatchSharded(listTemplates, images, method=5, N_object=_INF, score_threshold=0.5, maxOverlap=0.25,
                               searchBox=None, *, comm, context=None, streams=2, batch_fn=None):
    """``[matchTemplates(listTemplates, im, ...) for im in images]`` with the IMAGES cut into blocks over the ranks of
    ``comm``.  Rank r searches images ``block_bounds(len(images), world, r)`` with the whole template list
    (``mtm_match_templates_async`` on its GPU: NMS is local to an image, MTM/__init__.py:296), then ONE all-gather of
    fixed-size hit blocks (``mtm_gather_results``) hands every image's final hit list to every rank.  Returns the same
    list of hit lists on every rank."""
    images = [_native.as_image(im) for im in images]
    if maxOverlap < 0 or maxOverlap > 1:
        raise ValueError("Maximal overlap between bounding box is in range [0-1]")
    checked = [api._validate_search(listTemplates, im, N_object, searchBox) for im in images]   # every rank, before the collective
    if method == 0:
        raise ValueError("The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.")
    finite = N_object != _INF
    nms_threshold = (1 - score_threshold) if method == 1 else score_threshold
    if len(listTemplates) == 0 or (finite and N_object < 1) or nms_threshold < 0:
        return api.matchTemplatesBatch(listTemplates, images, method, N_object, score_threshold, maxOverlap, searchBox, context=context)
    n_dev = int(N_object) if finite else -1
    hits_per_image = min(max(n_dev, 1), _native.SLOT_HITS) if finite else _native.SLOT_HITS
    results = [None] * len(images)
    ctxs = None
    if batch_fn is None:
        ctx = context or _native.default_context(comm.device)
        ctxs = [ctx] + _native.helper_contexts(ctx.device, max(1, int(streams)) - 1, owner=context)
    round_images = comm.world * _native.MAX_INFLIGHT * (len(ctxs) if ctxs else 1)     # what one exchange can carry
    names = None
    for first in range(0, len(images), round_images):
        chunk = checked[first:first + round_images]
        lo, hi, per = block_bounds(len(chunk), comm.world, comm.rank)
        prepared = []
        for crop, _xo, _yo in chunk[lo:hi]:
            names, arrays, img, masks = api._prepare(listTemplates, crop, method)
            prepared.append((arrays, masks, img))
        if names is None:                                   # an idle rank still needs the labels
            names = [t[0] for t in listTemplates]
        if batch_fn is not None:
            hits, counts = batch_fn(prepared, method, n_dev, score_threshold, maxOverlap, per, hits_per_image)
        else:
            locks = [c.lock for c in ctxs]
            for lk in locks:
                lk.acquire()
            try:
                hits, counts = _device_batch(ctxs, comm, prepared, method, n_dev, score_threshold, maxOverlap, per, hits_per_image)
            finally:
                for lk in reversed(locks):
                    lk.release()
        for r in range(comm.world):
            for i in range(per):
                g = r * per + i                             # index inside the chunk
                if g >= len(chunk):
                    continue
                cnt = int(counts[r * per + i])
                _crop, xo, yo = chunk[g]
                if cnt >= 0:
                    results[first + g] = api._to_hits(hits[r * per + i, :cnt], names, xo, yo)
    for i, r in enumerate(results):                        # did not fit the fused fast path: replicated synchronous call
        if r is None:
            results[i] = api.matchTemplates(listTemplates, images[i], method, N_object, score_threshold, maxOverlap, searchBox,
                                            context=context)
    return results
