"""Template-sharded ``matchTemplates`` across GPUs (one process per GPU, torch.distributed).

The reference parallelises over templates with a thread pool
(``MTM/__init__.py:172-175``) and couples them again only in the NMS
(``MTM/__init__.py:294-296``).  The same cut works across GPUs:

* every rank holds the whole image and template list (KBs..MBs) and searches a
  CONTIGUOUS slice of the template list on its GPU (slices balanced by multiply-accumulate
  count, so that a list of mixed template sizes -- BASELINE.json configs[4] -- loads the
  ranks evenly) -> pre-NMS hits in the canonical order (template index, then the peak
  finder's order);
* one all-reduce(MAX) of the hit counts and ONE all-gather of fixed-size hit rows
  (6 x int32 per hit; NCCL over NVLink for CUDA tensors, gloo on CPU);
* concatenation in rank order IS the canonical global order, so every rank runs the
  identical global NMS and returns the identical list.

``matchTemplatesBatchSharded`` is the other cut of SURVEY.md 8e (configs[4]: a batch of
images): contiguous slices of the IMAGE list per rank, the whole template list everywhere,
NMS stays local to the image, and the same single all-gather returns every image's final
hit list to every rank.

``find_fn`` / ``nms_fn`` / ``batch_fn`` are injectable so the host logic is testable on CPU
with world_size 2 (tests/test_sharded_gloo.py uses the oracle there; the product default is
the CUDA path).
"""
import numpy as np

_INF = float("inf")


def shard_bounds(n_items, world, rank):
    """Contiguous, balanced [start, stop) slice of ``n_items`` for ``rank``."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


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


def pack_hits(hits, first_index):
    """(label '#k', bbox, score) hits -> int32 rows [global index, x, y, w, h, score bits]."""
    rows = np.zeros((len(hits), 6), np.int32)
    for i, (label, box, score) in enumerate(hits):
        rows[i, 0] = first_index + int(label[1:])
        rows[i, 1:5] = box
        rows[i, 5] = np.array([score], np.float32).view(np.int32)[0]
    return rows


def unpack_hits(rows, listTemplates):
    return [(listTemplates[int(r[0])][0], (int(r[1]), int(r[2]), int(r[3]), int(r[4])),
             np.array([r[5]], np.int32).view(np.float32)[0]) for r in rows]


def gather_rows(rows, group=None, device=None):
    """All ranks contribute ``rows`` (k_r x ncol int32, same ncol everywhere); returns the rank-ordered concatenation."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", device if device is not None else torch.cuda.current_device())
    else:
        dev = torch.device("cpu")
    cap = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
    cap = max(int(cap.item()), 1)
    ncol = rows.shape[1]
    buf = torch.zeros((cap + 1, ncol), dtype=torch.int32, device=dev)     # row 0 = header (count)
    buf[0, 0] = rows.shape[0]
    if rows.shape[0]:
        buf[1:1 + rows.shape[0]] = torch.from_numpy(rows).to(dev)
    gathered = torch.empty((world * (cap + 1), ncol), dtype=torch.int32, device=dev)   # concatenated form (gloo + nccl)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    g = gathered.cpu().numpy().reshape(world, cap + 1, ncol)
    return np.concatenate([g[r, 1:1 + int(g[r, 0, 0])] for r in range(world)], axis=0)


def matchTemplatesSharded(listTemplates, image, method=5, N_object=_INF, score_threshold=0.5, maxOverlap=0.25,
                          searchBox=None, *, group=None, device=None, find_fn=None, nms_fn=None):
    """Same contract as ``MTM.matchTemplates``; collective over ``group`` (default: WORLD)."""
    import torch.distributed as dist
    if find_fn is None or nms_fn is None:
        from . import api
        find_fn = find_fn or api.findMatches
        nms_fn = nms_fn or api.NMS
    if maxOverlap < 0 or maxOverlap > 1:
        raise ValueError("Maximal overlap between bounding box is in range [0-1]")
    from .api import _validate_search
    _validate_search(listTemplates, image, N_object, searchBox)          # the reference's errors, original labels
    distributed = dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    lo, hi = weighted_bounds(template_macs(listTemplates, image.shape, searchBox), world)[rank]
    # unique per-shard labels '#k' carry the template index through the label-only hit tuples
    mine = [("#%d" % k,) + tuple(entry[1:]) for k, entry in enumerate(listTemplates[lo:hi])]
    local = find_fn(mine, image, method, N_object, score_threshold, searchBox) if mine else []
    if method == 0:
        raise ValueError("The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.")
    rows = gather_rows(pack_hits(local, lo), group=group, device=device)
    return nms_fn(unpack_hits(rows, listTemplates), score_threshold, method == 1, N_object, maxOverlap)


def matchTemplatesBatchSharded(listTemplates, images, method=5, N_object=_INF, score_threshold=0.5, maxOverlap=0.25,
                               searchBox=None, *, group=None, device=None, batch_fn=None):
    """``[matchTemplates(listTemplates, im, ...) for im in images]`` with the IMAGES sharded over the ranks of ``group``.

    Rank r searches the contiguous slice ``shard_bounds(len(images), world, r)`` of the image list with the whole
    template list (``MTM.matchTemplatesBatch`` on its GPU: NMS is local to an image, MTM/__init__.py:296), then ONE
    all-gather of 7 x int32 rows [image, template, x, y, w, h, score bits] hands every image's final hit list to
    every rank.  Returns the same list of hit lists on every rank.
    """
    import torch.distributed as dist
    images = list(images)
    if batch_fn is None:
        from . import api
        batch_fn = api.matchTemplatesBatch
    if maxOverlap < 0 or maxOverlap > 1:
        raise ValueError("Maximal overlap between bounding box is in range [0-1]")
    from .api import _validate_search
    for im in images:                               # every rank raises the reference's errors BEFORE the collective
        _validate_search(listTemplates, im, N_object, searchBox)
    if method == 0:
        raise ValueError("The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.")
    distributed = dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    lo, hi = shard_bounds(len(images), world, rank)
    relabelled = [("#%d" % k,) + tuple(entry[1:]) for k, entry in enumerate(listTemplates)]
    local = batch_fn(relabelled, images[lo:hi], method, N_object, score_threshold, maxOverlap, searchBox) if hi > lo else []
    n_rows = sum(len(hits) for hits in local)
    rows = np.zeros((n_rows, 7), np.int32)
    k = 0
    for i, hits in enumerate(local):
        for label, box, score in hits:              # the order inside an image is the NMS order: kept as is
            rows[k, 0] = lo + i
            rows[k, 1] = int(label[1:])
            rows[k, 2:6] = box
            rows[k, 6] = np.array([score], np.float32).view(np.int32)[0]
            k += 1
    rows = gather_rows(rows, group=group, device=device)
    results = [[] for _ in images]
    for r in rows:
        results[int(r[0])].append((listTemplates[int(r[1])][0], (int(r[2]), int(r[3]), int(r[4]), int(r[5])),
                                   np.array([r[6]], np.int32).view(np.float32)[0]))
    return results
