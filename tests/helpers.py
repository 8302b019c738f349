"""Comparison helpers shared by the parity tests."""
import numpy as np


def canon(hits):
    """Order-independent form of a pre-NMS hit list: sorted (label, bbox) with scores."""
    return sorted(((h[0], tuple(int(v) for v in h[1])), float(h[2])) for h in hits)


def assert_hits_equal(got, want, tol=1e-4, ordered=True):
    """Post-NMS lists: identical (label, bbox) sequence, scores within ``tol``."""
    if not ordered:
        got, want = canon(got), canon(want)
        assert [g[0] for g in got] == [w[0] for w in want], (got, want)
        for g, w in zip(got, want):
            assert abs(g[1] - w[1]) <= tol, (g, w)
        return
    g_keys = [(h[0], tuple(int(v) for v in h[1])) for h in got]
    w_keys = [(h[0], tuple(int(v) for v in h[1])) for h in want]
    if g_keys != w_keys:
        # scores closer than tol may legitimately swap places in a descending sort
        assert sorted(g_keys) == sorted(w_keys), (got, want)
        gs = dict(zip(g_keys, [float(h[2]) for h in got]))
        for a, b in zip(g_keys, w_keys):
            assert abs(gs[a] - gs[b]) <= tol, "order differs beyond score tolerance: %r vs %r" % (got, want)
    for g, w in zip(sorted(zip(g_keys, [float(h[2]) for h in got])), sorted(zip(w_keys, [float(h[2]) for h in want]))):
        assert abs(g[1] - w[1]) <= tol, (g, w)


def assert_map_close(got, exact, cv=None, tol=1e-4):
    """|gpu - exact| <= tol*max(1,|exact|); and vs cv2: <= tol + |cv2 - exact|."""
    got = np.asarray(got, np.float64)
    exact = np.asarray(exact, np.float64)
    assert got.shape == exact.shape
    err = np.abs(got - exact)
    bound = tol * np.maximum(1.0, np.abs(exact))
    assert np.all(err <= bound), "max err %g at %s" % (err.max(), np.unravel_index(err.argmax(), err.shape))
    if cv is not None:
        cv = np.asarray(cv, np.float64)
        assert np.all(np.abs(got - cv) <= tol + np.abs(cv - exact) + 1e-12)
