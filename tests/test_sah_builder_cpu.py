"""The EID_ACCEL_FAST_TRACE topology builder (csrc/sah_host.cpp) through its host-only tap — no GPU needed.  What the GPU refit + 4-wide collapse
rely on: a full binary tree over a permutation (n - 1 inner nodes, root 0, contiguous position ranges, consistent parent links), identical
whatever the thread count; and what the build is for: a lower SAH cost than a median split, on the kind of scene the benchmark renders.
Degenerate inputs (coinciding centroids, one / two boxes, non-finite boxes) must terminate with a valid tree."""
import numpy as np
import pytest

import eidola_b200 as eid


def _check_tree(t, n):
    order = t["order"]
    assert sorted(order.tolist()) == list(range(n)), "order is not a permutation"
    if n < 2:
        return
    left, right, par, first, last, pleaf = (t[k] for k in ("left", "right", "parentInner", "rangeFirst", "rangeLast", "parentLeaf"))
    assert par[0] == -1 and first[0] == 0 and last[0] == n - 1, "node 0 must be the root over every position"
    seen_inner = np.zeros(n - 1, bool)
    seen_leaf = np.zeros(n, bool)
    stack = [0]
    while stack:
        i = stack.pop()
        assert not seen_inner[i], "inner node %d reached twice" % i
        seen_inner[i] = True

        def rng(c):
            return (~c, ~c) if c < 0 else (first[c], last[c])
        (lf, ll), (rf, rl) = rng(left[i]), rng(right[i])
        assert lf == first[i] and ll + 1 == rf and rl == last[i], "children of node %d do not tile its range" % i
        for c in (left[i], right[i]):
            if c < 0:
                assert pleaf[~c] == i and not seen_leaf[~c]
                seen_leaf[~c] = True
            else:
                assert par[c] == i
                stack.append(c)
    assert seen_inner.all() and seen_leaf.all()


def _sah_cost(t, lo, hi):
    """sum over inner nodes of area(node) (the classic SAH cost of the inner nodes, up to constants)"""
    n = lo.shape[0]
    lo_s, hi_s = lo[t["order"]], hi[t["order"]]
    # bottom-up boxes through an explicit post-order
    nlo = np.zeros((n - 1, 3)); nhi = np.zeros((n - 1, 3))
    done = np.zeros(n - 1, bool)
    stack = [0]
    while stack:
        i = stack[-1]
        kids = [c for c in (t["left"][i], t["right"][i]) if c >= 0 and not done[c]]
        if kids:
            stack.extend(kids)
            continue
        stack.pop()
        boxes = [(lo_s[~c], hi_s[~c]) if c < 0 else (nlo[c], nhi[c]) for c in (t["left"][i], t["right"][i])]
        nlo[i] = np.minimum(boxes[0][0], boxes[1][0]); nhi[i] = np.maximum(boxes[0][1], boxes[1][1])
        done[i] = True
    e = nhi - nlo
    return float((e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0]).sum())


def _height_field(q, seed=3):
    """boxes of a q x q-quad height field over 40 x 40 units + a cloud of small boxes above it (the shape of the C3 scene)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = np.linspace(-20.0, 20.0, q + 1)
    x, z = np.meshgrid(g, g)
    y = 0.5 * (np.sin(0.9 * x) * np.cos(0.7 * z) + 1.0) + 0.05 * rng.random(x.shape)
    p = np.stack([x, y, z], -1)
    a, b, c, d = p[:-1, :-1], p[1:, :-1], p[:-1, 1:], p[1:, 1:]
    tris = np.concatenate([np.stack([a, b, c], 2).reshape(-1, 3, 3), np.stack([c, b, d], 2).reshape(-1, 3, 3)])
    cl = np.stack([rng.uniform(-19, 19, 200), rng.uniform(10, 12, 200), rng.uniform(-19, 19, 200)], -1)
    lo = np.concatenate([tris.min(1), cl - 0.1]).astype(np.float32)
    hi = np.concatenate([tris.max(1), cl + 0.1]).astype(np.float32)
    perm = rng.permutation(lo.shape[0])
    return lo[perm], hi[perm]


def _median_tree_cost(lo, hi):
    """reference point: recursive median split on the widest centroid axis"""
    cen = 0.5 * (lo + hi)
    total = 0.0
    stack = [np.arange(lo.shape[0])]
    while stack:
        idx = stack.pop()
        if idx.size < 2:
            continue
        e = hi[idx].max(0) - lo[idx].min(0)
        total += float(e[0] * e[1] + e[1] * e[2] + e[2] * e[0])
        ax = int(np.argmax(cen[idx].max(0) - cen[idx].min(0)))
        o = idx[np.argsort(cen[idx, ax], kind="stable")]
        stack.append(o[: o.size // 2]); stack.append(o[o.size // 2:])
    return total


@pytest.mark.parametrize("n", [1, 2, 3, 5, 64, 1000])
def test_random_boxes_give_a_valid_tree(n):
    rng = np.random.Generator(np.random.PCG64(n))
    c = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
    h = rng.uniform(0.01, 0.5, (n, 3)).astype(np.float32)
    _check_tree(eid.sah_tree(c - h, c + h, threads=3), n)


def test_height_field_tree_is_valid_deterministic_and_better_than_median():
    lo, hi = _height_field(96)                       # 18 632 boxes: above the builder's parallel threshold (2^14)
    n = lo.shape[0]
    t1 = eid.sah_tree(lo, hi, threads=1)
    _check_tree(t1, n)
    for th in (2, 5, 0):
        t = eid.sah_tree(lo, hi, threads=th)
        for k in t1:
            assert np.array_equal(t1[k], t[k]), "%s differs with %d threads" % (k, th)
    sah, med = _sah_cost(t1, lo, hi), _median_tree_cost(lo, hi)
    assert sah < 0.9 * med, "SAH cost %.1f is not clearly below the median-split cost %.1f" % (sah, med)


def test_large_node_path_matches_single_thread():
    """nodes of >= 2^17 boxes are binned by all threads in chunks: the merged bins must give the single-thread tree"""
    lo, hi = _height_field(280)                      # 157 000 boxes
    t1, t4 = eid.sah_tree(lo, hi, threads=1), eid.sah_tree(lo, hi, threads=4)
    for k in t1:
        assert np.array_equal(t1[k], t4[k]), k
    _check_tree(t4, lo.shape[0])


def test_degenerate_inputs_terminate():
    n = 257
    same = np.zeros((n, 3), np.float32)
    _check_tree(eid.sah_tree(same - 1.0, same + 1.0, threads=2), n)          # every centroid coincides: median splits
    line = np.zeros((n, 3), np.float32); line[:, 0] = np.arange(n) // 8       # many ties on the only useful axis
    _check_tree(eid.sah_tree(line, line + 0.5, threads=2), n)
    bad = np.random.Generator(np.random.PCG64(1)).uniform(-1, 1, (n, 3)).astype(np.float32)
    lo, hi = bad - 0.1, bad + 0.1
    lo[5] = np.nan; hi[7] = np.inf; lo[9] = -np.inf                           # non-finite boxes must not hang or break the structure
    _check_tree(eid.sah_tree(lo, hi, threads=2), n)
    huge = np.full((n, 3), 3.0e38, np.float32)
    _check_tree(eid.sah_tree(-huge, huge, threads=1), n)


def test_null_arguments_are_rejected():
    with pytest.raises(eid.EidolaError):
        eid._check(eid.lib().eid_accel_sah_tap(None, None, 4, 1, None, None, None, None, None, None, None))
