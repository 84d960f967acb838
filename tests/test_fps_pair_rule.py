"""The pairing rule of the throughput sampler (csrc/fps_smem.cu: fps_pair_kernel), restated in numpy and checked against plain
furthest point sampling on the CPU: two samples are taken per traversal whenever the second one is provably the next, and the
sample order must not change.  Same distance function for the updates, the plain sampler and the rule's own test, as in the kernel
(the proof needs nothing else: monotone rounding for the box-diagonal bound, which float32 numpy arithmetic has)."""
import numpy as np
import pytest


def _d2(p, c):
    d = (p - c).astype(np.float32)
    t = d[..., 1] * d[..., 1]
    t = d[..., 0] * d[..., 0] + t
    return (d[..., 2] * d[..., 2] + t).astype(np.float32)


def _plain_fps(pts, m):
    n = len(pts)
    t = np.full(n, 1e10, np.float32)
    idx = [0]
    with np.errstate(invalid="ignore", over="ignore"):
        for _ in range(m - 1):
            t = np.fmin(t, _d2(pts, pts[idx[-1]]))        # fmin: a NaN distance leaves the running distance alone
            idx.append(int(np.flatnonzero(t == t.max())[0]))   # ties: the smallest index
    return idx, t


def _paired_fps(pts, m, bucket=32):
    """Buckets of `bucket` consecutive points (the kernel's are Morton-sorted; the rule does not care), per bucket the maximum,
    its index, the second largest value and the box diagonal; A = best bucket maximum, B = best maximum of the other buckets."""
    n = len(pts)
    nb = (n + bucket - 1) // bucket
    t = np.full(n, 1e10, np.float32)
    fin = np.isfinite(pts).all(1)
    diag = np.full(nb, np.inf, np.float32)
    for b in range(nb):
        sl = slice(b * bucket, min(n, (b + 1) * bucket))
        if fin[sl].all():
            ext = (pts[sl].max(0) - pts[sl].min(0)).astype(np.float32)
            diag[b] = _d2(ext[None], np.zeros(3, np.float32))[0] * np.float32(1.0001)
    idx, pending, pairs = [0], [0], 0
    with np.errstate(invalid="ignore", over="ignore"):
        while len(idx) < m:
            for c in pending:
                t = np.fmin(t, _d2(pts, pts[c]))
            best = []                                      # (value, index) of every bucket's maximum
            for b in range(nb):
                sl = slice(b * bucket, min(n, (b + 1) * bucket))
                v = t[sl]
                best.append((float(v.max()), b * bucket + int(np.flatnonzero(v == v.max())[0])))
            order = sorted(range(nb), key=lambda b: (-best[b][0], best[b][1]))
            a = best[order[0]][1]
            idx.append(a)
            pending = [a]
            if nb > 1 and len(idx) + 1 < m:                # B must not be the last sample
                tb, b_i = best[order[1]]
                sl = slice(order[0] * bucket, min(n, (order[0] + 1) * bucket))
                rest = np.delete(t[sl], a - order[0] * bucket)
                t2 = float(rest.max()) if len(rest) else -1.0
                d_ab = float(_d2(pts[b_i][None], pts[a])[0])
                if (tb > 0 and not (d_ab < tb) and min(t2, float(diag[order[0]])) < tb and fin[a] and fin[b_i]):
                    idx.append(b_i)
                    pending.append(b_i)
                    pairs += 1
    return idx, pairs


def _clouds():
    rng = np.random.default_rng(7)
    yield "uniform", rng.uniform(-10, 10, (700, 3)).astype(np.float32), 300
    yield "lattice", rng.integers(-3, 4, (600, 3)).astype(np.float32), 400          # heavy ties and duplicates
    yield "m>n", rng.integers(0, 2, (80, 3)).astype(np.float32), 120                 # every distance reaches 0
    line = rng.uniform(-5, 5, (500, 3)).astype(np.float32)
    line[100:400, 1:] = 0.0
    line[400:450] = line[399]
    line[470] = [1e6, -1e6, 1e6]
    yield "line+outlier", line, 200
    odd = rng.uniform(-5, 5, (400, 3)).astype(np.float32)
    odd[37, 0] = np.nan
    odd[251, 2] = np.inf
    yield "non-finite", odd, 60
    yield "sorted", np.sort(rng.uniform(0, 50, (640, 3)).astype(np.float32), axis=0), 320   # compact buckets: most samples pair


@pytest.mark.parametrize("name,pts,m", list(_clouds()), ids=[c[0] for c in _clouds()])
def test_paired_sampling_keeps_the_sample_order(name, pts, m):
    want, _ = _plain_fps(pts, m)
    got, pairs = _paired_fps(pts, m)
    assert got == want, f"{name}: first difference at sample {next(i for i, (g, w) in enumerate(zip(got, want)) if g != w)}"
    if name in ("uniform", "sorted"):
        assert pairs > m // 8, f"{name}: the rule should fire often ({pairs} pairs for {m} samples)"
