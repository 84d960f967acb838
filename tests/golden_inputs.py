"""Seeded inputs shared by tools/make_golden_train.py (which runs the reference's Python on them) and the tests that
compare against the resulting fixture: the fixture then only has to store OUTPUTS."""
import numpy as np

SUBSAMPLE_CASES = (("down", 21000, 16384, 3), ("up", 6000, 16384, 4), ("equal", 4096, 4096, 5), ("nofar", 5000, 4096, 6))


def subsample_inputs(tag, n, seed):
    r = np.random.default_rng(seed)
    pts_rect = r.normal(0, 20, (n, 3)).astype(np.float32)
    depth = np.abs(r.normal(30, 15, n)).astype(np.float32)
    if tag == "nofar":
        depth[:] = 10.0
    inten = r.random(n).astype(np.float32)
    return pts_rect, depth, inten


def rpn_loss_inputs(B=2, N=4096):
    rng = np.random.default_rng(98)
    rpn_cls = rng.normal(0, 2, (B, N, 1)).astype(np.float32)
    rpn_reg = rng.normal(0, 1, (B, N, 40)).astype(np.float32)
    label = np.where(rng.random((B, N)) < 0.08, rng.random((B, N)), 0.0).astype(np.float32)
    reg_label = (rng.normal(0, 2.5, (B, N, 3)) * (label[..., None] > 0)).astype(np.float32)
    return rpn_cls, rpn_reg, label, reg_label
