"""Golden vectors for the SURVEY.md section-8 "next" rows f2 / f3, produced IN THE AUTHORING CONTAINER by executing the
reference's own Python (read from /root/reference at generation time; nothing is copied into this repo):

  f2  lib/utils/iou3d/iou3d_utils.py::boxes_iou3d_gpu, imported unmodified.  Its only native call,
      iou3d_cuda.boxes_overlap_bev_gpu, is served by the CPU oracle (that kernel is pinned separately by
      tests/golden/iou3d.npz, which came from the reference's real kernel on a B200), and torch.cuda.FloatTensor is
      aliased to the CPU type because this container has no GPU.  The fixture keeps the diagonal, which is what
      lib/net/train_functions.py:258-260 keeps.
  f3  lib/utils/distance.py::distance_2 executed from its source text (its default argument calls .cuda() at import
      time, which is stripped), driven by the radius-NMS loop and the cylinder membership test of
      tools/eval_auto.py:272-279,:289-291,:327-343 (script-level code there, transcribed below line by line).

  f4  lib/datasets/kitti_rcnn_dataset.py::KittiRCNNDataset.generate_gaussian_training_labels: a static method of a
      module that cannot be imported here (easydict, matplotlib, cv2, a .cuda() call at import time), so its source
      text is cut out of the file and executed as it stands, with `cfg` = the three constants of lib/config.py:45-47.

    python tools/make_golden_next.py      ->  tests/golden/next_rows.npz
"""
import os
import re
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from ws3d_b200 import synth  # noqa: E402


def load_reference_iou3d_utils():
    fake = types.ModuleType("iou3d_cuda")

    def boxes_overlap_bev_gpu(a, b, out):
        out.copy_(torch.from_numpy(oracle.boxes_overlap_bev(a.numpy(), b.numpy())))
        return 1

    fake.boxes_overlap_bev_gpu = boxes_overlap_bev_gpu
    sys.modules["iou3d_cuda"] = fake
    sys.path.insert(0, REF)
    torch.cuda.FloatTensor = torch.FloatTensor   # no GPU here; same dtype, same arithmetic
    import lib.utils.iou3d.iou3d_utils as ref_iou
    return ref_iou


def load_reference_distance_2():
    src = open(os.path.join(REF, "lib/utils/distance.py")).read()
    line = [l for l in src.splitlines() if l.startswith("def distance_2(")][0]
    ns = {"torch": torch}
    exec(line.replace(".cuda()", ""), ns)
    return ns["distance_2"]


def load_reference_gaussian_labels():
    import math
    import textwrap
    from scipy.stats import multivariate_normal
    src = open(os.path.join(REF, "lib/datasets/kitti_rcnn_dataset.py")).read()
    a = src.index("    def generate_gaussian_training_labels(")
    b = src.index("    def generate_rpn_training_labels(")
    body = textwrap.dedent(src[a:b])
    cfg_src = open(os.path.join(REF, "lib/config.py")).read()
    vals = {k: float(re.search(r"__C\.RPN\.%s = ([0-9.]+)" % k, cfg_src).group(1)) for k in ("GAUSS_HEIGHT", "GAUSS_STATUS", "GAUSS_COV")}
    cfg = types.SimpleNamespace(RPN=types.SimpleNamespace(**vals))
    ns = {"np": np, "math": math, "multivariate_normal": multivariate_normal, "cfg": cfg}
    exec(body, ns)
    return ns["generate_gaussian_training_labels"], vals


def main():
    out = {}
    rng = np.random.default_rng(77)
    # ---- f2
    ref_iou = load_reference_iou3d_utils()
    pts = synth.make_scene(5)[:, :3]
    a = synth.make_boxes(pts, 600, seed=11)
    b = a.copy()
    b[:, [0, 2]] += rng.normal(0, 0.4, (600, 2)).astype(np.float32)      # predicted vs ground truth: mostly overlapping
    b[:, 1] += rng.normal(0, 0.2, 600).astype(np.float32)
    b[:, 3:6] *= (1 + rng.normal(0, 0.1, (600, 3))).astype(np.float32)
    b[:, 6] += rng.normal(0, 0.3, 600).astype(np.float32)
    b[:40] = a[:40]                                                      # identical boxes
    b[40:80, 0] += 30                                                    # far apart
    b[80:100, 1] -= 5                                                    # BEV overlap, no height overlap
    a[100:110, 6] = 0.0; b[100:110, 6] = 0.0                             # axis aligned
    iou2d, iou3d = ref_iou.boxes_iou3d_gpu(torch.from_numpy(a), torch.from_numpy(b))
    out["f2_boxes_a"], out["f2_boxes_b"] = a, b
    out["f2_iou2d_diag"] = torch.diagonal(iou2d).numpy().copy()
    out["f2_iou3d_diag"] = torch.diagonal(iou3d).numpy().copy()
    # ---- f3
    distance_2 = load_reference_distance_2()
    scene = synth.make_scene(9)
    inputs = torch.from_numpy(scene.copy())
    n_prop = 900
    seeds = rng.integers(0, scene.shape[0], 200)
    pick = seeds[rng.integers(0, 200, n_prop)]                          # clumps: most candidates have a close neighbour
    centres = scene[pick][:, [0, 2]] + rng.normal(0, 0.25, (n_prop, 2)).astype(np.float32)
    centres[50:60] = centres[40:50]                                      # exact duplicates
    centres[60:70] = centres[40:50] + np.float32(0.3) * np.array([[1, 0]], dtype=np.float32)  # at the threshold
    scores = rng.uniform(0, 1, n_prop).astype(np.float32)
    rpn_rois = torch.from_numpy(np.stack([centres[:, 0], np.zeros(n_prop, np.float32), centres[:, 1]], 1))
    rpn_scores_raw = torch.from_numpy(scores)
    # eval_auto.py:266-279
    sort_points = torch.argsort(-rpn_scores_raw)
    rpn_rois = rpn_rois[sort_points]
    keep_id = [0]
    prop_prop_distance = distance_2(rpn_rois[:, [0, 2]], rpn_rois[:, [0, 2]])
    for i in range(1, rpn_rois.shape[0]):
        if torch.min(prop_prop_distance[keep_id, i], dim=-1)[0] > 0.3:
            keep_id.append(i)
    rpn_center = rpn_rois[keep_id][:, [0, 2]]
    out["f3_centres"], out["f3_scores"] = centres, scores
    out["f3_sort"] = sort_points.numpy()
    out["f3_keep_id"] = np.asarray(keep_id, dtype=np.int64)
    # eval_auto.py:289-291 and :327-343
    point_center_distance = distance_2(rpn_center, inputs[:, [0, 2]])
    cur_proposal_points_index = (torch.min(point_center_distance, dim=-1)[0] < 4.0)
    out["f3_points"] = scene[:, :3].copy()
    out["f3_any"] = cur_proposal_points_index.numpy().astype(np.uint8)
    member = (point_center_distance < 4.0).numpy()                       # (points, centres); column c = :336
    out["f3_cnt"] = member.sum(0).astype(np.int32)
    cap = 1024
    idx = np.full((member.shape[1], cap), -1, dtype=np.int32)
    for c in range(member.shape[1]):
        w = np.nonzero(member[:, c])[0][:cap]
        idx[c, :len(w)] = w
    out["f3_idx"] = idx
    # ---- f4
    labels_fn, cfg_vals = load_reference_gaussian_labels()
    assert cfg_vals == {"GAUSS_HEIGHT": 0.707, "GAUSS_STATUS": 0.7, "GAUSS_COV": 1.5}, cfg_vals
    scene4 = synth.make_scene(12)[:, :3].copy()
    gt = synth.make_boxes(scene4, 14, seed=21)
    gt[3] = gt[2]                                    # duplicate box: argmin must take the first
    cls4, reg4 = labels_fn(scene4, gt)
    out["f4_points"], out["f4_boxes"] = scene4, gt
    out["f4_cls"], out["f4_reg"] = np.asarray(cls4, dtype=np.float64), np.asarray(reg4, dtype=np.float32)
    cls0, reg0 = labels_fn(scene4[:100], gt[:0])     # a scene without boxes
    out["f4_cls_empty"], out["f4_reg_empty"] = np.asarray(cls0, dtype=np.float64), np.asarray(reg0, dtype=np.float32)
    path = os.path.join(ROOT, "tests", "golden", "next_rows.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()}, "kept", len(keep_id), "max cnt", int(out["f3_cnt"].max()))


if __name__ == "__main__":
    main()
