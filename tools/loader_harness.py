"""Stage-1 data path harness: a loader with num_workers > 0 whose batches are finished on the
GPU, against the reference's per-sample path, on synthetic raw scenes.

The reference's KittiRCNNDataset.__getitem__ (lib/datasets/kitti_rcnn_dataset.py) calls CUDA from inside the dataset --
furthest_point_sample on every pasted ground-truth object followed by `.cpu()` (:305-313) -- which forces num_workers = 0
and one device synchronisation per object; it then subsamples to 16384 points (:424-452) and builds the Gaussian labels
(:529-573) in numpy.  Two arms over the same samples and the same random draws:

  reference-style   per sample, in the main process: per-object FPS + .cpu(), numpy paste / subsample / labels, H2D
  device            DataLoader workers (num_workers >= 2) only produce the raw arrays; per BATCH on the GPU: all objects'
                    FPS queued back to back (data_utils.sample_objects), paste, ONE host read (the near-point counts the
                    random draws depend on), subsample_points with the host's draws, labels for the batch in one launch

The two arms must produce identical network inputs and labels (float32 label Gaussian: 1e-6).  Prints one JSON line.

    python tools/loader_harness.py [--batches 6] [--batch 8] [--workers 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import data_utils, label_utils, pointnet2_utils, synth  # noqa: E402

NPOINTS, OBJ_POINTS = 16384, 100


class RawScenes(torch.utils.data.Dataset):
    """What a decoding worker hands over: the raw scene (more points than the network takes), its ground-truth boxes
    and the point clouds of the objects to paste.  numpy only: runs in worker processes."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        rng = np.random.default_rng(1000 + i)
        n_raw = int(rng.integers(17000, 22000))
        pts, boxes = synth.make_scene(i, n_raw, return_boxes=True)
        pts[:, 3] += 0.5                                      # raw intensity in [0, 1]; the subsampling shifts it (:444)
        objs = []
        for _ in range(int(rng.integers(4, 9))):               # pasted objects: 150..700 points inside a car-sized box
            c = np.array([rng.uniform(-30, 30), 0.9, rng.uniform(5, 65)])
            k = int(rng.integers(150, 700))
            objs.append((c + rng.uniform(-1, 1, (k, 3)) * np.array([1.9, 0.8, 0.8])).astype(np.float32))
        return {"pts": pts, "boxes": boxes, "objs": objs, "seed": 77 + i}


def collate(samples):
    return samples


def numpy_labels(xyz, boxes):
    """generate_gaussian_training_labels (kitti_rcnn_dataset.py:529-573) in vectorised float32 numpy."""
    d = np.sqrt((xyz[:, None, 0] - boxes[None, :, 0]) ** 2 + (xyz[:, None, 1] * np.float32(0.707)) ** 2 + (xyz[:, None, 2] - boxes[None, :, 2]) ** 2)
    centre = np.minimum(np.float32(100.0), np.clip(d - np.float32(0.7), 0, 100).min(axis=1))
    cls = np.exp(-0.5 * centre.astype(np.float64) ** 2 / 1.5).astype(np.float32)
    tgt = d.argmin(axis=1)
    reg = np.zeros((xyz.shape[0], 3), np.float32)
    fg = d.min(axis=1) < 4.0
    reg[fg, 0] = boxes[tgt[fg], 0] - xyz[fg, 0]
    reg[fg, 2] = boxes[tgt[fg], 2] - xyz[fg, 2]
    return cls, reg


def reference_style(sample, dev):
    """One sample the way the reference's __getitem__ does it (one sync per object, numpy for the rest)."""
    rs = np.random.RandomState(sample["seed"])
    pts = sample["pts"]
    for obj in sample["objs"]:
        t = torch.from_numpy(obj).to(dev).contiguous().view(1, -1, 3)
        sel = pointnet2_utils.furthest_point_sample(t, OBJ_POINTS).cpu().numpy().reshape(-1)       # :309-312 (synchronises)
        pts = np.concatenate([pts, np.concatenate([obj[sel], np.full((OBJ_POINTS, 1), 0.5, np.float32)], axis=1)], axis=0)
    depth = pts[:, 2]
    n_near = int((depth < 40.0).sum())
    perm, order = data_utils.draw_subsample(rs, pts.shape[0], n_near, NPOINTS)
    near, far = np.where(depth < 40.0)[0], np.where(~(depth < 40.0))[0]
    choice = np.concatenate([near[perm], far])[order]                                             # :424-441
    out = pts[choice].copy()
    out[:, 3] -= np.float32(0.5)                                                                  # :444
    cls, reg = numpy_labels(out[:, :3], sample["boxes"])
    return torch.from_numpy(out).to(dev), torch.from_numpy(cls).to(dev), torch.from_numpy(reg).to(dev)


def device_batch(samples, dev):
    """A whole batch on the GPU; the host reads back B integers once."""
    clouds = []
    for s in samples:
        raw = torch.from_numpy(s["pts"]).to(dev, non_blocking=True)
        objs = [torch.from_numpy(o).to(dev, non_blocking=True) for o in s["objs"]]
        sels = data_utils.sample_objects(objs, OBJ_POINTS)                                        # queued, no host read
        pasted = [torch.cat([o[i.long()], torch.full((OBJ_POINTS, 1), 0.5, device=dev)], dim=1) for o, i in zip(objs, sels)]
        clouds.append(torch.cat([raw] + pasted, dim=0))
    n_near = torch.stack([(c[:, 2] < 40.0).sum() for c in clouds]).cpu().tolist()                 # the ONE synchronisation
    outs = []
    for s, c, nn_ in zip(samples, clouds, n_near):
        perm, order = data_utils.draw_subsample(np.random.RandomState(s["seed"]), c.shape[0], nn_, NPOINTS)
        outs.append(data_utils.subsample_points(c, c[:, 2].contiguous(), NPOINTS, perm, order, nn_)[0])
    batch = torch.stack(outs)
    gmax = max(len(s["boxes"]) for s in samples)
    boxes = torch.zeros((len(samples), gmax, 7), device=dev)
    for k, s in enumerate(samples):
        boxes[k, :len(s["boxes"])] = torch.from_numpy(s["boxes"]).to(dev)
    cnt = torch.tensor([len(s["boxes"]) for s in samples], dtype=torch.int32, device=dev)
    cls, reg = label_utils.generate_gaussian_training_labels(batch[..., :3].contiguous(), boxes, cnt)
    return batch, cls, reg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=6)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--workers", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ds = RawScenes(args.batches * args.batch)
    # ---- parity on the first batch
    first = [ds[i] for i in range(args.batch)]
    b, cls, reg = device_batch(first, dev)
    for k, s in enumerate(first):
        rb, rc, rr = reference_style(s, dev)
        assert torch.equal(b[k], rb), f"sample {k}: network input differs"
        assert torch.equal(reg[k], rr), f"sample {k}: regression labels differ"
        assert float((cls[k] - rc).abs().max()) < 1e-6, f"sample {k}: classification labels differ"
    # ---- reference-style: main process, per-sample
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(len(ds)):
        reference_style(ds[i], dev)
    torch.cuda.synchronize()
    t_ref = time.perf_counter() - t0
    # ---- device path behind a multi-worker loader
    loader = torch.utils.data.DataLoader(ds, batch_size=args.batch, num_workers=args.workers, collate_fn=collate, persistent_workers=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for samples in loader:
        device_batch(samples, dev)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    syncs_ref = sum(len(ds[i]["objs"]) for i in range(len(ds)))
    print(json.dumps({"metric": "Stage-1 data path scenes/sec (synthetic raw scenes -> network input + labels)",
                      "scenes": len(ds), "batch": args.batch,
                      "reference_style": {"scenes_per_s": round(len(ds) / t_ref, 1), "num_workers": 0, "device_syncs": syncs_ref,
                                          "note": "per-object FPS + .cpu(), numpy paste / subsample / labels in the main process"},
                      "device": {"scenes_per_s": round(len(ds) / t_dev, 1), "num_workers": args.workers, "device_syncs": args.batches,
                                 "note": "workers produce raw arrays; FPS / paste / subsample / labels on the GPU, one host read per batch"},
                      "parity": "network inputs and regression labels bit-identical, classification labels within 1e-6 (first batch)"}))


if __name__ == "__main__":
    main()
