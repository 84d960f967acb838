"""Device-side pieces of the Stage-1 data path.

The reference's loader (lib/datasets/kitti_rcnn_dataset.py) runs with num_workers = 0 because __getitem__ itself calls
a CUDA op -- furthest_point_sample on every pasted ground-truth object, followed by `.cpu()` (:305-313), one device
synchronisation per object -- and then subsamples the scene to 16384 points (:424-452) and builds the Gaussian labels
(:529-573) in numpy.  Here the workers only read / decode (any num_workers), and the batch is finished on the device:

  * `sample_objects`   GT-paste FPS of all objects of a batch back to back on the current stream, ONE host read at the
                       end (B = 1 launches keep the reference's per-object block size, hence its exact tie-breaking);
  * `subsample_points` the 16384-point subsampling with the host's random draws (bit-identical to numpy's result);
  * label_utils.generate_gaussian_training_labels for the whole batch in one launch.
"""
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import native, pointnet2_utils

NEAR_DEPTH = 40.0   # kitti_rcnn_dataset.py:427


def draw_subsample(rng: np.random.RandomState, n: int, n_near: int, npoints: int) -> Tuple[np.ndarray, np.ndarray]:
    """The draws numpy makes in kitti_rcnn_dataset.py:424-441 for a cloud of n points, n_near of them nearer than 40 m:
    (perm, order) for `subsample_points`.  `rng` is consumed exactly as `np.random` is there."""
    if npoints < n:
        k = npoints - (n - n_near)
        if k < 0 or k > n_near:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")   # numpy's own error
        perm = rng.permutation(n_near)[:k]            # np.random.choice(near_idxs, k, replace=False)
    else:
        reps = -(-npoints // n)
        perm = rng.permutation(n * reps)[:npoints]    # np.random.choice(tiled arange, npoints, replace=False)
    order = np.arange(npoints)
    rng.shuffle(order)                                # np.random.shuffle(choice)
    return perm.astype(np.int32), order.astype(np.int32)


def subsample_points(pts: torch.Tensor, depth: torch.Tensor, npoints: int, perm, order, n_near: int,
                     intensity_shift: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """pts (n, 3 + C) CUDA rows [x, y, z, intensity...], depth (n) -> (pts_input (npoints, 3 + C), choice (npoints) int32,
    status (1) int32).  The last channel is shifted by -intensity_shift when C > 0 (kitti_rcnn_dataset.py:444).
    `perm` / `order`: draw_subsample()'s arrays (numpy or int32 tensors; copied to the device without a sync)."""
    dev = pts.device
    perm_t = perm if torch.is_tensor(perm) else torch.from_numpy(np.ascontiguousarray(perm, dtype=np.int32))
    order_t = order if torch.is_tensor(order) else torch.from_numpy(np.ascontiguousarray(order, dtype=np.int32))
    perm_t, order_t = perm_t.to(dev, non_blocking=True), order_t.to(dev, non_blocking=True)
    out = torch.empty((npoints, pts.shape[1]), dtype=torch.float32, device=dev)
    choice = torch.empty(npoints, dtype=torch.int32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    native.subsample_points(pts.contiguous(), depth.contiguous(), npoints, n_near, NEAR_DEPTH, intensity_shift, perm_t, order_t,
                            out, choice, status)
    return out, choice, status


def sample_objects(objects: Sequence[torch.Tensor], npoint: int = 100) -> List[torch.Tensor]:
    """FPS of every pasted ground-truth object (kitti_rcnn_dataset.py:309-311) without the per-object `.cpu()`:
    objects[i] is (n_i, 3) CUDA; returns the (npoint,) int32 index tensors, still on the device.  The launches are
    queued back to back; the caller reads them (or gathers with them) once per batch."""
    out = []
    for pts in objects:
        out.append(pointnet2_utils.furthest_point_sample(pts.contiguous().view(1, -1, 3), npoint).view(-1))
    return out
