"""Synthetic KITTI-shaped inputs: there is no dataset on the box.

A scene = 16384 points in the rect-camera frame (x in [-40,40], z in [0,70.4], y down): 70 % ground
points around y = 1.6 m, the rest inside 20-40 car-sized clusters; one intensity channel in
[-0.5, 0.5]; the points are shuffled like the reference loader does (kitti_rcnn_dataset.py:433).
Deterministic in (seed, scene_id); numpy only, so the CPU oracle and the GPU see identical bits.
"""
import numpy as np

CAR_SIZE = (1.52563191462, 1.62856739989, 3.88311640418)  # h, w, l (weaklyRPN.yaml:19)


def make_scene(scene_id: int = 0, num_points: int = 16384, seed: int = 1234, uniform: bool = False,
               return_boxes: bool = False):
    """-> (num_points, 4) float32 [x, y, z, intensity]; with `return_boxes` also the (n_cars, 7) ground-truth boxes
    [x, y(bottom), z, h, w, l, ry] of the car clusters (empty for the uniform variant)."""
    rng = np.random.default_rng(seed + scene_id)
    boxes = np.zeros((0, 7), np.float32)
    if uniform:
        xyz = np.stack([rng.uniform(-40, 40, num_points), rng.uniform(-1, 3, num_points),
                        rng.uniform(0, 70.4, num_points)], axis=1)
    else:
        n_ground = int(num_points * 0.7)
        ground = np.stack([rng.uniform(-40, 40, n_ground),
                           np.clip(1.6 + rng.normal(0, 0.1, n_ground), -3, 3),
                           rng.uniform(0, 70.4, n_ground)], axis=1)
        n_obj = num_points - n_ground
        n_cars = int(rng.integers(20, 41))
        centres = np.stack([rng.uniform(-35, 35, n_cars), np.full(n_cars, 0.9), rng.uniform(5, 65, n_cars)], axis=1)
        ry = rng.uniform(-np.pi, np.pi, n_cars)
        owner = rng.integers(0, n_cars, n_obj)
        local = np.stack([rng.uniform(-CAR_SIZE[2] / 2, CAR_SIZE[2] / 2, n_obj),
                          rng.uniform(-1.0, 0.7, n_obj),
                          rng.uniform(-CAR_SIZE[1] / 2, CAR_SIZE[1] / 2, n_obj)], axis=1)
        c, s = np.cos(ry[owner]), np.sin(ry[owner])
        obj = np.stack([local[:, 0] * c + local[:, 2] * s, local[:, 1], -local[:, 0] * s + local[:, 2] * c], axis=1)
        xyz = np.concatenate([ground, obj + centres[owner]], axis=0)
        # the clusters fill y in [centre - 1.0, centre + 0.7] (y points down): box bottom at centre + 0.7
        boxes = np.concatenate([centres[:, 0:1], centres[:, 1:2] + 0.7, centres[:, 2:3],
                                np.tile(np.asarray(CAR_SIZE)[None, :], (n_cars, 1)), ry[:, None]], axis=1).astype(np.float32)
    intensity = rng.uniform(0, 1, num_points) - 0.5
    pts = np.concatenate([xyz, intensity[:, None]], axis=1).astype(np.float32)
    pts = pts[rng.permutation(num_points)]
    return (pts, boxes) if return_boxes else pts


def make_gt_boxes(batch: int, num_points: int = 16384, first_scene: int = 0, seed: int = 1234, pad_to: int = 48):
    """Ground-truth boxes of scenes [first_scene, first_scene + batch): (batch, pad_to, 7) float32 zero padded and the
    per-scene counts (batch,) int32 -- the `gt_boxes3d` a loader would hand to the label generator."""
    out = np.zeros((batch, pad_to, 7), np.float32)
    cnt = np.zeros(batch, np.int32)
    for i in range(batch):
        _, bx = make_scene(first_scene + i, num_points, seed, return_boxes=True)   # (the RNG stream depends on num_points)
        cnt[i] = min(len(bx), pad_to)
        out[i, :cnt[i]] = bx[:cnt[i]]
    return out, cnt


def make_batch(batch: int, num_points: int = 16384, seed: int = 1234, first_scene: int = 0, uniform: bool = False):
    return np.stack([make_scene(first_scene + i, num_points, seed, uniform) for i in range(batch)], axis=0)


def make_boxes(points_xyz: np.ndarray, num_boxes: int, seed: int = 4321) -> np.ndarray:
    """Proposals for one scene: (num_boxes, 7) float32 [x, y(bottom), z, h, w, l, ry] with centres drawn
    from the scene's points (+N(0,0.3)), car-sized, random heading."""
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, points_xyz.shape[0], num_boxes)
    ctr = points_xyz[pick, :3] + rng.normal(0, 0.3, (num_boxes, 3))
    size = np.asarray(CAR_SIZE)[None, :] + rng.normal(0, 0.1, (num_boxes, 3))
    ry = rng.uniform(-np.pi, np.pi, num_boxes)
    boxes = np.concatenate([ctr[:, 0:1], ctr[:, 1:2] + size[:, 0:1] / 2, ctr[:, 2:3], size, ry[:, None]], axis=1)
    return boxes.astype(np.float32)


def boxes3d_to_bev(boxes3d: np.ndarray) -> np.ndarray:
    """numpy twin of kitti_utils.boxes3d_to_bev_torch."""
    cu, cv = boxes3d[:, 0], boxes3d[:, 2]
    hl, hw = boxes3d[:, 5] / 2, boxes3d[:, 4] / 2
    return np.stack([cu - hl, cv - hw, cu + hl, cv + hw, boxes3d[:, 6]], axis=1).astype(np.float32)
