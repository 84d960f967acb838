"""One launch of every kernel family outside the backbone forward, at the BASELINE shapes -- the command `ncu --set full`
wraps for the iou3d / roipool3d / gradient / training captures (profiles/r2_ncu_ops_*.csv):

    config 4   boxes_iou_bev 16384 x 16384, nms_gpu 16384 boxes, roipool3d 16384 boxes x 16384 points (C = 1, 128)
    gradients  three_interpolate_grad (FP0 shape), group_points_grad (SA2 scale 1 shape)
    training   one shared-MLP layer forward + backward at the SA2 scale 1 shape (99 -> 64, 16 x 1024 x 32 columns)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import native, pointnet2_utils, pytorch_utils, synth, train_mlp, workloads  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
sc = workloads.proposals_scene(dev, 16384)
nb = 16384
ans = torch.empty((nb, nb), device=dev)
native.boxes_iou_bev_gpu(sc["bev"], sc["bev"], ans)
sb = sc["bev"][sc["scores"].sort(descending=True)[1]].contiguous()
native.nms_device(sb, 0.85)
for c in (1, 128):
    f = sc["features"][..., :c].contiguous()
    pooled = torch.zeros((1, nb, 512, 3 + c), device=dev)
    flag = torch.zeros((1, nb), dtype=torch.int32, device=dev)
    native.roipool3d_forward(sc["xyz"], sc["boxes3d"], f, pooled, flag)
    del pooled
# gradient kernels
B = 16
pts = torch.from_numpy(synth.make_batch(B, 16384)).to(dev)
xyz = pts[..., :3].contiguous()
_, known = pointnet2_utils.sample_and_gather(xyz, 4096)
idx, w = pointnet2_utils.three_nn_weights(xyz, known)
feats = torch.randn(B, 128, 4096, device=dev, requires_grad=True)
pointnet2_utils.three_interpolate(feats, idx, w).sum().backward()
_, c2 = pointnet2_utils.sample_and_gather(known, 1024)
bq = pointnet2_utils.ball_query(1.0, 32, known, c2)
f2 = torch.randn(B, 96, 4096, device=dev, requires_grad=True)
pointnet2_utils.group_concat(known, c2, f2, bq, True).sum().backward()
# one training layer
mlp = pytorch_utils.SharedMLP([99, 64], bn=True).to(dev).train()
x = torch.randn(B, 99, 1024 * 32, device=dev, requires_grad=True)
train_mlp.shared_mlp_train(mlp, x, pool=32).sum().backward()
torch.cuda.synchronize()
print("ok")
