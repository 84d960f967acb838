"""TEST INFRASTRUCTURE ONLY: the PointNet++ backbone forward on the CPU, composed from the CPU
oracle ops (oracle/ws3d_oracle.c) in the reference's sequencing (pointnet2_modules.py:19-55,
127-156; lib/net/pointnet2_msg.py:56-70) with the module's own MLP weights evaluated by PyTorch on
the host.  Used by the tests as the checker for the GPU forward and by bench.py as the CPU baseline
(the reference has no CPU implementation of these ops)."""
import numpy as np
import torch
import torch.nn.functional as F

import oracle


def sa_forward(sa, xyz, feats):
    idx = oracle.furthest_point_sample(xyz, sa.npoint)
    new_xyz = np.take_along_axis(xyz, idx[..., None].astype(np.int64), 1)
    outs = []
    for grouper, mlp in zip(sa.groupers, sa.mlps):
        g = oracle.query_and_group(grouper.radius, grouper.nsample, xyz, new_xyz, feats, grouper.use_xyz)
        with torch.no_grad():
            y = mlp(torch.from_numpy(g))
            y = F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1)
        outs.append(y.numpy())
    return new_xyz, np.concatenate(outs, axis=1)


def fp_forward(fp, unknown, known, unknown_feats, known_feats):
    d2, idx = oracle.three_nn(unknown, known)
    with torch.no_grad():
        dist = torch.sqrt(torch.from_numpy(d2))
        recip = 1.0 / (dist + 1e-8)
        weight = (recip / torch.sum(recip, dim=2, keepdim=True)).numpy()
    interp = oracle.three_interpolate(known_feats, idx, weight)
    x = interp if unknown_feats is None else np.concatenate([interp, unknown_feats], axis=1)
    with torch.no_grad():
        return fp.mlp(torch.from_numpy(x).unsqueeze(-1)).squeeze(-1).numpy()


def backbone_forward(model, pts):
    """model: ws3d_b200.models.Pointnet2MSG on the CPU in eval mode; pts (B,N,3+C) float32 numpy."""
    xyz = np.ascontiguousarray(pts[..., :3])
    feats = np.ascontiguousarray(np.transpose(pts[..., 3:], (0, 2, 1))) if pts.shape[-1] > 3 else None
    l_xyz, l_feats = [xyz], [feats]
    for sa in model.SA_modules:
        nx, nf = sa_forward(sa, l_xyz[-1], l_feats[-1])
        l_xyz.append(nx)
        l_feats.append(nf)
    for i in range(-1, -(len(model.FP_modules) + 1), -1):
        l_feats[i - 1] = fp_forward(model.FP_modules[i], l_xyz[i - 1], l_xyz[i], l_feats[i - 1], l_feats[i])
    return l_xyz[0], l_feats[0]
