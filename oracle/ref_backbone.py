"""TEST / BASELINE INFRASTRUCTURE ONLY: the PointNet++ backbone forward driven by the REFERENCE's own CUDA kernels
(oracle/_ref/pointnet2_cuda.so: the unmodified reference sources recompiled for sm_100a by oracle/build_ref.sh) in the
reference's unfused sequencing (pointnet2_modules.py:19-55,127-156; pointnet2_utils.py:241-264), with the module's own
PyTorch / cuDNN MLPs.  This is BASELINE.md section 3a's arm -- "the reference kernels on the same B200" -- used by
bench.py's `oracle_gpu` record, tools/backbone_compare.py and the parity tests.  The product never imports it."""
import importlib.machinery
import importlib.util
import os

import torch
import torch.nn.functional as F

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_cache = {}


def load_ref(name: str):
    """name in {pointnet2_cuda, iou3d_cuda, roipool3d_cuda}: the reference extension under a private handle (never
    registered in sys.modules, so it cannot shadow the product's drop-in of the same name); None when not built."""
    if name not in _cache:
        path = os.path.join(REF_DIR, name + ".so")
        mod = None
        if os.path.exists(path):
            loader = importlib.machinery.ExtensionFileLoader(name, path)
            spec = importlib.util.spec_from_loader(name, loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
        _cache[name] = mod
    return _cache[name]


class RefOps:
    """The reference wrappers' call pattern (pointnet2_utils.py) on the reference extension module."""

    def __init__(self, mod=None):
        self.m = mod or load_ref("pointnet2_cuda")
        assert self.m is not None, "oracle/_ref/pointnet2_cuda.so is not built (oracle/build_ref.sh)"

    def fps(self, xyz, npoint):
        B, N, _ = xyz.shape
        out = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, device=xyz.device)
        self.m.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out)
        return out

    def gather(self, feats, idx):
        B, C, N = feats.shape
        out = torch.empty((B, C, idx.shape[1]), device=feats.device)
        self.m.gather_points_wrapper(B, C, N, idx.shape[1], feats, idx, out)
        return out

    def ball_query(self, r, k, xyz, new_xyz):
        B, N, _ = xyz.shape
        idx = torch.zeros((B, new_xyz.shape[1], k), dtype=torch.int32, device=xyz.device)
        self.m.ball_query_wrapper(B, N, new_xyz.shape[1], r, k, new_xyz, xyz, idx)
        return idx

    def group(self, feats, idx):
        B, C, N = feats.shape
        out = torch.empty((B, C, idx.shape[1], idx.shape[2]), device=feats.device)
        self.m.group_points_wrapper(B, C, N, idx.shape[1], idx.shape[2], feats, idx, out)
        return out

    def three_nn(self, unknown, known):
        B, N, _ = unknown.shape
        d2 = torch.empty((B, N, 3), device=unknown.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
        self.m.three_nn_wrapper(B, N, known.shape[1], unknown, known, d2, idx)
        return torch.sqrt(d2), idx

    def interpolate(self, feats, idx, w):
        B, c, m = feats.shape
        out = torch.empty((B, c, idx.shape[1]), device=feats.device)
        self.m.three_interpolate_wrapper(B, c, m, idx.shape[1], feats, idx, w, out)
        return out


def ref_forward(model, ops: RefOps, pc: torch.Tensor):
    """model: ws3d_b200.models.Pointnet2MSG (its SharedMLPs are evaluated by PyTorch); pc (B,N,3+C) CUDA."""
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    l_xyz, l_f = [xyz], [feats]
    for sa in model.SA_modules:
        x, f = l_xyz[-1], l_f[-1]
        xt = x.transpose(1, 2).contiguous()
        new_xyz = ops.gather(xt, ops.fps(x, sa.npoint)).transpose(1, 2).contiguous()       # pointnet2_modules.py:30-35
        outs = []
        for g, mlp in zip(sa.groupers, sa.mlps):
            idx = ops.ball_query(g.radius, g.nsample, x, new_xyz)                          # pointnet2_utils.py:249
            gx = ops.group(x.transpose(1, 2).contiguous(), idx)
            gx -= new_xyz.transpose(1, 2).unsqueeze(-1)                                    # :252
            y = mlp(torch.cat([gx, ops.group(f, idx)], dim=1))                             # :257, modules :40
            outs.append(F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1))           # :41-44
        l_xyz.append(new_xyz)
        l_f.append(torch.cat(outs, dim=1))
    for i in range(-1, -(len(model.FP_modules) + 1), -1):
        dist, idx = ops.three_nn(l_xyz[i - 1], l_xyz[i])                                   # :139
        recip = 1.0 / (dist + 1e-8)
        w = recip / torch.sum(recip, dim=2, keepdim=True)                                  # :140-142
        interp = ops.interpolate(l_f[i], idx, w)
        nf = torch.cat([interp, l_f[i - 1]], dim=1) if l_f[i - 1] is not None else interp
        l_f[i - 1] = model.FP_modules[i].mlp(nf.unsqueeze(-1)).squeeze(-1)                 # :154
    return l_xyz[0], l_f[0]
