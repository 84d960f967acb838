"""GPU: the tcgen05 shared-MLP layer vs PyTorch (conv1x1 + BN(eval) + ReLU [+ max-pool]).
TF32 inputs / FP32 accumulate.  The FP32 torch result (TF32 disabled) is the reference; tolerance
TOL = 3e-3 of the output scale over up to three chained layers: weights and intermediate activations
are rounded to nearest TF32 (2^-11 relative each), the raw first-layer input is truncated by the
tensor core (2^-10) -- the same class of error as cuDNN's default TF32 convolutions."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda:0"
TOL = 3e-3


def _mlp(spec, seed):
    from ws3d_b200 import pytorch_utils as pt_utils
    torch.manual_seed(seed)
    mlp = pt_utils.SharedMLP(list(spec), bn=True).to(dev).eval()
    g = torch.Generator(device="cpu").manual_seed(seed)
    for m in mlp.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)
    return mlp


def _ref(mlp, x, pool):
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            y = mlp(x)
            if pool:
                y = F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return y


@pytest.mark.parametrize("spec,B,M,K", [((4, 16, 16, 32), 2, 512, 16), ((4, 32, 32, 64), 1, 1024, 32), ((99, 64, 96, 128), 2, 256, 32),
                                        ((259, 128, 196, 256), 2, 64, 16), ((515, 256, 384, 512), 1, 64, 32), ((35, 40), 3, 100, 4),
                                        ((131, 128, 128, 256), 3, 32, 64), ((20, 24), 2, 8, 128), ((20, 70), 2, 6, 64)])
def test_folded_mlp_with_pool_matches_torch(spec, B, M, K):
    from ws3d_b200 import fused_mlp
    mlp = _mlp(spec, seed=sum(spec))
    x = torch.randn(B, spec[0], M, K, device=dev)
    want = _ref(mlp, x, pool=True)
    with torch.no_grad():
        got = fused_mlp.FoldedMLP(mlp)(x.view(B, spec[0], M * K), pool=K)
    assert got.shape == want.shape
    scale = float(want.abs().max()) + 1e-6
    assert float((got - want).abs().max()) <= TOL * scale + 1e-4, float((got - want).abs().max()) / scale
    # and without pooling (every layer's full output)
    want_full = _ref(mlp, x, pool=False).view(B, spec[-1], M * K)
    with torch.no_grad():
        got_full = fused_mlp.FoldedMLP(mlp)(x.view(B, spec[0], M * K))
    assert float((got_full - want_full).abs().max()) <= TOL * (float(want_full.abs().max()) + 1e-6) + 1e-4


@pytest.mark.parametrize("c1,c2,spec,n", [(256, 1, (257, 128, 128), 1024), (512, 96, (608, 256, 256), 512), (1024, 512, (1536, 512, 512), 64),
                                           (40, 0, (40, 24), 300)])
def test_two_input_first_layer_equals_concat(c1, c2, spec, n):
    from ws3d_b200 import fused_mlp
    mlp = _mlp(spec, seed=c1 + c2)
    B = 2
    a = torch.randn(B, c1, n, device=dev)
    b = torch.randn(B, c2, n, device=dev) if c2 else None
    x = a if b is None else torch.cat([a, b], dim=1)
    want = _ref(mlp, x.unsqueeze(-1), pool=False).squeeze(-1)
    with torch.no_grad():
        got = fused_mlp.FoldedMLP(mlp, first_split=(c1, c2))(a, b)
    assert float((got - want).abs().max()) <= TOL * (float(want.abs().max()) + 1e-6) + 1e-4


def test_backbone_fused_eval_path_close_to_fp32_path():
    from ws3d_b200 import models, synth
    torch.manual_seed(0)
    model = models.Pointnet2MSG(input_channels=1).to(dev).eval()
    pts = torch.from_numpy(synth.make_batch(2)).to(dev)
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = True
        _, fused = model(pts)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        _, exact = model(pts)
        torch.backends.cudnn.allow_tf32 = True
    rel = float((fused - exact).abs().max()) / (float(exact.abs().max()) + 1e-6)
    assert rel < 2e-2, rel     # 8 layers of TF32 rounding, same order as cuDNN's TF32 path
    assert torch.isfinite(fused).all()


def test_rpn_heads_on_the_layer_kernel_match_pytorch():
    """models.RPN: the cls / reg heads (Conv1d + BN + ReLU, Dropout, Conv1d with bias) through FoldedMLP."""
    from ws3d_b200 import models, synth
    torch.manual_seed(1)
    rpn = models.RPN().to(dev).eval()
    feats = torch.randn(2, 128, 4096, device=dev)
    with torch.no_grad():
        old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        want_cls, want_reg = rpn.rpn_cls_layer(feats), rpn.rpn_reg_layer(feats)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        got_cls, got_reg = rpn._head("rpn_cls_layer", feats), rpn._head("rpn_reg_layer", feats)
    assert got_cls.shape == want_cls.shape == (2, 1, 4096) and got_reg.shape == want_reg.shape == (2, 40, 4096)
    for got, want in ((got_cls, want_cls), (got_reg, want_reg)):
        assert float((got - want).abs().max()) <= TOL * (float(want.abs().max()) + 1e-6) + 1e-4


def _sa_reference(sa, xyz, feat, new_xyz):
    """The module's own FP32 PyTorch path (grouping kernels + cuDNN convs with TF32 off + max_pool2d)."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            return sa(xyz, feat, new_xyz=new_xyz)[1]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("n,m,c,radii,nsamples,mlps", [
    (4096, 1024, 1, [0.5, 1.0], [16, 32], [[1, 16, 16, 32], [1, 32, 32, 64]]),          # SA1 widths
    (2048, 512, 96, [1.0, 2.0], [16, 32], [[96, 64, 64, 128], [96, 64, 96, 128]]),      # SA2 widths
    (1500, 100, 5, [1.0, 2.0], [16, 64], [[5, 24, 40, 40], [5, 16, 16, 72]]),           # ragged tiles, odd widths, cross-warp pooling
    (1024, 33, 0, [2.0], [128], [[0, 16, 32, 48]]),                                       # no features, one centre per tile
    (600, 64, 128, [3.0], [16], [[128, 128, 128, 128]]),                                  # Stage-2 widths: 512 TMEM columns, 208 KB of weights
])
def test_fused_sa_scale_matches_module_fp32_path(n, m, c, radii, nsamples, mlps, monkeypatch):
    """csrc/sa_fused.cu (grouping + 3 layers + max-pool in one kernel, activations in tensor memory) against the same
    module on its FP32 PyTorch path; and against the per-layer tcgen05 path, which rounds identically except for
    the first-layer input (truncated there, rounded to nearest here)."""
    from ws3d_b200 import _C, pointnet2_modules, pointnet2_utils, synth
    torch.manual_seed(n + m)
    sa = pointnet2_modules.PointnetSAModuleMSG(npoint=m, radii=radii, nsamples=nsamples, mlps=[list(s) for s in mlps],
                                              use_xyz=True).to(dev).eval()
    g = torch.Generator(device="cpu").manual_seed(m)
    for mod in sa.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.2)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.2)
    pts = torch.from_numpy(synth.make_batch(2, n)).to(dev)
    xyz = pts[..., :3].contiguous()
    feat = torch.randn(2, c, n, device=dev) if c else None
    _, new_xyz = pointnet2_utils.sample_and_gather(xyz, m)
    want = _sa_reference(sa, xyz, feat, new_xyz)
    before = _C.launch_count()
    with torch.no_grad():
        got = sa(xyz, feat, new_xyz=new_xyz)[1]
    launched = _C.launch_count() - before
    assert len(sa.__dict__.get("_fused_scales", {})) == len(radii)            # the fused path was taken ...
    # ... ball queries (grid build + query) + ONE kernel per scale (+ one pass that packs channel-major features into operand rows)
    assert launched <= 2 * ((len(radii) + 1) // 2) + len(radii) + (1 if c >= 4 else 0), launched
    assert got.shape == want.shape
    scale = float(want.abs().max()) + 1e-6
    assert float((got - want).abs().max()) <= TOL * scale + 1e-4, float((got - want).abs().max()) / scale
    monkeypatch.setenv("WS3D_SA_FUSED", "0")
    with torch.no_grad():
        layered = sa(xyz, feat, new_xyz=new_xyz)[1]
    assert float((got - layered).abs().max()) <= TOL * scale + 1e-4


def test_three_interpolate_affine_epilogue():
    """ws3d_three_interpolate_affine = three_interpolate + scale1[c] * row1 + shift[c] (+ ReLU, + TF32 rounding)."""
    from ws3d_b200 import native, pointnet2_utils
    torch.manual_seed(3)
    B, c, m, n = 2, 24, 512, 2048
    pts = torch.randn(B, c, m, device=dev)
    idx = torch.randint(0, m, (B, n, 3), device=dev, dtype=torch.int32)
    w = torch.rand(B, n, 3, device=dev)
    w = w / w.sum(-1, keepdim=True)
    scale1, row1, shift = torch.randn(c, device=dev), torch.randn(B, n, device=dev), torch.randn(c, device=dev)
    base = pointnet2_utils.three_interpolate(pts, idx, w)
    out = torch.empty_like(base)
    native.three_interpolate_affine(B, c, m, n, pts, idx, w, None, None, None, 0, out)
    assert torch.equal(out, base)                                   # no epilogue: the plain kernel
    native.three_interpolate_affine(B, c, m, n, pts, idx, w, scale1, row1, shift, 1, out)
    want = torch.relu(base + scale1[None, :, None] * row1[:, None, :] + shift[None, :, None])
    assert float((out - want).abs().max()) < 1e-5
    native.three_interpolate_affine(B, c, m, n, pts, idx, w, None, None, shift, 2, out)
    assert int((out.view(torch.int32) & 0x1FFF).abs().max()) == 0   # rounded to TF32
    assert float((out - (base + shift[None, :, None])).abs().max()) < 2e-3 * float(base.abs().max())
    with pytest.raises(RuntimeError):
        native.three_interpolate_affine(B, 6, m, n, pts[:, :6].contiguous(), idx, w, None, None, shift[:6].contiguous(), 0,
                                        out[:, :6].contiguous())     # c % 4 != 0 is not served


@pytest.mark.parametrize("c_known,c_skip,spec,m,n", [(256, 1, (257, 128, 128), 1024, 4096), (64, 0, (64, 32, 48), 512, 1024),
                                                     (40, 1, (41, 24), 256, 300)])
def test_fp_module_premultiplied_first_layer(c_known, c_skip, spec, m, n, monkeypatch):
    """PointnetFPModule with a thin skip input: W_a applied to the known points, the product interpolated, skip term +
    shift + ReLU in the interpolation epilogue -- against the module's FP32 PyTorch path and the per-layer path."""
    from ws3d_b200 import pointnet2_modules, synth
    torch.manual_seed(c_known + n)
    fp = pointnet2_modules.PointnetFPModule(mlp=list(spec)).to(dev).eval()
    g = torch.Generator(device="cpu").manual_seed(5)
    for mod in fp.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.2)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.2)
    pts = torch.from_numpy(synth.make_batch(2, n)[..., :3].copy()).to(dev)
    known = pts[:, :m].contiguous()
    kf = torch.randn(2, c_known, m, device=dev)
    sf = torch.randn(2, c_skip, n, device=dev) if c_skip else None
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        want = fp(pts, known, sf, kf)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    with torch.no_grad():
        got = fp(pts, known, sf, kf)
    assert "_premul" in fp.__dict__ and got.shape == want.shape
    scale = float(want.abs().max()) + 1e-6
    assert float((got - want).abs().max()) <= TOL * scale + 1e-4, float((got - want).abs().max()) / scale
    monkeypatch.setenv("WS3D_FP_PREMUL", "0")
    with torch.no_grad():
        layered = fp(pts, known, sf, kf)
    assert float((got - layered).abs().max()) <= TOL * scale + 1e-4


@pytest.mark.parametrize("n,m,c,radii,nsamples,mlps", [
    (1024, 256, 256, [1.0, 2.0], [16, 32], [[256, 128, 196, 256], [256, 128, 196, 256]]),      # SA3 of the backbone
    (256, 64, 512, [2.0, 4.0], [16, 32], [[512, 256, 256, 512], [512, 256, 384, 512]]),        # SA4
    (2000, 100, 6, [1.5], [8], [[6, 24, 40]]),                                                   # two layers, ragged
])
def test_sa_first_layer_premultiplied(n, m, c, radii, nsamples, mlps, monkeypatch):
    """Scales outside the one-kernel path: W_f applied to the source points, ws3d_group_affine gathers the product and
    adds the FP32 coordinate term, shift and ReLU -- against the module's FP32 PyTorch path and the grouped path."""
    from ws3d_b200 import pointnet2_modules, pointnet2_utils, synth
    torch.manual_seed(n + c)
    sa = pointnet2_modules.PointnetSAModuleMSG(npoint=m, radii=radii, nsamples=nsamples, mlps=[list(s) for s in mlps],
                                              use_xyz=True).to(dev).eval()
    g = torch.Generator(device="cpu").manual_seed(c)
    for mod in sa.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.2)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.2)
    pts = torch.from_numpy(synth.make_batch(2, n)).to(dev)
    xyz = pts[..., :3].contiguous()
    feat = torch.randn(2, c, n, device=dev)
    _, new_xyz = pointnet2_utils.sample_and_gather(xyz, m)
    want = _sa_reference(sa, xyz, feat, new_xyz)
    with torch.no_grad():
        got = sa(xyz, feat, new_xyz=new_xyz)[1]
    assert len(sa.__dict__.get("_premul", {})) == len(radii) and "_fused_scales" not in sa.__dict__
    scale = float(want.abs().max()) + 1e-6
    assert float((got - want).abs().max()) <= TOL * scale + 1e-4, float((got - want).abs().max()) / scale
    monkeypatch.setenv("WS3D_SA_PREMUL", "0")
    with torch.no_grad():
        grouped_path = sa(xyz, feat, new_xyz=new_xyz)[1]
    assert float((got - grouped_path).abs().max()) <= TOL * scale + 1e-4
