"""CPU: host-side mirror of the reference interface -- module names / state_dict layout, box
helpers, synthetic data determinism, sharding arithmetic."""
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_rpn_state_dict_matches_reference_layout():
    """Names and shapes of every parameter/buffer equal the reference RPN's (fixture generated from
    /root/reference by tools/make_statedict_fixture.py): checkpoints are interchangeable."""
    from ws3d_b200 import models
    want = json.load(open(os.path.join(GOLD, "rpn_state_dict_keys.json")))
    got = {k: list(v.shape) for k, v in models.RPN().state_dict().items()}
    assert list(got.keys()) == list(want.keys())
    assert got == want
    assert sum(p.numel() for p in models.RPN().parameters()) == 3046201


def test_sa_module_spec_is_widened_in_place_like_the_reference():
    from ws3d_b200 import pointnet2_modules
    spec = [[4, 8, 16], [4, 8, 16]]
    sa = pointnet2_modules.PointnetSAModuleMSG(npoint=8, radii=[0.1, 0.2], nsamples=[4, 8], mlps=spec, use_xyz=True)
    assert spec[0][0] == 7 and sa.mlps[0].layer0.conv.weight.shape == (8, 7, 1, 1)
    single = pointnet2_modules.PointnetSAModule(mlp=[3, 8], npoint=None, use_xyz=True)   # GroupAll variant
    out = single(torch.rand(2, 10, 3), torch.rand(2, 3, 10))
    assert out[0] is None and out[1].shape == (2, 8, 1)


def test_box_helpers_match_reference_formulas():
    from ws3d_b200 import kitti_utils, synth
    b = torch.tensor([[1.0, 2.0, 3.0, 1.5, 1.6, 3.9, 0.3]])
    bev = kitti_utils.boxes3d_to_bev_torch(b)
    torch.testing.assert_close(bev, torch.tensor([[1 - 1.95, 3 - 0.8, 1 + 1.95, 3 + 0.8, 0.3]]))
    np.testing.assert_allclose(synth.boxes3d_to_bev(b.numpy()), bev.numpy(), rtol=1e-6)
    big = kitti_utils.enlarge_box3d(b, 0.5)
    torch.testing.assert_close(big, torch.tensor([[1.0, 2.5, 3.0, 2.5, 2.6, 4.9, 0.3]]))
    assert b[0, 1] == 2.0                                                # input untouched
    np.testing.assert_allclose(kitti_utils.enlarge_box3d(b.numpy(), 0.5), big.numpy())


def test_synthetic_scenes_are_deterministic_and_kitti_shaped():
    from ws3d_b200 import synth
    a, b = synth.make_scene(3), synth.make_scene(3)
    assert a.shape == (16384, 4) and a.dtype == np.float32 and np.array_equal(a, b)
    assert not np.array_equal(a, synth.make_scene(4))
    assert a[:, 0].min() >= -40 and a[:, 0].max() <= 40 and a[:, 2].min() >= 0 and a[:, 2].max() <= 70.4
    assert -0.5 <= a[:, 3].min() and a[:, 3].max() <= 0.5
    batch = synth.make_batch(2, 1024, first_scene=5)
    assert np.array_equal(batch[1], synth.make_scene(6, 1024))


def test_sharding_arithmetic():
    from ws3d_b200 import sharding
    assert sharding.scene_range(3, 8, 16) == (48, 64)
    parts = sharding.split_scenes(34, 8)
    assert parts[0] == (0, 5) and parts[-1] == (30, 34) and sum(e - s for s, e in parts) == 34
    assert all(parts[i][1] == parts[i + 1][0] for i in range(7))
    assert sharding.aggregate_throughput(16 * 16384, 8, 10.0) == 16 * 16384 * 8 / 0.01
    assert sharding.max_over_ranks(3.5) == 3.5


def test_folded_mlp_weights_equal_conv_bn_eval_on_cpu():
    """Host logic of the tensor-core MLP path: BN(eval) folding, TF32 rounding, padding and row replication
    (ws3d_b200/fused_mlp.py) -- checked on CPU with a plain matmul against the PyTorch module."""
    import torch
    from ws3d_b200 import fused_mlp, pytorch_utils as pt_utils
    torch.manual_seed(3)
    mlp = pt_utils.SharedMLP([7, 20, 40], bn=True).eval()
    for m in mlp.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    folded = fused_mlp.FoldedMLP(mlp)
    x = torch.randn(2, 7, 16, 1)
    with torch.no_grad():
        want = mlp(x).squeeze(-1)
    cur = x.squeeze(-1)
    for lay in folded.layers:
        assert lay.w.shape[0] % 128 == 0 and lay.w.shape[1] % 32 == 0
        copies = 1 << lay.rep_log2
        rows = 128 // copies
        for r in range(1, copies):   # replicated rows are exact copies
            assert torch.equal(lay.w[r * rows:r * rows + lay.c_out], lay.w[:lay.c_out])
            assert torch.equal(lay.shift[r * rows:r * rows + lay.c_out], lay.shift[:lay.c_out])
        assert torch.equal(lay.w, fused_mlp._round_tf32(lay.w))       # already TF32 values
        y = torch.einsum("oc,bce->boe", lay.w[:lay.c_out, :cur.shape[1]], cur) + lay.shift[:lay.c_out, None]
        cur = torch.relu(y) if lay.relu else y
    assert float((cur - want).abs().max()) < 5e-3 * float(want.abs().max())   # only the weights were rounded
    # _round_tf32: nearest, ties away from zero, 13 low mantissa bits cleared
    t = torch.tensor([1.0 + 2 ** -11, 1.0 + 2 ** -12, -(1.0 + 2 ** -11), 3.0])
    r = fused_mlp._round_tf32(t)
    assert r.tolist() == [1.0 + 2 ** -10, 1.0, -(1.0 + 2 ** -10), 3.0]
    assert fused_mlp.supported(4096, 32) and fused_mlp.supported(4096, 0) and not fused_mlp.supported(4098, 0) \
        and fused_mlp.supported(4096, 64) and not fused_mlp.supported(4096, 256) and not fused_mlp.supported(4096, 24)


def _randomize_bn_cpu(module, seed):
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)


def test_premultiplied_first_layers_fold_the_same_function_on_cpu():
    """Host logic of fused_mlp.FoldedSAFirstLayer / FoldedFPFirstLayer / FusedSAScale (BN folding, the [xyz ; features]
    column order, padding): the folded tensors, combined the way the kernels combine them, reproduce the module's
    conv + BN(eval) + ReLU on CPU.  (Weights are rounded to TF32 in the folded copies: 2e-3 of the output scale.)"""
    from ws3d_b200 import fused_mlp
    from ws3d_b200 import pytorch_utils as pt_utils
    torch.manual_seed(0)
    B, C, N, M, K = 2, 10, 50, 7, 4
    mlp = pt_utils.SharedMLP([C + 3, 12, 20, 24], bn=True).eval()
    _randomize_bn_cpu(mlp, 1)
    xyz, ctr, feat = torch.randn(B, N, 3), torch.randn(B, M, 3), torch.randn(B, C, N)
    idx = torch.randint(0, N, (B, M, K))
    gx = torch.gather(xyz, 1, idx.view(B, -1, 1).expand(-1, -1, 3)).view(B, M, K, 3) - ctr[:, :, None, :]
    gf = torch.gather(feat, 2, idx.view(B, 1, -1).expand(-1, C, -1)).view(B, C, M, K)
    grouped = torch.cat([gx.permute(0, 3, 1, 2), gf], dim=1)                    # xyz channels first (pointnet2_utils.py:257)
    with torch.no_grad():
        want1 = mlp[0](grouped)
        want3 = mlp(grouped)
    # set abstraction: P = W_f f on the source points, gathered, + W_x (xyz[i] - centre) + shift, ReLU
    first = fused_mlp.FoldedSAFirstLayer(mlp[0], C)
    assert fused_mlp.FoldedSAFirstLayer.eligible(mlp[0], C, 48) and not fused_mlp.FoldedSAFirstLayer.eligible(mlp[0], C, 50)
    P = torch.einsum("oc,bcn->bon", first.wf[:first.c_out, :C], feat)
    Pg = torch.gather(P, 2, idx.view(B, 1, -1).expand(-1, first.c_out, -1)).view(B, first.c_out, M, K)
    got1 = torch.relu(Pg + torch.einsum("oa,bmka->bomk", first.wx, gx) + first.shift[None, :, None, None])
    assert float((got1 - want1).abs().max()) < 2e-3 * float(want1.abs().max())
    # the one-kernel scale: three folded layers chained on the grouped tensor
    scale = fused_mlp.FusedSAScale(mlp)
    assert scale.widths == [12, 20, 24] and [tuple(w.shape) for w in scale.w] == [(16, 32), (32, 32), (32, 32)]
    act, k_real = grouped, C + 3
    for w, sh, c_out in zip(scale.w, scale.shift, scale.widths):
        act = torch.relu(torch.einsum("oc,bcmk->bomk", w[:c_out, :k_real], act) + sh[None, :c_out, None, None])
        k_real = c_out
    assert float((act - want3).abs().max()) < 3e-3 * float(want3.abs().max())
    # feature propagation: W_a on the known points, interpolated, + W_b[c] * skip + shift, ReLU
    fp_mlp = pt_utils.SharedMLP([C + 1, 16], bn=True).eval()
    _randomize_bn_cpu(fp_mlp, 2)
    n, m = 256, 8
    known, skip = torch.randn(B, C, m), torch.randn(B, 1, n)
    nn_idx = torch.randint(0, m, (B, n, 3))
    w3 = torch.rand(B, n, 3)
    w3 = w3 / w3.sum(-1, keepdim=True)

    def interp(f):
        return sum(torch.gather(f, 2, nn_idx[..., k].unsqueeze(1).expand(-1, f.shape[1], -1)) * w3[..., k].unsqueeze(1) for k in range(3))

    with torch.no_grad():
        want_fp = fp_mlp(torch.cat([interp(known), skip], dim=1).unsqueeze(-1)).squeeze(-1)
    assert fused_mlp.FoldedFPFirstLayer.eligible(fp_mlp[0], C, 1, m, n) and not fused_mlp.FoldedFPFirstLayer.eligible(fp_mlp[0], C, 2, m, n)
    ff = fused_mlp.FoldedFPFirstLayer(fp_mlp[0], C, 1)
    pre = torch.einsum("oc,bcm->bom", ff.wa[:ff.c_out, :C], known)
    got_fp = torch.relu(interp(pre) + ff.scale1[None, :, None] * skip + ff.shift[None, :, None])
    assert float((got_fp - want_fp).abs().max()) < 2e-3 * float(want_fp.abs().max())


def test_boundary_rejects_wrong_dtypes_and_short_buffers():
    """the ctypes boundary passes raw pointers, so dtype and size are enforced before the call --
    int64 indices or float64 / half features raise RuntimeError as the reference's pybind `tensor.data<T>()` does."""
    import pytest
    import torch

    from ws3d_b200 import _C
    f32, i32 = torch.float32, torch.int32
    ok = torch.zeros(8, dtype=f32)
    assert _C.require("t", (ok, f32, 8), (None, i32, 4), cuda=False)
    with pytest.raises(RuntimeError, match="must be torch.int32"):
        _C.require("t", (torch.zeros(8, dtype=torch.int64), i32, None), cuda=False)
    with pytest.raises(RuntimeError, match="must be torch.float32"):
        _C.require("t", (torch.zeros(8, dtype=torch.float64), f32, None), cuda=False)
    with pytest.raises(RuntimeError, match="must be torch.float32"):
        _C.require("t", (torch.zeros(8, dtype=torch.float16), f32, None), cuda=False)
    with pytest.raises(RuntimeError, match="need 9"):
        _C.require("t", (ok, f32, 9), cuda=False)
    with pytest.raises(RuntimeError, match="contiguous"):
        _C.require("t", (torch.zeros(4, 4)[:, 0], f32, None), cuda=False)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _C.require("t", (ok, f32, None))


def test_native_wrappers_validate_before_touching_the_library():
    """Every wrapper of the reference's function tables raises on an int64 idx / CPU tensor without a launch."""
    import pytest
    import torch

    from ws3d_b200 import native
    x = torch.zeros(1, 16, 3)
    with pytest.raises(RuntimeError):
        native.furthest_point_sampling_wrapper(1, 16, 4, x, torch.zeros(1, 16), torch.zeros(1, 4, dtype=torch.int32))   # CPU tensors
    with pytest.raises(RuntimeError):
        native.boxes_iou_bev_gpu(torch.zeros(4, 7), torch.zeros(4, 5), torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        native.nms_gpu(torch.zeros(4, 5), torch.zeros(4, dtype=torch.int64), 0.5)


def test_streamed_runner_rejects_more_buffer_sets_than_scratch_arenas():
    import inspect

    from ws3d_b200 import graphs
    src = inspect.getsource(graphs.StreamedBackboneRunner.__init__)
    assert "num_arenas" in src and "raise ValueError" in src


def test_plan_tensors_are_listed_in_one_fixed_order():
    """The cold-start capture of StreamedBackboneRunner copies its plan into the throughput capture's plan tensor by tensor:
    both flattenings must pair the same entries whatever the dict insertion order."""
    from ws3d_b200.graphs import _plan_tensors
    t = [torch.full((2,), float(k)) for k in range(7)]
    a = {"xyz": [t[0], t[1]], "idx": [(t[2], t[3])], "nn": [(t[4], t[5])], "xyz0": t[6], "note": "x", "feat0": None}
    b = {"feat0": None, "xyz0": t[6], "nn": [(t[4], t[5])], "note": "y", "idx": [(t[2], t[3])], "xyz": [t[0], t[1]]}
    fa, fb = _plan_tensors(a), _plan_tensors(b)
    assert len(fa) == 7 and all(x is y for x, y in zip(fa, fb))
    assert [int(x[0]) for x in fa] == [2, 3, 4, 5, 0, 1, 6]          # keys sorted: idx, nn, xyz, xyz0


def test_throughput_sampler_packing_rule_is_exposed_without_a_gpu():
    """ws3d_fps_clouds_per_cta: packed CTAs in throughput mode only (two clouds at 16384 points, four at 8192, eight at 4096),
    one cloud per CTA otherwise until the batch exceeds the SM count."""
    from ws3d_b200 import native
    prev = native.set_fps_mode(1)
    try:
        assert native.fps_clouds_per_cta(16, 16384) == 2
        assert native.fps_clouds_per_cta(16, 8192) == 4
        assert native.fps_clouds_per_cta(16, 4096) == 8
        assert native.fps_clouds_per_cta(3, 4096) == 3           # never more than the batch
        native.set_fps_mode(0)
        assert native.fps_clouds_per_cta(16, 16384) == 1
    finally:
        native.set_fps_mode(prev)
