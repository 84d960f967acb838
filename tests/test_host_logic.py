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
