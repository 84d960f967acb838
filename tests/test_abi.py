"""CPU: the C-ABI library loads, exports exactly what include/ws3d_ops.h declares, and its host
entry points (no GPU involved) agree with the oracle and with the reference's own CPU functions."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import oracle
from refmods import load_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ws3d_ops.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ws3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ws3d_b200 import _C, build
    build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", _C.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T ws3d_" in l)
    assert exported == _declared()
    assert sorted(_C.EXPORTS) == _declared()          # the ctypes table covers the whole header
    assert _C.lib().ws3d_abi_version() == 1
    assert _C.last_error() == ""


def test_library_is_sm100a_only_and_uses_tma_and_clusters():
    from ws3d_b200 import _C
    sass = subprocess.run(["cuobjdump", "-sass", _C.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass and "sm_90" not in sass
    assert "UBLKCP" in sass                              # cp.async.bulk (TMA) staging
    assert "REDUX" in sass                               # warp arg-max in FPS
    assert re.search(r"UCGABAR|CGABAR", sass)            # cluster barrier of the FPS cluster kernel
    assert re.search(r"UTC\w*MMA", sass)                 # tcgen05.mma: the shared-MLP layer and the fused SA scale
    assert "UTMALDG" in sass                             # cp.async.bulk.tensor operand / weight loads
    assert "LDTM" in sass and "STTM" in sass             # tcgen05.ld epilogues; tcgen05.st: gathered tile + activations kept in TMEM
    assert "HMMA." not in sass.replace("UTCHMMA", "")    # no legacy mma.sync / wmma path


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ws3d_b200 import _C
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _C.lib()


def test_cuda_entry_points_reject_cpu_tensors():
    from ws3d_b200 import native
    x = torch.zeros(1, 8, 3)
    with pytest.raises(RuntimeError):
        native.ball_query_wrapper(1, 8, 2, 1.0, 4, x[:, :2], x, torch.zeros(1, 2, 4, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        native.boxes_iou_bev_gpu(torch.zeros(2, 5), torch.zeros(2, 5), torch.zeros(2, 2))


def _scene(n=3000, m=40, c=3, seed=3):
    from ws3d_b200 import synth
    rng = np.random.default_rng(seed)
    pts = synth.make_scene(seed, n)
    xyz = np.ascontiguousarray(pts[:, :3])
    feat = rng.normal(size=(n, c)).astype(np.float32)
    boxes = synth.make_boxes(xyz, m, seed=seed)
    boxes[:5, 0] += 300
    boxes[-1, 3:6] = 40.0
    return xyz, feat, boxes


def test_host_roipool_entry_points_match_oracle_and_reference():
    from ws3d_b200 import native, roipool3d_utils
    xyz, feat, boxes = _scene()
    s = 64
    pp = torch.zeros(boxes.shape[0], s, 3)
    pf = torch.zeros(boxes.shape[0], s, feat.shape[1])
    fl = torch.zeros(boxes.shape[0], dtype=torch.int64)
    native.roipool3d_cpu(torch.from_numpy(xyz), torch.from_numpy(boxes), torch.from_numpy(feat), pp, pf, fl)
    epp, epf, efl = oracle.roipool3d_cpu(xyz, boxes, feat, s)
    np.testing.assert_array_equal(pp.numpy(), epp)
    np.testing.assert_array_equal(pf.numpy(), epf)
    np.testing.assert_array_equal(fl.numpy(), efl)
    assert 0 < int(fl.sum()) < boxes.shape[0]
    masks = roipool3d_utils.pts_in_boxes3d_cpu(torch.from_numpy(xyz), torch.from_numpy(boxes))
    eflag = oracle.pts_in_boxes3d_cpu(xyz, boxes)
    np.testing.assert_array_equal(torch.stack(masks).numpy(), eflag > 0)
    ref = load_ref("roipool3d_cuda")
    if ref is not None:  # the reference's own CPU implementation (the only CPU code on its hot path)
        rpp, rpf, rfl = torch.zeros_like(pp), torch.zeros_like(pf), torch.zeros_like(fl)
        ref.roipool3d_cpu(torch.from_numpy(xyz), torch.from_numpy(boxes), torch.from_numpy(feat), rpp, rpf, rfl)
        assert torch.equal(rpp, pp) and torch.equal(rpf, pf) and torch.equal(rfl, fl)
        rflag = torch.zeros(boxes.shape[0], xyz.shape[0], dtype=torch.int64)
        ref.pts_in_boxes3d_cpu(rflag, torch.from_numpy(xyz), torch.from_numpy(boxes))
        np.testing.assert_array_equal(rflag.numpy(), eflag)


def test_next_row_entry_points_reject_cpu_tensors_and_bad_shapes():
    from ws3d_b200 import iou3d_utils, label_utils, native, proposal_utils
    with pytest.raises(RuntimeError):
        iou3d_utils.boxes_iou3d_aligned(torch.zeros(3, 7), torch.zeros(3, 7))
    with pytest.raises(RuntimeError):
        proposal_utils.cylinder_crop(torch.zeros(10, 3), torch.zeros(2, 2))
    with pytest.raises(RuntimeError):
        label_utils.generate_gaussian_training_labels(torch.zeros(10, 3), torch.zeros(2, 7))
    with pytest.raises(RuntimeError):
        native.radius_nms_device(torch.zeros(4, 2), 0.3)
    assert native.set_fps_mode(1) == 0 and native.set_fps_mode(0) == 1            # host-only knobs work without a GPU
    assert native.set_workspace_arena(3) == 0 and native.set_workspace_arena(0) == 3
    assert native.set_sm_budget(100) == 0 and native.set_sm_budget(0) == 100
    assert native.sa_mlp_fused_supported(96, 32, 64, 96, 128) and not native.sa_mlp_fused_supported(256, 16, 128, 196, 256)
    assert not native.sa_mlp_fused_supported(1, 4, 16, 16, 32)                    # nsample below 16 is not served
