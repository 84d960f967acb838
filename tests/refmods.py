"""Loaders for the checkers used by the tests: the reference's own extension modules built into
oracle/_ref/ (TEST INFRASTRUCTURE, see oracle/build_ref.sh) under private names, so they never
shadow the product's drop-in modules.  On a machine with a CUDA device a missing oracle/_ref is a
test FAILURE (require_ref), not a silent downgrade to oracle-only checks."""
import os

import pytest

from oracle.ref_backbone import REF_DIR, load_ref  # noqa: F401


def have_ref(name: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, name + ".so"))


def require_ref(name: str):
    """The reference extension `name`, or a hard test failure: the parity claim against the reference's own kernels
    must not evaporate because oracle/_ref did not travel to the GPU box."""
    mod = load_ref(name)
    if mod is None:
        pytest.fail(f"oracle/_ref/{name}.so is missing: run oracle/build_ref.sh where /root/reference exists "
                    "(it is git-ignored but travels with the gpurun snapshot)")
    return mod


REFPY_ZIP = os.path.join(REF_DIR, "refpy.zip")
_REF_PACKAGES = ("pointnet2_lib", "lib")


def load_reference_python(natives: dict):
    """The reference's own Python op wrappers (oracle/_ref/refpy.zip: pointnet2_utils / pointnet2_modules / pytorch_utils /
    iou3d_utils / roipool3d_utils, UNMODIFIED) imported with `natives` = {"pointnet2_cuda": mod, "iou3d_cuda": mod,
    "roipool3d_cuda": mod} as the extension modules they bind at import time.  Every call returns a FRESH set of module
    objects, so the same files can be loaded once over the product's drop-ins and once over the reference extensions."""
    import sys
    import types
    if not os.path.exists(REFPY_ZIP):
        return None

    def purge():
        for name in list(sys.modules):
            if name.split(".")[0] in _REF_PACKAGES:
                del sys.modules[name]

    saved = {k: sys.modules.get(k) for k in natives}
    purge()
    sys.modules.update(natives)
    sys.path.insert(0, REFPY_ZIP)
    try:
        import lib.utils.iou3d.iou3d_utils as iou3d_utils
        import lib.utils.roipool3d.roipool3d_utils as roipool3d_utils
        import pointnet2_lib.pointnet2.pointnet2_modules as pointnet2_modules
        import pointnet2_lib.pointnet2.pointnet2_utils as pointnet2_utils
    finally:
        sys.path.remove(REFPY_ZIP)
        purge()
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return types.SimpleNamespace(pointnet2_utils=pointnet2_utils, pointnet2_modules=pointnet2_modules, iou3d_utils=iou3d_utils,
                                 roipool3d_utils=roipool3d_utils)


def require_reference_python(natives: dict):
    ns = load_reference_python(natives)
    if ns is None:
        pytest.fail("oracle/_ref/refpy.zip is missing: run oracle/build_ref.sh where /root/reference exists")
    return ns


# ---- TF32 emulation of the training-mode shared MLP (checker for ws3d_b200/train_mlp.py) ----------------------------------------
def emulated_shared_mlp_train(mlp, x1, x2=None, pool: int = 0, product_dtype=None):
    """Plain PyTorch (float64 products, autograd) restatement of `train_mlp.shared_mlp_train` WITH the operand rounding of the
    tensor-core path: a GEMM reads its operands truncated to TF32 (10 explicit mantissa bits: the tcgen05 kind::tf32 read),
    the activations handed to the next layer are rounded to the nearest TF32 value, everything else is FP32 / FP64.
    Against this reference a ReLU mask or a max-pool arg-max flips only on ties at FP32 rounding level, so forward values
    and gradients can be held to tolerances two orders tighter than against the cuDNN TF32 path (whose rounding differs).
    `product_dtype=torch.float32` forms the products with FP32 accumulation instead (cuBLAS SGEMM, TF32 off): the same
    arithmetic up to summation order, used by the whole-network test as its noise floor."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    class Trunc(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)

        @staticmethod
        def backward(ctx, g):
            return g

    class RoundNearest(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

        @staticmethod
        def backward(ctx, g):
            return g

    mods = list(mlp)
    blocks = [m for m in mods if not isinstance(m, nn.Dropout)]
    cur = x1 if x2 is None else torch.cat([x1, x2], dim=1)
    seen = 0
    for m in mods:
        if isinstance(m, nn.Dropout):
            cur = F.dropout(cur, m.p, training=m.training)
            continue
        seen += 1
        last = seen == len(blocks)
        conv = m.conv
        bn = m.bn.bn if hasattr(m, "bn") else None
        w = Trunc.apply(conv.weight.reshape(conv.weight.shape[0], -1))
        if product_dtype is torch.float32:
            assert not torch.backends.cuda.matmul.allow_tf32
            y = torch.matmul(w, Trunc.apply(cur))
        else:
            y = torch.matmul(w.double(), Trunc.apply(cur).double()).float()
        if bn is not None:
            y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)
            bn.num_batches_tracked.add_(1)
        elif conv.bias is not None:
            y = y + conv.bias[None, :, None]
        if hasattr(m, "activation"):
            y = F.relu(y)
        cur = y if last else RoundNearest.apply(y)
    if pool:
        B, C, cols = cur.shape
        cur = cur.view(B, C, cols // pool, pool).max(dim=3)[0]
    return cur
