"""Training-mode SharedMLP on this library's kernels (BASELINE config 3).

The reference trains its shared MLPs as cuDNN conv -> BatchNorm -> ReLU (-> max-pool) with autograd's mirror images
(pytorch_utils.py:5-32, pointnet2_modules.py:40-44).  `shared_mlp_train` runs the same layers -- same parameters, same
batch statistics, same running-statistics update, same gradients -- as one autograd Function per layer on top of
csrc/train_mlp.cu and the tcgen05 layer kernel:

    forward   Y = W X (+ per-channel sum / sum of squares from the GEMM epilogue) -> mean / invstd / running stats ->
              Z = relu(Y * scale + shift), the last layer of a set-abstraction scale max-pooled with its arg-max
    backward  (sum dA, sum dA xhat) -> dY -> dX = W^T dY (tcgen05) and dW = dY X^T (tcgen05, split-K)

on the channel-major (B, C, cols) activations the grouping / interpolation kernels produce: no NCHW <-> NHWC copies, no
separate BatchNorm / ReLU / max-pool passes, no torch.cat for the two-input layers of feature propagation.
Numerics: TF32 operands, FP32 accumulation (what cuDNN uses for these convolutions by default); statistics in double.
"""
import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused_mlp, native

DESCRIPTION = ("this repo's kernels: tcgen05 TF32 GEMMs for Y = W X (batch statistics reduced in the epilogue), dX = W^T dY and "
               "dW = dY X^T (split-K), fused BatchNorm(batch statistics) + ReLU (+ max-pool / arg-max) forward and backward")

_TILE_M, _CHUNK_K = 128, 32


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def enabled_for(mlp: nn.Sequential, x: torch.Tensor, pool: int = 0) -> bool:
    """Training with autograd on a CUDA tensor, TF32 convolutions allowed, every block conv1x1 -> [BatchNorm] -> [ReLU]
    (fused_mlp.block_foldable), columns a multiple of 4, pooling groups a power of two in [4, 128]."""
    if os.environ.get("WS3D_TRAIN_MLP", "1") == "0" or not x.is_cuda or x.dtype != torch.float32:
        return False
    if not (torch.is_grad_enabled() and torch.backends.cudnn.allow_tf32):
        return False
    blocks = [m for m in mlp if not isinstance(m, nn.Dropout)]
    if not blocks or not all(fused_mlp.block_foldable(b) for b in blocks):
        return False
    for b in blocks:
        bn = b.bn.bn if hasattr(b, "bn") else None
        if bn is not None and not (bn.training and bn.affine and bn.track_running_stats and bn.momentum is not None):
            return False
    cols = x.shape[2] if x.dim() == 3 else x.shape[2] * x.shape[3]
    return cols % 4 == 0 and (pool == 0 or (4 <= pool <= 128 and pool & (pool - 1) == 0 and cols % pool == 0))


def _buffer(owner: nn.Module, key, shape, device, dtype=torch.float32) -> torch.Tensor:
    """A zero-initialised scratch tensor cached on the module (padded weight images: only the live block is rewritten
    every step, the padding stays zero)."""
    cache = owner.__dict__.setdefault("_ws3d_train_buf", {})
    t = cache.get(key)
    if t is None or t.shape != torch.Size(shape) or t.device != device:
        t = cache[key] = torch.zeros(shape, dtype=dtype, device=device)
    return t


class _TrainLayer(torch.autograd.Function):
    """z = [maxpool_K] act(bn(W [x1 ; x2] (+ bias))) with batch statistics; see the module docstring."""

    @staticmethod
    def forward(ctx, x1, x2, weight, gamma, beta, bias, conv, bn, relu: bool, pool: int, round_out: bool):
        B, c1, cols = x1.shape
        c2 = 0 if x2 is None else x2.shape[1]
        c_out, c_in = weight.shape[0], c1 + c2
        assert weight.numel() == c_out * c_in, "weight does not match the input channels"
        dev = x1.device
        w2d = weight.detach().reshape(c_out, c_in)
        c_out_pad, k1, k2 = _ceil(c_out, _TILE_M), _ceil(c1, _CHUNK_K), (_ceil(c2, _CHUNK_K) if c2 else 0)
        wp = _buffer(conv, ("w", c1, c2), (c_out_pad, k1 + k2), dev)
        wp[:c_out, :c1].copy_(w2d[:, :c1])
        if c2:
            wp[:c_out, k1:k1 + c2].copy_(w2d[:, c1:])
        zero_shift = _buffer(conv, ("zs",), (c_out_pad,), dev)
        y = torch.empty((B, c_out, cols), dtype=torch.float32, device=dev)
        stats = torch.zeros(2 * c_out, dtype=torch.float64, device=dev)
        x1c = x1.contiguous()
        x2c = None if x2 is None else x2.contiguous()
        native.mlp_layer_stats(B, c_out, c_out_pad, c1, c2, cols, wp, zero_shift, x1c, x2c, y, stats)
        scale = torch.empty(c_out, dtype=torch.float32, device=dev)
        shift, mean, invstd = torch.empty_like(scale), torch.empty_like(scale), torch.empty_like(scale)
        count = float(B) * cols
        if bn is not None:
            native.bn_finalize(c_out, count, stats, gamma.detach(), beta.detach(), bn.eps, bn.momentum, bn.running_mean, bn.running_var,
                               scale, shift, mean, invstd)
            bn.num_batches_tracked.add_(1)
        else:   # no batch norm: z = act(y + bias)
            scale.fill_(1.0)
            invstd.fill_(1.0)
            mean.zero_()
            if bias is not None:
                shift.copy_(bias.detach())
            else:
                shift.zero_()
        flags = int(relu) | (2 if round_out else 0)
        if pool:
            z = torch.empty((B, c_out, cols // pool), dtype=torch.float32, device=dev)
            arg = torch.empty((B, c_out, cols // pool), dtype=torch.uint8, device=dev)
        else:
            z, arg = torch.empty_like(y), None
        native.bn_relu_apply(B, c_out, cols, pool, y, scale, shift, flags, z, arg)
        ctx.save_for_backward(x1c, x2c, y, arg, scale, shift, mean, invstd, weight)
        ctx.conv, ctx.has_bn, ctx.relu, ctx.pool, ctx.count, ctx.has_bias = conv, bn is not None, relu, pool, count, bias is not None
        return z

    @staticmethod
    def backward(ctx, dz):
        x1, x2, y, arg, scale, shift, mean, invstd, weight = ctx.saved_tensors
        B, c_out, cols = y.shape
        c1 = x1.shape[1]
        c2 = 0 if x2 is None else x2.shape[1]
        c_in, dev, pool, flags = c1 + c2, y.device, ctx.pool, int(ctx.relu)
        dz = dz.contiguous()
        sums = torch.zeros(2 * c_out, dtype=torch.float64, device=dev)
        native.bn_relu_bwd_reduce(B, c_out, cols, pool, y, dz, arg, scale, shift, mean, invstd, flags, sums)
        dy = torch.empty_like(y)
        native.bn_relu_bwd_apply(B, c_out, cols, pool, y, dz, arg, scale, shift, mean, invstd, flags, sums, ctx.count if ctx.has_bn else 0.0, dy)
        need = ctx.needs_input_grad
        dgamma = sums[c_out:].float() if (ctx.has_bn and need[3]) else None
        dbeta = sums[:c_out].float() if (ctx.has_bn and need[4]) else None
        dbias = sums[:c_out].float() if (ctx.has_bias and need[5]) else None
        dw = None
        if need[2]:
            dw2d = torch.zeros((c_out, c_in), dtype=torch.float32, device=dev)
            native.mlp_wgrad(B, c_out, c1, cols, dy, x1, dw2d, c_in, 0)
            if c2:
                native.mlp_wgrad(B, c_out, c2, cols, dy, x2, dw2d, c_in, c1)
            dw = dw2d.view_as(weight)
        dx1 = dx2 = None
        w2d = weight.detach().reshape(c_out, c_in)
        k_pad = _ceil(c_out, _CHUNK_K)

        def dgrad(lo: int, hi: int) -> torch.Tensor:
            """dX[:, lo:hi] = W[:, lo:hi]^T dY as its own contiguous tensor (the two inputs of a feature-propagation layer get
            one launch each: dY is read twice, but neither consumer has to copy a strided slice of a joint dX)."""
            rows, rows_pad = hi - lo, _ceil(hi - lo, _TILE_M)
            wt = _buffer(ctx.conv, ("wt", lo, hi, c_out), (rows_pad, k_pad), dev)
            wt[:rows, :c_out].copy_(w2d[:, lo:hi].t())
            zero_shift = _buffer(ctx.conv, ("zst", rows_pad), (rows_pad,), dev)
            out = torch.empty((B, rows, cols), dtype=torch.float32, device=dev)
            native.mlp_layer(B, rows, rows_pad, c_out, 0, cols, wt, zero_shift, dy, None, out, 0, 0)
            return out

        if need[0]:
            dx1 = dgrad(0, c1)
        if c2 and need[1]:
            dx2 = dgrad(c1, c_in)
        return dx1, dx2, dw, dgamma, dbeta, dbias, None, None, None, None, None


def shared_mlp_train(mlp: nn.Sequential, x1: torch.Tensor, x2: Optional[torch.Tensor] = None, pool: int = 0) -> torch.Tensor:
    """A SharedMLP / head (sequence of conv1x1 [+ BN] [+ ReLU] blocks, nn.Dropout in between allowed) in TRAINING mode.
    x1 (B, c1, cols) [, x2 (B, c2, cols): second input of the first layer, e.g. the skip features of feature propagation]
    -> (B, c_last, cols), or (B, c_last, cols / pool) with the last layer max-pooled over runs of `pool` columns."""
    mods = list(mlp)
    blocks = [m for m in mods if not isinstance(m, nn.Dropout)]
    cur1, cur2, seen = x1, x2, 0
    for m in mods:
        if isinstance(m, nn.Dropout):
            cur1 = F.dropout(cur1, m.p, training=m.training)
            continue
        seen += 1
        last = seen == len(blocks)
        conv = m.conv
        bn = m.bn.bn if hasattr(m, "bn") else None
        relu = hasattr(m, "activation")
        cur1 = _TrainLayer.apply(cur1, cur2, conv.weight, None if bn is None else bn.weight, None if bn is None else bn.bias, conv.bias,
                                 conv, bn, relu, pool if last else 0, not last)
        cur2 = None
    return cur1
