"""Inference-time SharedMLP on the tcgen05 tensor cores.

`FoldedMLP` is a read-only view of a `pytorch_utils.SharedMLP` (conv1x1 [+ BN] [+ ReLU] per layer):
BatchNorm (eval statistics) is folded into the conv weight and a per-channel shift, weights are
zero-padded to the kernel's tile sizes once, and every layer runs as one `ws3d_mlp_layer` launch
(GEMM + shift + ReLU, and for the last layer of a set-abstraction scale the max-pool over nsample).
Numerics: TF32 inputs with FP32 accumulation -- what PyTorch's cuDNN convolutions use by default
(`torch.backends.cudnn.allow_tf32`); when TF32 is disallowed the modules keep the PyTorch FP32 path.
Training (batch statistics, autograd) always uses the PyTorch path.
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import native

_TILE_M = 128
_CHUNK_K = 32


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def _round_tf32(t: torch.Tensor) -> torch.Tensor:
    """Round FP32 to the nearest TF32 (10-bit mantissa), ties away from zero."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def block_foldable(block: nn.Module) -> bool:
    """A block the folded kernels reproduce: children exactly [conv 1x1, optional BatchNorm wrapper (`bn.bn`), optional
    nn.ReLU] in that order.  pytorch_utils also builds preact blocks (norm / activation BEFORE the convolution), instance
    norm (`in`) and arbitrary activations: those keep the PyTorch path."""
    names = [n for n, _ in block.named_children()]
    if not names or names[0] != "conv" or names != [n for n in ("conv", "bn", "activation") if n in names]:
        return False
    conv = block.conv
    if not isinstance(conv, (nn.Conv1d, nn.Conv2d)) or any(k != 1 for k in conv.kernel_size) or any(k != 1 for k in conv.stride):
        return False
    if any(p != 0 for p in conv.padding) or conv.groups != 1:
        return False
    if "bn" in names and not isinstance(getattr(block.bn, "bn", None), (nn.BatchNorm1d, nn.BatchNorm2d)):
        return False
    return "activation" not in names or isinstance(block.activation, nn.ReLU)


def mlp_foldable(mlp) -> bool:
    blocks = [m for m in mlp if not isinstance(m, nn.Dropout)]   # dropout is the identity in eval mode
    return len(blocks) > 0 and all(block_foldable(b) for b in blocks)


def _module_foldable(module: nn.Module) -> bool:
    """Structure check of every shared MLP the module owns (`mlps` of an SA layer, `mlp` of an FP layer, the heads of
    the RPN), cached on the module."""
    ok = module.__dict__.get("_ws3d_foldable")
    if ok is None:
        seqs = []
        if isinstance(getattr(module, "mlps", None), nn.ModuleList):
            seqs += list(module.mlps)
        if isinstance(getattr(module, "mlp", None), nn.Sequential):
            seqs.append(module.mlp)
        for name in ("rpn_cls_layer", "rpn_reg_layer"):
            if isinstance(getattr(module, name, None), nn.Sequential):
                seqs.append(getattr(module, name))
        ok = module.__dict__["_ws3d_foldable"] = all(mlp_foldable(q) for q in seqs)
    return ok


def enabled_for(module: nn.Module) -> bool:
    """The fused path is taken in eval mode, without autograd, when TF32 convolutions are allowed and every block of the
    module's shared MLPs is conv -> [BatchNorm] -> [ReLU] (anything else keeps the PyTorch path)."""
    return ((not module.training) and (not torch.is_grad_enabled()) and torch.backends.cudnn.allow_tf32
            and _module_foldable(module))


class _Layer:
    __slots__ = ("w", "shift", "c_out", "c_out_pad", "relu", "splits", "rep_log2")


class FoldedMLP:
    """Folded, padded weights of one SharedMLP.  `first_split=(c1, c2)` says the first layer's input
    arrives as two tensors (channels c1 then c2), which are read in place instead of concatenated."""

    def __init__(self, mlp: nn.Sequential, first_split: Optional[Tuple[int, int]] = None):
        self.layers: List[_Layer] = []
        self._versions = None
        self._mlp = mlp
        self._first_split = first_split
        self._build()

    def _params(self):
        return [p for p in self._mlp.parameters()] + [b for b in self._mlp.buffers()]

    def _stamp(self):
        return tuple((p.data_ptr(), p._version) for p in self._params())

    @staticmethod
    def fold_block(block: nn.Module):
        """(W', shift, relu) of one conv1x1 [+ BatchNorm(eval)] [+ ReLU] block: y = relu(W' x + shift)."""
        conv = block.conv
        assert conv.kernel_size in ((1, 1), (1,)) and conv.stride in ((1, 1), (1,)), "shared MLPs are 1x1 convolutions"
        w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
        shift = conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
        if hasattr(block, "bn"):
            bn = block.bn.bn
            scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
            w = w * scale[:, None]
            shift = (shift - bn.running_mean) * scale + bn.bias.detach()
        return w, shift, hasattr(block, "activation")

    def _matrices(self):
        return [self.fold_block(block) for block in self._mlp]

    def _build(self):
        self.layers = []
        for li, (w, shift, relu) in enumerate(self._matrices()):
            c_out, c_in = w.shape
            splits = self._first_split if (li == 0 and self._first_split is not None and self._first_split[1] > 0) else (c_in, 0)
            assert sum(splits) == c_in
            k1, k2 = _ceil(splits[0], _CHUNK_K), (_ceil(splits[1], _CHUNK_K) if splits[1] else 0)
            lay = _Layer()
            lay.c_out, lay.c_out_pad = c_out, _ceil(c_out, _TILE_M)
            # narrow layers: replicate the rows over the 128 MMA rows so that every epilogue warp has channels
            lay.rep_log2 = 2 if c_out <= 32 else (1 if c_out <= 64 else 0)
            w = _round_tf32(w)  # the tensor core truncates FP32 operands to TF32: round the weights to nearest first
            wp = torch.zeros((lay.c_out_pad, k1 + k2), dtype=torch.float32, device=w.device)
            wp[:c_out, :splits[0]] = w[:, :splits[0]]
            if splits[1]:
                wp[:c_out, k1:k1 + splits[1]] = w[:, splits[0]:]
            sp = torch.zeros(lay.c_out_pad, dtype=torch.float32, device=w.device)
            sp[:c_out] = shift
            for r in range(1, 1 << lay.rep_log2):
                o = r * (_TILE_M >> lay.rep_log2)
                wp[o:o + c_out] = wp[:c_out]
                sp[o:o + c_out] = sp[:c_out]
            lay.w, lay.shift, lay.splits = wp.contiguous(), sp, splits
            lay.relu = relu
            self.layers.append(lay)
        self._versions = self._stamp()

    def __call__(self, x1: torch.Tensor, x2: Optional[torch.Tensor] = None, pool: int = 0,
                 out: Optional[torch.Tensor] = None, out_coff: int = 0) -> torch.Tensor:
        """x1 (B, c1, cols) [, x2 (B, c2, cols)] -> (B, c_last, cols) or (B, c_last, cols // pool).
        `out` (B, c_total, .) / `out_coff`: the last layer writes its channels into that slot instead of a new tensor."""
        if self._versions != self._stamp():
            self._build()  # parameters were updated (training step, load_state_dict, .to())
        B, _, cols = x1.shape
        cur1, cur2 = x1, x2
        for li, lay in enumerate(self.layers):
            last = li == len(self.layers) - 1
            p = pool if last else 0
            into = out if (last and out is not None) else None
            res = into if into is not None else torch.empty((B, lay.c_out, cols // p if p else cols), dtype=torch.float32,
                                                            device=x1.device)
            c1 = cur1.shape[1]
            c2 = cur2.shape[1] if cur2 is not None else 0
            assert (c1, c2) == tuple(lay.splits), ((c1, c2), lay.splits)
            rep = lay.rep_log2
            while p and (_TILE_M >> rep) % p:   # a pooling group must fit the column share of one epilogue warp
                rep -= 1                        # (fewer copies are used; the extra replicated rows are simply masked)
            flags = int(lay.relu) | (0 if last else 2) | (rep << 4)  # intermediates are stored TF32-rounded
            if into is not None:
                native.mlp_layer(B, lay.c_out, lay.c_out_pad, c1, c2, cols, lay.w, lay.shift, cur1, cur2, res, flags, p,
                                 out_ctot=into.shape[1], out_coff=out_coff)
            else:
                native.mlp_layer(B, lay.c_out, lay.c_out_pad, c1, c2, cols, lay.w, lay.shift, cur1, cur2, res, flags, p)
            cur1, cur2 = res, None
        return cur1


class FoldedHeads(FoldedMLP):
    """Several per-point heads of the same depth over the same input (the RPN's classification and regression heads,
    lib/net/rpn.py:31-45) as ONE chain of launches: the first layers are stacked (they read the same features), the
    deeper ones are block-diagonal.  Output channels are the heads' outputs in order."""

    def __init__(self, heads):
        self._heads = [nn.Sequential(*[m for m in h if not isinstance(m, nn.Dropout)]) for h in heads]
        assert len({len(h) for h in self._heads}) == 1, "heads must have the same number of layers"
        self.out_channels = [h[-1].conv.out_channels for h in self._heads]
        super().__init__(nn.Sequential(*[b for h in self._heads for b in h]))   # (parameters for the version stamp)

    def _matrices(self):
        per_head = [[self.fold_block(b) for b in h] for h in self._heads]
        out = []
        for li in range(len(per_head[0])):
            ws, shifts, relus = zip(*[ph[li] for ph in per_head])
            assert len(set(relus)) == 1, "heads must agree on the activation of every layer"
            w = torch.cat(ws, dim=0) if li == 0 else torch.block_diag(*ws)
            out.append((w, torch.cat(shifts), relus[0]))
        return out


def supported(cols: int, pool: int) -> bool:
    """Shapes the kernel accepts: 16-byte aligned rows; pooling groups (nsample) that are a power of two <= 128, so
    that a group never straddles the column share of one epilogue warp."""
    return cols % 4 == 0 and (pool == 0 or (pool & (pool - 1) == 0 and pool <= 128 and cols % pool == 0))


class FusedSAScale:
    """Folded weights of a 3-layer SharedMLP in the layout of `ws3d_sa_mlp_fused` (csrc/sa_fused.cu): one kernel per
    set-abstraction scale does grouping + the three layers + the max-pool with every activation kept in TMEM.
    Layer l's weight is (roundup(c_l, 16), 32 * ceil(K_l / 32)) with K_1 = roundup(3 + C, 8), K_{l+1} = roundup(c_l, 16)."""

    def __init__(self, mlp: nn.Sequential):
        self._mlp = mlp
        self._versions = None
        self._build()

    @staticmethod
    def eligible(mlp: nn.Sequential, c_feat: int, nsample: int) -> bool:
        blocks = list(mlp)
        if len(blocks) != 3 or not all(hasattr(b, "conv") and hasattr(b, "activation") for b in blocks):
            return False
        if any(not isinstance(b.activation, nn.ReLU) for b in blocks):
            return False   # the pooled epilogue orders non-negative floats as unsigned integers
        if any(hasattr(b, "bn") and not hasattr(b.bn, "bn") for b in blocks):
            return False
        if blocks[0].conv.in_channels != 3 + c_feat:
            return False
        widths = [b.conv.out_channels for b in blocks]
        return bool(native.sa_mlp_fused_supported(c_feat, nsample, *widths))

    def _stamp(self):
        return tuple((p.data_ptr(), p._version) for p in list(self._mlp.parameters()) + list(self._mlp.buffers()))

    def _build(self):
        self.w, self.shift, self.widths = [], [], []
        k_in = None
        for block in self._mlp:
            conv = block.conv
            w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
            shift = conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
            if hasattr(block, "bn"):
                bn = block.bn.bn
                scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
                w = w * scale[:, None]
                shift = (shift - bn.running_mean) * scale + bn.bias.detach()
            c_out, c_in = w.shape
            n_pad = _ceil(c_out, 16)
            k_pad = _ceil(_ceil(c_in, 8) if k_in is None else k_in, _CHUNK_K)
            wp = torch.zeros((n_pad, k_pad), dtype=torch.float32, device=w.device)
            wp[:c_out, :c_in] = _round_tf32(w)
            sp = torch.zeros(n_pad, dtype=torch.float32, device=w.device)
            sp[:c_out] = shift
            self.w.append(wp.contiguous())
            self.shift.append(sp)
            self.widths.append(c_out)
            k_in = n_pad
        self._versions = self._stamp()

    def __call__(self, xyz, new_xyz, features, idx, out, c_off, rows=None, out_pm=None, pm_xyz=False):
        """xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) or None, idx (B,M,K) int32; writes out[:, c_off:c_off+c3, :].
        `rows` (B,N,ld) = [xyz | features | zeros] point-major: gather from it instead (same result, 16-byte loads);
        `out_pm` (B,M,ld_pm): the pooled channels also land there point-major (the next level's `rows`)."""
        if self._versions != self._stamp():
            self._build()
        B, N, _ = xyz.shape
        M, K = idx.shape[1], idx.shape[2]
        c_feat = 0 if features is None else features.shape[1]
        if rows is not None:
            native.sa_mlp_fused_rows(B, N, M, K, c_feat, rows, new_xyz, idx, self.widths, self.w, self.shift, out, out.shape[1], c_off,
                                     out_pm, pm_xyz)
            return
        native.sa_mlp_fused(B, N, M, K, c_feat, xyz, new_xyz, features, idx, self.widths, self.w, self.shift, out,
                            out.shape[1], c_off)


class FoldedFPFirstLayer:
    """First layer of a feature-propagation MLP with the 1x1 convolution moved IN FRONT of the interpolation.

    PointnetFPModule computes relu(bn(W [interp(f_known) ; skip])) (pointnet2_modules.py:139-154).  The interpolation
    weights do not depend on the channel, so W_a interp(f) = interp(W_a f): the product is formed on the m known points
    (a quarter of the columns), the c_out-channel result is interpolated (instead of the c_in-channel input) and the
    skip term, shift and ReLU ride in the interpolation kernel's epilogue (`ws3d_three_interpolate_affine`).  At FP0
    (256 + 1 -> 128 channels, 4096 -> 16384 points) that replaces 0.94 GB of HBM traffic by 0.30 GB.  Served when the
    skip input has at most one channel (the intensity row of FP0) and c_out % 4 == 0."""

    def __init__(self, block: nn.Module, c_interp: int, c_skip: int):
        assert c_skip in (0, 1)
        self._block, self._dims, self._versions = block, (c_interp, c_skip), None
        self._build()

    @staticmethod
    def eligible(block: nn.Module, c_interp: int, c_skip: int, m: int, n: int) -> bool:
        conv = getattr(block, "conv", None)
        if conv is None or c_skip > 1 or conv.in_channels != c_interp + c_skip or conv.out_channels % 4:
            return False
        act = getattr(block, "activation", None)
        return (act is None or isinstance(act, nn.ReLU)) and m % 4 == 0 and 0 < m <= 8192 and n >= 256

    def _stamp(self):
        return tuple((p.data_ptr(), p._version) for p in list(self._block.parameters()) + list(self._block.buffers()))

    def _build(self):
        block, (c_interp, c_skip) = self._block, self._dims
        conv = block.conv
        w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
        shift = conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
        if hasattr(block, "bn"):
            bn = block.bn.bn
            scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
            w = w * scale[:, None]
            shift = (shift - bn.running_mean) * scale + bn.bias.detach()
        self.c_out = w.shape[0]
        self.c_out_pad = _ceil(self.c_out, _TILE_M)
        wa = torch.zeros((self.c_out_pad, _ceil(c_interp, _CHUNK_K)), dtype=torch.float32, device=w.device)
        wa[:self.c_out, :c_interp] = _round_tf32(w[:, :c_interp])
        self.wa = wa.contiguous()
        self.zero_shift = torch.zeros(self.c_out_pad, dtype=torch.float32, device=w.device)
        self.scale1 = w[:, c_interp].contiguous() if c_skip else None
        self.shift = shift.contiguous()
        self.relu = hasattr(block, "activation")
        self._versions = self._stamp()

    def __call__(self, known_feats, skip, idx, weight, n: int, round_out: bool):
        """known_feats (B, c_interp, m), skip (B, 1, n) or None, idx / weight (B, n, 3) -> (B, c_out, n)."""
        if self._versions != self._stamp():
            self._build()
        B, c_interp, m = known_feats.shape
        pre = torch.empty((B, self.c_out, m), dtype=torch.float32, device=known_feats.device)
        native.mlp_layer(B, self.c_out, self.c_out_pad, c_interp, 0, m, self.wa, self.zero_shift, known_feats, None, pre, 0, 0)
        out = torch.empty((B, self.c_out, n), dtype=torch.float32, device=known_feats.device)
        row1 = None if skip is None else skip.reshape(B, n)
        native.three_interpolate_affine(B, self.c_out, m, n, pre, idx, weight, self.scale1 if skip is not None else None, row1,
                                        self.shift, int(self.relu) | (2 if round_out else 0), out)
        return out


class FoldedSAFirstLayer:
    """First layer of a set-abstraction MLP with the feature part of the 1x1 convolution moved IN FRONT of the grouping.

    W [xyz[i] - centre ; f[i]] = (W_f f)[i] + W_x (xyz[i] - centre): P = W_f f is one tensor-core launch over the N
    source points (4-8x fewer columns than npoint * nsample), and `ws3d_group_affine` gathers it, adds the coordinate
    term in FP32, the shift and the ReLU and writes the layer's OUTPUT -- the (3 + C)-channel grouped tensor and the big
    first GEMM disappear.  Used for the scales whose weights do not fit the one-kernel path (SA3, SA4)."""

    def __init__(self, block: nn.Module, c_feat: int):
        self._block, self._c_feat, self._versions = block, c_feat, None
        self._build()

    @staticmethod
    def eligible(block: nn.Module, c_feat: int, n: int) -> bool:
        conv = getattr(block, "conv", None)
        act = getattr(block, "activation", None)
        return (conv is not None and c_feat > 0 and conv.in_channels == 3 + c_feat and n % 4 == 0
                and (act is None or isinstance(act, nn.ReLU)))

    def _stamp(self):
        return tuple((p.data_ptr(), p._version) for p in list(self._block.parameters()) + list(self._block.buffers()))

    def _build(self):
        block, c_feat = self._block, self._c_feat
        conv = block.conv
        w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
        shift = conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
        if hasattr(block, "bn"):
            bn = block.bn.bn
            scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
            w = w * scale[:, None]
            shift = (shift - bn.running_mean) * scale + bn.bias.detach()
        self.c_out = w.shape[0]
        self.c_out_pad = _ceil(self.c_out, _TILE_M)
        wf = torch.zeros((self.c_out_pad, _ceil(c_feat, _CHUNK_K)), dtype=torch.float32, device=w.device)
        wf[:self.c_out, :c_feat] = _round_tf32(w[:, 3:])
        self.wf = wf.contiguous()
        self.zero_shift = torch.zeros(self.c_out_pad, dtype=torch.float32, device=w.device)
        self.wx = w[:, :3].contiguous()
        self.shift = shift.contiguous()
        self.relu = hasattr(block, "activation")
        self._versions = self._stamp()

    def __call__(self, xyz, new_xyz, features, idx, round_out: bool):
        """xyz (B,N,3), new_xyz (B,M,3), features (B,C,N), idx (B,M,K) -> first-layer output (B, c_out, M*K)."""
        if self._versions != self._stamp():
            self._build()
        B, C, N = features.shape
        M, K = idx.shape[1], idx.shape[2]
        pre = torch.empty((B, self.c_out, N), dtype=torch.float32, device=features.device)
        native.mlp_layer(B, self.c_out, self.c_out_pad, C, 0, N, self.wf, self.zero_shift, features, None, pre, 0, 0)
        out = torch.empty((B, self.c_out, M * K), dtype=torch.float32, device=features.device)
        native.group_affine(B, N, M, self.c_out, K, pre, xyz, new_xyz, self.wx, self.shift, idx,
                            int(self.relu) | (2 if round_out else 0), out)
        return out
