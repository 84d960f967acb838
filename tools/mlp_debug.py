"""Layout diagnostics for ws3d_mlp_layer (run on the GPU box): structured inputs whose outputs
reveal which operand index went wrong, then random cases with error statistics."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ws3d_b200 import native

dev = "cuda:0"


def run(w, shift, x1, x2, relu, pool, c_out):
    B, c1, cols = x1.shape
    c2 = 0 if x2 is None else x2.shape[1]
    c_out_pad = w.shape[0]
    out = torch.full((B, c_out, cols // pool if pool else cols), float("nan"), device=dev)
    native.mlp_layer(B, c_out, c_out_pad, c1, c2, cols, w, shift, x1, x2, out, relu, pool)
    torch.cuda.synchronize()
    return out


def pad_w(w, c1, c2):
    c_out, c_in = w.shape
    k1 = (c1 + 31) // 32 * 32
    k2 = (c2 + 31) // 32 * 32 if c2 else 0
    wp = torch.zeros(((c_out + 127) // 128 * 128, k1 + k2), device=dev)
    wp[:c_out, :c1] = w[:, :c1]
    if c2:
        wp[:c_out, k1:k1 + c2] = w[:, c1:]
    return wp.contiguous()


def case(c_in, c_out, cols, B=1, c2=0, relu=False, pool=0, exact_ints=False, tag=""):
    torch.manual_seed(c_in * 131 + c_out * 7 + cols)
    c1 = c_in - c2
    if exact_ints:
        # small integers: exact in TF32, so any mismatch is an indexing error, not rounding
        w = torch.randint(-3, 4, (c_out, c_in), device=dev).float()
        x = torch.randint(-3, 4, (B, c_in, cols), device=dev).float()
        shift = torch.randint(-2, 3, (c_out,), device=dev).float()
    else:
        w = torch.randn(c_out, c_in, device=dev)
        x = torch.randn(B, c_in, cols, device=dev)
        shift = torch.randn(c_out, device=dev)
    sp = torch.zeros((c_out + 127) // 128 * 128, device=dev)
    sp[:c_out] = shift
    x1 = x[:, :c1].contiguous()
    x2 = x[:, c1:].contiguous() if c2 else None
    want = torch.einsum("oc,bce->boe", w.double(), x.double()) + shift.double()[None, :, None]
    if relu:
        want = want.clamp_min(0)
    if pool:
        want = want.view(B, c_out, cols // pool, pool).amax(-1)
    got = run(pad_w(w, c1, c2), sp, x1, x2, relu, pool, c_out).double()
    err = (got - want).abs()
    scale = want.abs().max().item() + 1e-9
    bad = ~(err <= (0 if exact_ints else 4e-3 * scale))
    print(f"{tag or 'case'} c_in={c_in} (c2={c2}) c_out={c_out} cols={cols} B={B} relu={relu} pool={pool} ints={exact_ints}: "
          f"max_err={err.nan_to_num(1e30).max().item():.3e} rel={err.nan_to_num(1e30).max().item() / scale:.3e} "
          f"bad={int(bad.sum())}/{bad.numel()} nan={int(got.isnan().sum())}", flush=True)
    if bad.any() and exact_ints:
        idx = bad.nonzero()[:8]
        for b, o, e in idx.tolist():
            print(f"   [b={b} co={o} e={e}] got={got[b, o, e].item():.1f} want={want[b, o, e].item():.1f}")
        print("   bad per channel-quarter:", [int(bad[:, q * 32:(q + 1) * 32].sum()) for q in range((c_out + 31) // 32)][:8])
        cb = bad.any(dim=1)[0]
        print("   bad column blocks of 32:", [int(cb[i:i + 32].sum()) for i in range(0, min(cb.numel(), 512), 32)])
    return int(bad.sum()) == 0


def main():
    ok = True
    # identity probes first
    ok &= case(32, 128, 256, exact_ints=True, tag="one K chunk, one tile")
    ok &= case(8, 16, 256, exact_ints=True, tag="OOB K rows + masked channels")
    ok &= case(64, 128, 512, exact_ints=True, tag="two K chunks, two column tiles")
    ok &= case(160, 256, 1024, B=2, exact_ints=True, tag="5 K chunks (ring wraps), 2 M tiles, 2 clouds")
    ok &= case(99, 64, 768, B=2, exact_ints=True, relu=True, tag="ragged K, relu")
    ok &= case(99, 64, 1000, B=2, exact_ints=True, relu=True, tag="ragged columns")
    ok &= case(99, 64, 1024, B=2, exact_ints=True, relu=True, pool=32, tag="pool 32")
    ok &= case(99, 64, 1024 + 16, B=2, exact_ints=True, relu=True, pool=16, tag="pool 16 ragged tile")
    ok &= case(257, 128, 512, B=2, c2=1, exact_ints=True, tag="two inputs (256 + 1)")
    ok &= case(608, 256, 512, B=1, c2=96, exact_ints=True, tag="two inputs (512 + 96)")
    ok &= case(515, 512, 2048, B=2, relu=True, pool=32, tag="random tf32")
    ok &= case(1536, 512, 64, B=2, c2=512, relu=True, tag="random tf32 two inputs")
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
