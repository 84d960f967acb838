"""Where do this library's training layers and the TF32-emulating autograd restatement (tests/refmods.py) part ways?
Runs one Stage-1 training forward / backward on both, compares every module's output and the gradient arriving at it.

    python tools/train_debug.py [scenes]
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refmods import emulated_shared_mlp_train  # noqa: E402

from ws3d_b200 import label_utils, models, synth, train_functions, train_mlp  # noqa: E402

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if os.environ.get("DBG_ONE_STREAM"):
    os.environ["WS3D_TWO_STREAMS"] = "0"


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm()) / max(1e-30, float(b.norm()))


def run(model, emulate, product_dtype=None):
    rec = {"fwd": {}, "grad": {}}
    calls = []
    orig = train_mlp.shared_mlp_train
    fn = (lambda m, a, b=None, pool=0: emulated_shared_mlp_train(m, a, b, pool=pool, product_dtype=product_dtype)) if emulate else orig

    def wrapped(mlp, x1, x2=None, pool=0):
        k = len(calls)
        calls.append((mlp, x1.detach().clone(), None if x2 is None else x2.detach().clone(), pool))
        if x1.requires_grad:
            x1.register_hook(lambda g, k=k: rec["grad"].__setitem__(f"mlp{k:02d}.dx1", g.detach().clone()))
        if x2 is not None and x2.requires_grad:
            x2.register_hook(lambda g, k=k: rec["grad"].__setitem__(f"mlp{k:02d}.dx2", g.detach().clone()))
        out = fn(mlp, x1, x2, pool=pool)
        rec["fwd"][f"mlp{k:02d}.out c{out.shape[1]} pool{pool}"] = out.detach().clone()
        out.register_hook(lambda g, k=k: rec["grad"].__setitem__(f"mlp{k:02d}.dout", g.detach().clone()))
        return out

    train_mlp.shared_mlp_train = wrapped
    try:
        torch.manual_seed(3)
        out = model({"pts_input": pts})
        loss, _ = train_functions.get_rpn_loss(out["rpn_cls"], out["rpn_reg"], cls_label, reg_label)
        loss.backward()
    finally:
        train_mlp.shared_mlp_train = orig
    torch.cuda.synchronize()
    rec["calls"] = calls
    return float(loss.detach()), rec


torch.manual_seed(0)
net = models.RPN().to(dev).train()
ref = copy.deepcopy(net)
ref32 = copy.deepcopy(net)
pts = torch.from_numpy(synth.make_batch(B, 16384)).to(dev)
gt, cnt = synth.make_gt_boxes(B)
cls_label, reg_label = label_utils.generate_gaussian_training_labels(pts[..., :3].contiguous(), torch.from_numpy(gt).to(dev),
                                                                     torch.from_numpy(cnt).to(dev))
l1, r1 = run(net, False)
l2, r2 = run(ref, True)
print("loss", l1, l2)
for k in r1["fwd"]:
    print(f"fwd  {k:32s} {rel(r1['fwd'][k], r2['fwd'][k]):.3e}")
for k in sorted(r1["grad"]):
    if k in r2["grad"]:
        print(f"grad {k:32s} {rel(r1['grad'][k], r2['grad'][k]):.3e}   |ref| {float(r2['grad'][k].norm()):.3e}")
errs = {}
for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
    if q.grad is not None and float(q.grad.abs().max()) > 1e-6:
        errs[n] = rel(p.grad, q.grad)
for n, e in errs.items():
    print(f"param {n:60s} {e:.3e}")


l3, r3 = run(ref32, True, torch.float32)
floor, mine = {}, {}
for (n, p), (_, q), (_, q32) in zip(net.named_parameters(), ref.named_parameters(), ref32.named_parameters()):
    if q.grad is not None and float(q.grad.abs().max()) > 1e-6:
        floor[n], mine[n] = rel(q32.grad, q.grad), rel(p.grad, q.grad)
med = lambda d: sorted(d.values())[len(d) // 2]
print(f"NOISE FLOOR (emulation FP32 products vs FP64 products): loss {l3} median {med(floor):.3e} worst {max(floor.values()):.3e}")
print(f"THIS LIBRARY vs FP64 products:                          loss {l1} median {med(mine):.3e} worst {max(mine.values()):.3e}")
for n in floor:
    print(f"   {n:60s} floor {floor[n]:.3e}  mine {mine[n]:.3e}")
print("---- every shared MLP in isolation on its ACTUAL input, (a) random dout (b) the dout of the step ----")
ref_calls = r2["calls"]
for k, (mlp, x1, x2, pool) in enumerate(r1["calls"]):
    rmlp = ref_calls[k][0]
    for tag in ("rand", "step"):
        xa, xb = x1.clone().requires_grad_(True), x1.clone().requires_grad_(True)
        x2a = None if x2 is None else x2.clone().requires_grad_(True)
        x2b = None if x2 is None else x2.clone().requires_grad_(True)
        torch.manual_seed(11)
        oa = train_mlp.shared_mlp_train(mlp, xa, x2a, pool=pool)
        torch.manual_seed(11)
        ob = emulated_shared_mlp_train(rmlp, xb, x2b, pool=pool)
        key = [n for n in r2["grad"] if n.startswith(f"mlp{k:02d}.dout")]
        g = torch.randn_like(ob) if tag == "rand" or not key else r2["grad"][key[0]]
        for q in list(mlp.parameters()) + list(rmlp.parameters()):
            q.grad = None
        (oa * g).sum().backward()
        (ob * g).sum().backward()
        d = (xa.grad - xb.grad).double()
        chan = d.mean(dim=(0, 2), keepdim=True).expand_as(d)
        perr = max(rel(p.grad, q.grad) for p, q in zip(mlp.parameters(), rmlp.parameters()) if q.grad is not None)
        print(f"mlp{k:02d} {tag}: c_in {x1.shape[1]}+{0 if x2 is None else x2.shape[1]} cols {x1.shape[2]} pool {pool}  fwd {rel(oa, ob):.2e}  dx1 {rel(xa.grad, xb.grad):.2e} "
              f"(per-channel-constant part {float(chan.norm()) / max(1e-30, float(xb.grad.double().norm())):.2e})  worst param {perr:.2e}"
              + ("" if x2 is None else f"  dx2 {rel(x2a.grad, x2b.grad):.2e}"))
