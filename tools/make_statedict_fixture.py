"""Writes tests/golden/rpn_state_dict_keys.json: parameter/buffer names and shapes of the reference's
Stage-1 RPN (lib/net/rpn.py + lib/net/pointnet2_msg.py), built by importing the reference modules
from /root/reference with stub extension modules (no kernels run).  Run in the authoring container."""
import json
import os
import sys
import types

REF = "/root/reference"
sys.path[:0] = [REF, os.path.join(REF, "lib", "net")]
for name in ("pointnet2_cuda", "iou3d_cuda", "roipool3d_cuda"):
    sys.modules[name] = types.ModuleType(name)


class EasyDict(dict):  # minimal stand-in for the missing `easydict` package
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in (d or {}).items():
            setattr(self, k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        self[k] = v


ed = types.ModuleType("easydict")
ed.EasyDict = EasyDict
sys.modules["easydict"] = ed

from lib.config import cfg  # noqa: E402

cfg.RPN.LOC_SCOPE, cfg.RPN.LOC_BIN_SIZE = 4.0, 0.8          # tools/cfgs/weaklyRPN.yaml:37-38
cfg.RPN.LOSS_CLS, cfg.RPN.FOCAL_GAMMA = "SigmoidFocalLoss", 2.0
from lib.net.rpn import RPN  # noqa: E402

m = RPN()
keys = {k: list(v.shape) for k, v in m.state_dict().items()}
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rpn_state_dict_keys.json")
json.dump(keys, open(out, "w"), indent=0)
print(len(keys), "entries,", sum(p.numel() for p in m.parameters()), "parameters ->", out)
