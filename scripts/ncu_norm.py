"""Targeted launches for `ncu --set full` of the HBM-bound norm kernels at the north-star level-0 / level-1 shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops
DEV = "cuda"
for (imgs, s, c) in [(28, 9216, 320), (28, 2304, 640)]:
    x = torch.randn(imgs * s, c, device=DEV).half()
    out = torch.empty_like(x)
    nw = ops.prep_norm(torch.ones(c), torch.zeros(c), DEV)
    for _ in range(2):
        ops.layer_norm(x, nw, out=out)
        ops.group_norm(x, nw, imgs, 1e-5, True, out=out)
    cat = torch.empty(imgs * s, 2 * c, device=DEV, dtype=torch.half)
    ops.concat_channels(x.view(imgs, s, c), out.view(imgs, s, c))
torch.cuda.synchronize()
print("done")
