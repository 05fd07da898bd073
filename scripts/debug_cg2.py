import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops
DEV = "cuda"
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
M, K, N = 256, 64, 64
A = torch.zeros(M, K)
for m in range(M):
    A[m, m % 64] = 1.0
B = torch.zeros(N, K)
for n in range(N):
    for k in range(K):
        B[n, k] = n * 8 + (k % 8) + (k // 8) * 0.125      # exactly representable in fp16
gw = ops.prep_linear(B, None, DEV)
y = ops.linear(A.half().to(DEV), gw).float().cpu()
ref = A.half().float() @ B.half().float().t()
bad = (y - ref).abs() > 1e-3
print("cg", os.environ.get("WIW_GEMM_CG"), "bn", gw.bn, "bad total", int(bad.sum()))
for r0 in (0, 128):
    for c0 in (0, 32):
        print(f"rows {r0}-{r0+127} cols {c0}-{c0+31}: bad {int(bad[r0:r0+128, c0:c0+32].sum())}")
print("y[0:4, 0:8]\n", y[0:4, 0:8]); print("ref[0:4,0:8]\n", ref[0:4, 0:8])
print("y[128:132, 0:8]\n", y[128:132, 0:8]); print("ref[128:132,0:8]\n", ref[128:132, 0:8])
print("y[0:4, 32:40]\n", y[0:4, 32:40]); print("ref[0:4,32:40]\n", ref[0:4, 32:40])
print("y[128:132, 32:40]\n", y[128:132, 32:40])
# second: random A to see magnitude relation
A2 = torch.randn(M, K); y2 = ops.linear(A2.half().to(DEV), gw).float().cpu(); ref2 = A2.half().float() @ B.half().float().t()
print("random A: rel err per quadrant", [float(((y2-ref2)[r0:r0+128, c0:c0+32]).abs().max() / ref2.abs().max()) for r0 in (0,128) for c0 in (0,32)])
