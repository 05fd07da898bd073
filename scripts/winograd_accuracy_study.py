"""Numerical study (CPU, no kernels): error of a Winograd F(2x2,3x3) convolution with fp16 storage of the transformed
operands and fp32 accumulation, against the direct convolution with fp16 operands and fp32 accumulation (what conv_gemm does),
both measured against an fp64 reference.  Decides whether Winograd is admissible for the 3x3 convs of the UNet."""
import torch
import torch.nn.functional as F

torch.manual_seed(0)
N, C, Co, H, W = 2, 320, 320, 36, 64
x = F.silu(torch.randn(N, C, H, W, dtype=torch.float64))                 # post GroupNorm+SiLU statistics
w = torch.randn(Co, C, 3, 3, dtype=torch.float64) * (0.7 / (9 * C) ** 0.5)
ref = F.conv2d(x, w, padding=1)
xh, wh = x.half().double(), w.half().double()                             # fp16 storage of the operands
direct = F.conv2d(xh, wh, padding=1).float().half().double()             # fp32-ish accumulation, fp16 output

G = torch.tensor([[1, 0, 0], [0.5, 0.5, 0.5], [0.5, -0.5, 0.5], [0, 0, 1]], dtype=torch.float64)
Bt = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float64)
At = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float64)
U = torch.einsum("ij,ocjk,lk->ocil", G, wh, G).half().double()            # transformed weights stored fp16
xp = F.pad(xh, (1, 1, 1, 1))
tiles = xp.unfold(2, 4, 2).unfold(3, 4, 2)                                # [N, C, H/2, W/2, 4, 4]
V = torch.einsum("ij,nchwjk,lk->nchwil", Bt, tiles, Bt).half().double()   # transformed input stored fp16
M = torch.einsum("ocil,nchwil->nohwil", U, V)                             # 16 GEMMs, fp32-ish accumulation
Y = torch.einsum("ij,nohwjk,lk->nohwil", At, M, At)                       # [N, Co, H/2, W/2, 2, 2]
wino = Y.permute(0, 1, 2, 4, 3, 5).reshape(N, Co, H, W).float().half().double()


def rel(a):
    return float(((a - ref).pow(2).mean() / ref.pow(2).mean()).sqrt())


print(f"direct (fp16 operands, fp16 output): rel rms {rel(direct):.3e}")
print(f"winograd F(2x2,3x3), fp16 transformed operands: rel rms {rel(wino):.3e}  ({rel(wino) / rel(direct):.1f}x the direct error)")
print(f"max |transformed input| / max |input| = {float(V.abs().max() / xh.abs().max()):.2f}")
