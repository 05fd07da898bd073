"""Pipeline timeline of the two-threads-per-row TS attention kernel: clock64 stamps of CTA 0 over 8 consecutive KV tiles
(wiw_debug_attn_trace), printed as cycles relative to the first stamp.  Slots: 0/2 MMA warp saw P_t(j); 1/3 PV(t,j)+QK(t,j+1) issued;
4/8 primary saw S_t(j); 5/9 pass 1 done; 6/10 helper released; 7/11 primary's P stored; 12/13 helper's P stored (t = 0 / 1)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import lib as L, ops
lib = ctypes.CDLL(L.LIB_PATH)
imgs, s, c = 4, 9216, 320
qkv = torch.randn(imgs * s, 3 * c, device="cuda").half()
for _ in range(2):
    ops.attention_spatial(qkv, imgs, s, c)
torch.cuda.synchronize()
assert lib.wiw_debug_attn_trace(1, None) == 0
ops.attention_spatial(qkv, imgs, s, c)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (8 * 16))()
assert lib.wiw_debug_attn_trace(0, buf) == 0
v = [[buf[i * 16 + k] for k in range(16)] for i in range(8)]
t0 = min(x for row in v for x in row if x)
ts = os.environ.get("WIW_ATTN_TS", "4")
if ts == "2":
    names = ["mma:P0", "mma:iss0", "mma:P1", "mma:iss1", "pri0:S", "pri0:max", "hlp0:go", "pri0:P", "pri1:S", "pri1:max", "hlp1:go", "pri1:P", "hlp0:P", "hlp1:P"]
elif ts == "4":   # one Q tile, 128-key steps: MMA warp saw P(j) | issued PV(j) + QK(j+2) + commit | softmax saw S(j) | row max known | P(j) stored
    names = ["mma:P", "mma:iss", "smx:S", "smx:max", "smx:P"]
else:   # ts3: two 64-key steps (even | odd) per row; per step: MMA saw P0, issued 0, saw P1, issued 1, primary0 saw S, stored P, primary1 saw S, stored P
    names = [f"{e}:{n}" for e in ("e", "o") for n in ("mP0", "mI0", "mP1", "mI1", "p0S", "p0P", "p1S", "p1P")]
print("iter " + " ".join(f"{n:>8s}" for n in names))
for i, row in enumerate(v):
    print(f"{i:4d} " + " ".join(f"{(row[k] - t0) if row[k] else -1:8d}" for k in range(len(names))))
per = [(v[i + 1][0] - v[i][0]) for i in range(7)]
print("cycles per 128 keys (slot 0 to slot 0):", per)
