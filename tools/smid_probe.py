"""Is the spread of CTA finish times in a saturating weight stream tied to the SM (systematic) or random?"""
# NOTE: needs a debug build of the library: PC_NVCC_EXTRA=-DPC_DEBUG python tools/smid_probe.py  (product builds compile the env / trace hooks out)
import os, sys, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptcache_b200 import _native as nat
DEV = "cuda"
ws = nat.linear_workspace(DEV)
q, k, n = 16, 4096, 22016
W = (torch.randn(4, n, k, device=DEV) * 0.02).half()
WT = [nat.tile_weights(W[i], nat.PC_LIN_SILU_MUL) for i in range(4)]
x = torch.randn(q, k, device=DEV).half()
out = torch.empty(q, n // 2, dtype=torch.half, device=DEV)
tr = torch.zeros(8 * 148, dtype=torch.int64, device=DEV)
os.environ["PC_LIN_TRACE_PTR"] = hex(tr.data_ptr())
runs = []
for i in range(12):
    tr.zero_()
    nat.linear_small(x, W[i % 4], out, ws, silu_mul=True, w_tiled=WT[i % 4])
    torch.cuda.synchronize()
    t = tr.cpu().numpy().reshape(148, 8).astype(np.float64)
    t0 = t[:, 0].min()
    fin = np.zeros(148)
    fin[t[:, 7].astype(int)] = (t[:, 3] - t0) / 1000  # last MMA committed, indexed by SM id
    runs.append(fin)
runs = np.array(runs[4:])
mean = runs.mean(0)
print("per-run spread (us): ", [f"{r.min():.1f}-{r.max():.1f}" for r in runs])
print("mean over runs per SM: min %.1f max %.1f std %.2f; run-to-run std per SM (mean) %.2f" % (mean.min(), mean.max(), mean.std(), runs.std(0).mean()))
cc = np.corrcoef(runs)
print("mean pairwise correlation of per-SM finish times between runs: %.2f" % ((cc.sum() - len(cc)) / (len(cc) * (len(cc) - 1))))
order = np.argsort(mean)
print("fastest SMs", order[:12].tolist(), "slowest SMs", order[-12:].tolist())
