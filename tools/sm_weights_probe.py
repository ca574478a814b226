"""Does a stream-K partition weighted by per-SM streaming speed remove the finish-time spread?  (experiment, debug build:
   PC_NVCC_EXTRA="-DPC_DEBUG -DPC_LIN_BOUNDS" python tools/sm_weights_probe.py)  gate/up+SiLU 22016x4096, q=16, hybrid partition: every CTA
   owns one whole tile (32 units) + a share of the 24-tile remainder (768 units): uniform (5.19 each) vs weighted."""
import os, sys, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptcache_b200 import _native as nat
DEV = "cuda"
ws = nat.linear_workspace(DEV)
q, k, n = 16, 4096, 22016
NW = 4
W = (torch.randn(NW, n, k, device=DEV) * 0.02).half()
WT = [nat.tile_weights(W[i], nat.PC_LIN_SILU_MUL) for i in range(NW)]
x = torch.randn(q, k, device=DEV).half()
out = torch.empty(q, n // 2, dtype=torch.half, device=DEV)
tr = torch.zeros(8 * 148, dtype=torch.int64, device=DEV)
WHOLE, R, G = 32, 768, 148


def traced(runs=8):
    os.environ["PC_LIN_TRACE_PTR"] = hex(tr.data_ptr())
    fins = []
    for i in range(runs + 2):
        tr.zero_()
        nat.linear_small(x, W[i % NW], out, ws, silu_mul=True, w_tiled=WT[i % NW])
        torch.cuda.synchronize()
        t = tr.cpu().numpy().reshape(148, 8).astype(np.float64)
        fin = np.zeros(148)
        fin[t[:, 7].astype(int)] = (t[:, 3] - t[:, 0].min()) / 1000
        assert len(set(t[:, 7].astype(int).tolist())) == 148, "two CTAs on one SM"
        fins.append(fin)
        ends.append(((t[:, 6].max() - t[:, 0].min()) / 1000, (t[:, 4].max() - t[:, 0].min()) / 1000, (t[:, 0].max() - t[:, 0].min()) / 1000))
    del os.environ["PC_LIN_TRACE_PTR"]
    rel = (t[:, :7] - t[:, 0].min()) / 1000
    top = np.argsort(-rel[:, 6])[:5].tolist() + np.argsort(rel[:, 6])[:3].tolist()
    for c in top:
        print(f"      cta {c:3d} sm {int(t[c, 7]):3d}: start {rel[c, 0]:.1f} setup {rel[c, 1]:.1f} first-mma {rel[c, 2]:.1f} last-mma {rel[c, 3]:.1f} "
              f"last-epi-start {rel[c, 4]:.1f} last-partial-arrival {rel[c, 5]:.1f} end {rel[c, 6]:.1f}")
    e = np.array(ends[-runs:]).mean(0)
    print(f"   trace: last epilogue end {e[0]:.1f} us, last first-epilogue {e[1]:.1f} us, CTA start spread {e[2]:.1f} us")
    return np.array(fins[2:]).mean(0)


ends = []


def single(reps=12):
    ts = []
    for i in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        nat.linear_small(x, W[i % NW], out, ws, silu_mul=True, w_tiled=WT[i % NW])
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def timed(reps=24):
    for i in range(4):
        nat.linear_small(x, W[i % NW], out, ws, silu_mul=True, w_tiled=WT[i % NW])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(reps):
        nat.linear_small(x, W[i % NW], out, ws, silu_mul=True, w_tiled=WT[i % NW])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def bounds_for(speed):
    lo, hi = 0.0, 1e3
    for _ in range(60):
        T = (lo + hi) / 2
        if np.maximum(1, T * speed - WHOLE).sum() > R: hi = T
        else: lo = T
    # every partition keeps at least one unit: the stream-K bookkeeping counts every CTA between a tile's first and last
    # contributor as a contributor
    r = np.maximum(1, np.floor(lo * speed - WHOLE)).astype(np.int64)
    order = np.argsort(-(lo * speed - WHOLE - r))
    for j in range(int(R - r.sum())):  # hand the rounding remainder to the partitions that were rounded down the most
        r[order[j % G]] += 1
    while r.sum() > R:
        r[np.argmax(r)] -= 1
    b = np.concatenate([[0], np.cumsum(r)]).astype(np.int64)
    assert b[-1] == R and (np.diff(b) >= 1).all()
    return b, np.diff(b)


ref = out.clone()
fin = traced()
units = np.full(148, WHOLE + R / G)
print(f"uniform: finish {fin.min():.1f}-{fin.max():.1f} us (mean {fin.mean():.1f}), kernel back-to-back {timed():.2f} us, alone {single():.2f} us")
nat.linear_small(x, W[0], ref, ws, silu_mul=True, w_tiled=WT[0])
for it in range(1):
    speed = units / np.maximum(fin - 1.5, 1.0)  # ~1.5 us from kernel start to the first stage
    b, r = bounds_for(speed)
    bt = torch.tensor(b, dtype=torch.int64, device=DEV)
    os.environ["PC_LIN_BOUNDS_PTR"] = hex(bt.data_ptr())
    units = WHOLE + r.astype(np.float64)
    fin = traced()
    t = timed()
    got = torch.empty_like(ref)
    nat.linear_small(x, W[0], got, ws, silu_mul=True, w_tiled=WT[0])
    torch.cuda.synchronize()
    print(f"weighted #{it}: remainder units per SM {r.min()}-{r.max()}, finish {fin.min():.1f}-{fin.max():.1f} us (mean {fin.mean():.1f}), "
          f"kernel back-to-back {t:.2f} us, alone {single():.2f} us, max |diff| vs uniform {float((got.float() - ref.float()).abs().max()):.3e}")
del os.environ["PC_LIN_BOUNDS_PTR"]
print(f"uniform again: kernel {timed():.2f} us")
