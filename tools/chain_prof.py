"""Debug: per-role cycle counters of the warp-specialised kernel (HG_FB_PROF=1), see hg_fb_debug_profile."""
import os, sys
os.environ["HG_FB_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import huygens_b200 as hb
from huygens_b200 import design, _lib as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 10
mode = sys.argv[3] if len(sys.argv) > 3 else "ordered"
Ls = int(secs * 48000)
freqs = design.log_spaced(N)
fwd, back, gains = design.resonator_bank(freqs, 0.999)
fb = hb.Filterbank(2, N, 0.001, 1.0, dtype=np.float32, sum_mode=mode)
fb.coefficients(0, fwd, back); fb.mix(gains); fb.boost(np.ones(N))
x = torch.from_numpy(design.chirp(Ls, dtype=np.float32)).cuda()
out = torch.empty_like(x)
for _ in range(3):
    fb.process(x, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fb.process(x, out=out); e1.record(); torch.cuda.synchronize()
info = fb.last_launch()
G = info["ctas"]
buf = np.zeros((G, 8), np.int64)
L.check(L.lib().hg_fb_debug_profile(fb._h, buf.ctypes.data_as(C.c_void_p), G))
tile = {0: 128, 1: 64, 2: 32}
ms = e0.elapsed_time(e1)
print(f"N={N} L={Ls} sum={mode} ctas={G} kernel {ms:.3f} ms = {ms*1e-3*1.965e9/Ls:.2f} cycles/sample")
names = ["S wait y tile", "S wait running row", "S mixdown", "S reserve+send", "R wait u", "R wait slot (gated by S)", "R total"]
if info["filters_per_thread"] == -2:
    names = ["S1 wait running row", "S1 wait y tiles", "S1 mixdown", "S1 hand-over", "RU0 gated by S", "RU0 wait x", "RUlast gated by S", "RUlast total"]
print(info)
for i, n in enumerate(names):
    col = buf[:, i].astype(np.float64)
    print(f"  {n:28s} per sample: mean {col.mean()/Ls:8.2f}  first {col[0]/Ls:8.2f}  mid {col[G//2]/Ls:8.2f}  last {col[-1]/Ls:8.2f}  max {col.max()/Ls:8.2f}")
if os.environ.get("HG_PROF_TABLE"):
    print("cta: " + " ".join(f"{n[:10]:>10s}" for n in names))
    step = max(1, G // 40)
    rows = sorted(set(list(range(0, G, step)) + list(range(max(0, G // 2 - 12), min(G, G // 2 + 12)))))
    for c in rows:
        print(f"{c:4d} " + " ".join(f"{buf[c, i] / Ls:10.2f}" for i in range(len(names))))
