"""Debug: progress of the chain over time — when S warp 1 of every CTA finished its tiles (HG_PROF_G build)."""
import os, sys
os.environ["HG_FB_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import huygens_b200 as hb
from huygens_b200 import design, _lib as L
N = int(sys.argv[1]); secs = float(sys.argv[2]); Ls = int(secs * 48000)
freqs = design.log_spaced(N)
fwd, back, gains = design.resonator_bank(freqs, 0.999)
fb = hb.Filterbank(2, N, 0.001, 1.0, dtype=np.float32)
fb.coefficients(0, fwd, back); fb.mix(gains); fb.boost(np.ones(N))
x = torch.from_numpy(design.chirp(Ls, dtype=np.float32)).cuda()
out = torch.empty_like(x)
for _ in range(2):
    fb.process(x, out=out)
torch.cuda.synchronize()
fb.process(x, out=out); torch.cuda.synchronize()
G = fb.last_launch()["ctas"]
buf = np.zeros((4096, 8), np.int64)
L.check(L.lib().hg_fb_debug_profile(fb._h, buf.ctypes.data_as(C.c_void_p), 4096))
st = buf[1024:1024 + 16 * G].reshape(G, 128)
t0 = st[st > 0].min()
NS = 7
print("columns: time (us) at which S warp 1 finished its tile number m*32 (tile index k = 1 + NS*32*m); rows: CTA")
cols = list(range(0, 66, 6))
print("cta  " + " ".join(f"m={m:3d}  " for m in cols))
rows = sorted(set(list(range(0, G, 8)) + [G - 1] + list(range(56, min(G, 90)))))
for c in rows:
    print(f"{c:4d} sm{buf[c, 5] // Ls:4d} " + " ".join(f"{(st[c, m] - t0) / 1e3:8.1f}" if st[c, m] > 0 else "       -" for m in cols))
