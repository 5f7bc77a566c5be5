"""ncu target: one ws launch of N filters x secs seconds (HG_FB_WS_NF picks the CTA width)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import huygens_b200 as hb
from huygens_b200 import design
N = int(sys.argv[1]); secs = float(sys.argv[2]); Ls = int(secs * 48000)
fwd, back, gains = design.resonator_bank(design.log_spaced(N), 0.999)
fb = hb.Filterbank(2, N, 0.001, 1.0, dtype=np.float32)
fb.coefficients(0, fwd, back); fb.mix(gains); fb.boost(np.ones(N))
x = torch.from_numpy(design.chirp(Ls, dtype=np.float32)).cuda()
out = torch.empty_like(x)
for _ in range(3):
    fb.process(x, out=out)
torch.cuda.synchronize()
print(fb.last_launch())
