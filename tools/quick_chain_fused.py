"""Quick A/B timing of the whole-chain kernels on the GPU box (not a bench line):
   python tools/quick_chain_fused.py [B ...]   -> ms/step of the compile-time and the run-time kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exposure_b200.chain import FusedFilterChain   # noqa: E402
from oracle import filters as F                    # noqa: E402

CHAIN = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]
for B in [int(a) for a in sys.argv[1:]] or [64, 256]:
  x = F.synth_images(2, 512, 512, seed=1).cuda().repeat(B // 2, 1, 1, 1).contiguous()
  gout = torch.randn_like(x)
  for static in (True, False):
    fused = FusedFilterChain(CHAIN, B, torch.device("cuda"), static_chain=static)
    fused.set_logits([(F.synth_logits(f, B, seed=9) * 0.5).cuda() for f in CHAIN])
    y, gx = torch.empty_like(x), torch.empty_like(x)
    for _ in range(5):
      fused.forward_backward(x, gout, y_out=y, gx_out=gx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
      fused.forward_backward(x, gout, y_out=y, gx_out=gx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print("B=%d %s chain8 fwd+bwd: %.4f ms/step, %.0f images/s, algorithmic(60N) %.0f GB/s, actual(48B/px) %.0f GB/s"
          % (B, "compile-time" if static else "run-time   ", ms, B / ms * 1e3, B * 512 * 512 * 480 / ms / 1e6,
             B * 512 * 512 * 48 / ms / 1e6))
