#!/usr/bin/env python
"""A/B of two library builds on the bench's parity frame (GOP 0, frame 4): prints bits and symbol digests.
   B200VC_LIB=<other .so> python tools/ab_parity.py   -- compare the printed lines of two runs."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-compression_b200")):
    sys.path.insert(0, p)
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = False
import bench
from b200vc import synthetic
from b200vc.lhbdc import reflect_pad64
dev = torch.device("cuda", 0)
model = bench.build_lhbdc(dev)
frames = reflect_pad64(synthetic.make_sequence(9, 1080, 1920, seed=1234).to(dev))
with torch.no_grad():
    x_hat, bits, p = model.forward_device(frames[0:1], frames[4:5], frames[8:9], return_parts=True)
dig = lambda t: hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:16]
print("lib", os.environ.get("B200VC_LIB", "default"), "bits %.6f" % bits.item(), "mv_y", dig(p["mv"]["y_symbols"]),
      "res_y", dig(p["res"]["y_symbols"]), "x_hat", dig(x_hat))
