"""sha256 of encode_images on a seeded full-size model / input (bit-level A/B of two library builds: run once per
RADVLM_B200_LIB and compare the lines)."""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import golden_inputs as gi  # noqa: E402
from radvlm_b200 import synthetic  # noqa: E402

host = synthetic.build_host(hidden_size=3584, vocab=64, seed=gi.FULL_SEED, dtype=torch.bfloat16, device="cuda")
for n in (1, 10, 37):
    x = gi.encoder_pixels(n, seed=40 + n).cuda().bfloat16()
    f = host.encode_images(x)
    torch.cuda.synchronize()
    print("tiles %d sha256 %s" % (n, hashlib.sha256(f.view(torch.int16).cpu().numpy().tobytes()).hexdigest()))
