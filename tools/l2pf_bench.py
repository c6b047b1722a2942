"""Megakernel decode step vs L2-prefetch distance / cap (B=8, large-v3 shape, random xkv)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from crisperwhisper_b200 import weights as Wt, _lib as L
from crisperwhisper_b200.engine import Engine
B = int(os.environ.get("B", "8")); T = int(os.environ.get("T", "128"))
eng = Engine(0)
cfg = Wt.large_v3_config()
eng.load_weights(Wt.synthetic_weights(cfg, eng.device, seed=0))
xkv = (torch.randn(32, B, 1500, 2, 20, 64, device="cuda") * 0.5).to(torch.bfloat16)
prompt = torch.tensor([[50258, 50259, 50360]] * B, dtype=torch.int32, device="cuda")
ref = None
for dist, mb in ((0, 64), (1, 64), (2, 64), (3, 64), (1, 32), (2, 96), (1, 16)):
    os.environ["CW_MEGA_L2PF"] = str(dist); os.environ["CW_MEGA_L2PF_MB"] = str(mb)
    flags = L.CW_DEC_SUPPRESS_EOS
    for _ in range(2):
        eng.decode(xkv, prompt, T, flags=flags, want_align=True)
    eng.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = eng.decode(xkv, prompt, T, flags=flags, want_align=True)
    eng.sync(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tok = out["tokens"].cpu()
    if ref is None: ref = tok
    print(f"l2pf dist={dist} cap={mb}MB  {1000*dt/(T+2):.3f} ms/step  same_tokens={bool((tok==ref).all())}", flush=True)
