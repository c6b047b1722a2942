"""Decode-step microbenchmark: ms per step for graph/eager x PDL/no-PDL (B=8, large-v3 shape, random xkv)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from crisperwhisper_b200 import weights as Wt, _lib as L
from crisperwhisper_b200.engine import Engine
B = int(os.environ.get("B", "8")); T = int(os.environ.get("T", "64"))
eng = Engine(0)
cfg = Wt.large_v3_config()
eng.load_weights(Wt.synthetic_weights(cfg, eng.device, seed=0))
xkv = (torch.randn(32, B, 20, 2, 1500, 64, device="cuda") * 0.5).to(torch.bfloat16)
prompt = torch.tensor([[50258, 50259, 50360]] * B, dtype=torch.int32, device="cuda")
CASES = (("mega (default)", 0), ("mega eager", L.CW_DEC_NO_GRAPH), ("ops graph+pdl", L.CW_DEC_NO_MEGA),
         ("ops eager+pdl", L.CW_DEC_NO_MEGA | L.CW_DEC_NO_GRAPH))
if os.environ.get("ONLY_MEGA"):
    CASES = CASES[:1]
for name, fl in CASES:
    flags = L.CW_DEC_SUPPRESS_EOS | fl
    for _ in range(2):
        eng.decode(xkv, prompt, T, flags=flags, want_align=True)
    eng.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = eng.decode(xkv, prompt, T, flags=flags, want_align=True)
    eng.sync(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:10s} {1000*dt/(T+2):.3f} ms/step  tokens[0,:6]={out['tokens'][0,:6].tolist()}", flush=True)
