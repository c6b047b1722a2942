import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
os.environ["CW_MEGA_DEBUG"] = "1"
from crisperwhisper_b200 import weights as Wt, _lib as L
from crisperwhisper_b200.engine import Engine
B, T = 8, 32
eng = Engine(0)
cfg = Wt.large_v3_config()
eng.load_weights(Wt.synthetic_weights(cfg, eng.device, seed=0))
xkv = (torch.randn(32, B, 20, 2, 1500, 64, device="cuda") * 0.5).to(torch.bfloat16)
prompt = torch.tensor([[50258, 50259, 50360]] * B, dtype=torch.int32, device="cuda")
eng.decode(xkv, prompt, T, flags=L.CW_DEC_SUPPRESS_EOS)
eng.decode(xkv, prompt, T, flags=L.CW_DEC_SUPPRESS_EOS)
eng.sync()
