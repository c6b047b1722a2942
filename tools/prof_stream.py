"""One short decode on the streaming step kernel at BASELINE cfg-2 shape (B=8, large-v3) for ncu:
    ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 1 -c 1 -o gpurun_out/prof python tools/prof_stream.py
PROF_T new tokens (default 13 -> one 16-step launch incl. the prompt), random cross K/V."""
import os, sys
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from crisperwhisper_b200 import weights as Wt, _lib as L
from crisperwhisper_b200.engine import Engine
B = int(os.environ.get("PROF_B", "8")); T = int(os.environ.get("PROF_T", "29"))
eng = Engine(0)
cfg = Wt.large_v3_config()
eng.load_weights(Wt.synthetic_weights(cfg, eng.device, seed=0))
xkv = (torch.randn(32, B, 20, 2, 1500, 64, device="cuda") * 0.5).to(torch.bfloat16)
prompt = torch.tensor([[50258, 50259, 50360]] * B, dtype=torch.int32, device="cuda")
out = eng.decode(xkv, prompt, T, flags=L.CW_DEC_SUPPRESS_EOS)
eng.sync()
print("prof_stream done: launches", eng.launch_count(), out["tokens"][0, :8].tolist())
