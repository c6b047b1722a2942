import os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from crisperwhisper_b200.engine import Engine
eng = Engine(0)
N = int(os.environ.get("N", "32"))
al = torch.softmax(torch.randn(N, 20, 448, 1500, device="cuda") * 3, -1)
Tl = torch.full((N,), 448, dtype=torch.int32, device="cuda"); Fl = torch.full((N,), 1500, dtype=torch.int32, device="cuda")
for _ in range(2): eng.align(al, Tl, Fl, 7)
eng.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(eng.stream)
for _ in range(5): eng.align(al, Tl, Fl, 7)
e1.record(eng.stream); eng.sync()
ms = e0.elapsed_time(e1) / 5
print(f"align N={N}: {ms:.3f} ms  {N*(20*448*1500*4)/ms/1e6:.1f} GB/s  ({1000*ms/N:.1f} us/utt)")
