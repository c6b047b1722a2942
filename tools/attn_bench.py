"""Encoder self-attention microbenchmark (cw_attention_enc): B x 1500 x 20 heads x 64, bf16.  Prints ms and TFLOP/s and
the max abs error against torch fp32 attention on one (sample, head)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crisperwhisper_b200.engine import Engine

B, S, H = int(os.environ.get("B", "8")), 1500, 20
eng = Engine(0)
g = torch.Generator(device="cuda").manual_seed(1)
qkv = (torch.randn(B * S, 3 * H * 64, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
out = eng.attention_enc(qkv, B, S, H)
eng.sync()
q, k, v = [qkv[:S, i * H * 64: i * H * 64 + 64].float() for i in range(3)]
ref = torch.softmax(q @ k.T, -1) @ v      # q is pre-scaled in the real model; here the raw product is the test
err = (out[:S, :64].float() - ref).abs().max().item()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = []
for _ in range(10):
    e0.record(eng.stream); out = eng.attention_enc(qkv, B, S, H); e1.record(eng.stream); eng.sync()
    ms.append(e0.elapsed_time(e1))
t = float(np.median(ms))
print(f"attention_enc B={B}: {t:.3f} ms  {4.0 * S * S * 64 * H * B / t / 1e9:.1f} TFLOP/s  max|err| vs fp32 {err:.4f}")
