"""Stage-3 microbenchmark: cw_align (align_reduce_kernel + dtw_kernel) on N utterances of H x T x F float32 rows.

    python tools/align_bench.py [--n 128] [--heads 20] [--t 448] [--f 1500] [--iters 10] [--ragged]

Prints the CUDA-event time of the whole call and the algorithmic HBM rate (H*T*F*4 bytes read per utterance).  Run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split, or `ncu --set full -k regex:align_reduce` for the capture."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crisperwhisper_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--heads", type=int, default=20)
    ap.add_argument("--t", type=int, default=448)
    ap.add_argument("--f", type=int, default=1500)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--width", type=int, default=7)
    ap.add_argument("--ragged", action="store_true", help="random T_len / F_len per utterance")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    eng = Engine(dev)
    N, H, T, F = args.n, args.heads, args.t, args.f
    al = torch.empty(N, H, T, F, dtype=torch.float32, device=dev)
    tt = torch.arange(T, device=dev, dtype=torch.float32)[None, None, :, None]
    ff = torch.arange(F, device=dev, dtype=torch.float32)[None, None, None, :]
    peak = 6.0 * torch.exp(-(((ff - tt * F / T) / 20.0) ** 2))
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    for j in range(0, N, 8):
        k = min(8, N - j)
        al[j:j + k] = torch.softmax(3.0 * torch.randn(k, H, T, F, generator=gen, device=dev) + peak, -1)
    if args.ragged:
        rs = np.random.RandomState(3)
        Tl = torch.from_numpy(rs.randint(1, T + 1, size=N).astype(np.int32)).to(dev)
        Fl = torch.from_numpy(rs.randint(1, F + 1, size=N).astype(np.int32)).to(dev)
    else:
        Tl = torch.full((N,), T, dtype=torch.int32, device=dev)
        Fl = torch.full((N,), F, dtype=torch.int32, device=dev)
    out = eng.align(al, Tl, Fl, args.width)
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(args.iters):
        e0.record(eng.stream)
        out = eng.align(al, Tl, Fl, args.width)
        e1.record(eng.stream)
        eng.sync()
        ms.append(e0.elapsed_time(e1))
    med = float(np.median(ms))
    nbytes = float((Tl.double() * Fl.double()).sum().item()) * H * 4
    print(f"align N={N} H={H} T={T} F={F} ragged={args.ragged}: {med:.3f} ms (min {min(ms):.3f})  {nbytes / med / 1e6:.1f} GB/s algorithmic"
          f"  checksum {int(out.to(torch.int64).sum().item())}")


if __name__ == "__main__":
    main()
