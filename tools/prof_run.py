"""One pass of the hot path at BASELINE cfg-2 shapes for ncu (launch list / --set full captures):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_run.py
Decode runs with direct launches (CW_DEC_NO_GRAPH) so every kernel is visible; T defaults to 4 new tokens."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from crisperwhisper_b200 import weights as Wt, _lib as L
from crisperwhisper_b200.engine import Engine
from crisperwhisper_b200.asr_pipeline import mel_filters_slaney

B = int(os.environ.get("PROF_B", "8")); T = int(os.environ.get("PROF_T", "4")); N5 = int(os.environ.get("PROF_N5", "8"))
what = set((os.environ.get("PROF_WHAT", "logmel,encode,decode,align")).split(","))
eng = Engine(0)
cfg = Wt.large_v3_config()
eng.load_weights(Wt.synthetic_weights(cfg, eng.device, seed=0))
wave = torch.from_numpy(np.stack([(np.random.default_rng(i).standard_normal(480000) * 0.1).astype(np.float32) for i in range(B)])).cuda()
filt = torch.from_numpy(mel_filters_slaney(128)).cuda()
_, tm, _ = eng.logmel(wave, filt, None, want_f32=True)
if "encode" in what or "decode" in what:
    xkv, _ = eng.encode(tm)
if "decode" in what:
    prompt = torch.tensor([[50258, 50259, 50360]] * B, dtype=torch.int32, device="cuda")
    mega = 0 if os.environ.get("PROF_MEGA", "0") == "1" else L.CW_DEC_NO_MEGA
    out = eng.decode(xkv, prompt, T, flags=L.CW_DEC_SUPPRESS_EOS | L.CW_DEC_NO_GRAPH | mega)
if "align" in what:
    al = torch.softmax(torch.randn(N5, 20, 448, 1500, device="cuda") * 3, -1)
    eng.align(al, torch.full((N5,), 448, dtype=torch.int32), torch.full((N5,), 1500, dtype=torch.int32), 7)
eng.sync()
print("prof_run done: launches", eng.launch_count())
