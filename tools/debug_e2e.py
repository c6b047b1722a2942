"""Debug helper (GPU box): token-level comparison of HF generate vs crisperwhisper_b200.generate on the golden cases."""
import json, os, sys, warnings
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from oracle import hf_harness as H
from crisperwhisper_b200 import weights as Wt, generate as G, audio as A
from crisperwhisper_b200.engine import Engine
from crisperwhisper_b200.asr_pipeline import mel_filters_slaney
from transformers import WhisperFeatureExtractor

eng = Engine(0)
g = json.load(open(os.path.join(ROOT, "tests/golden/pipeline_hf.json")))
long70 = np.concatenate([H.speechlike(3), H.noise(4), H.noise(5, 160000)])
waves = {"clip5s": H.noise(0, 80000), "clip70s": long70, "clip70s_bs1": long70, "clip12s_80": H.speechlike(6, 12 * 16000)}
report = {}
for name in (sys.argv[1:] or list(waves)):
    c = g[name]
    m = H.build_model(H.tiny_hf_config(n_mels=c["n_mels"]), seed=c["seed"], logit_scale=c["logit_scale"], pos_scale=c["pos_scale"])
    pw = Wt.pack_hf_model(m, device=eng.device)
    pw.config["lang_id"], pw.config["task_id"] = H.TOK_IDS["en"], H.TOK_IDS["transcribe"]
    eng.load_weights(pw)
    fe = WhisperFeatureExtractor(feature_size=c["n_mels"])
    wave = waves[name]
    plan = A.chunk_plan(len(wave), 30.0)
    bs = c["batch_size"]
    filt = torch.from_numpy(mel_filters_slaney(c["n_mels"])).to(eng.device)
    rep = []
    for b0 in range(0, len(plan), bs):
        items = plan[b0:b0 + bs]
        chunks = [wave[s:s + ln] for (s, ln, _, _, _) in items]
        feats = fe(chunks, sampling_rate=16000, return_tensors="pt", return_attention_mask=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gc = m.generation_config
            out = m.generate(feats["input_features"], attention_mask=feats["attention_mask"], return_timestamps=True,
                             return_token_timestamps=True, return_segments=True, max_new_tokens=c["max_new_tokens"], num_beams=1)
        host = np.zeros((len(items), 480000), np.float32)
        for k, ch in enumerate(chunks):
            host[k, :len(ch)] = ch
        nv = torch.tensor([len(ch) for ch in chunks], dtype=torch.int32, device=eng.device)
        f32, tm, frames = eng.logmel(torch.from_numpy(host).to(eng.device), filt, nv, want_f32=True)
        eng.sync()
        st = {}
        mine = G.generate(eng, tm, frames.cpu().numpy(), G.GenOptions(max_new_tokens=c["max_new_tokens"]), st)
        for k in range(len(items)):
            hf_tok = torch.cat([s["tokens"] for s in out["segments"][k]]).tolist() if out["segments"][k] else []
            hf_ts = torch.cat([s["token_timestamps"] for s in out["segments"][k]]).tolist() if out["segments"][k] else []
            rep.append({"chunk": b0 + k, "frames": int(frames[k]), "hf_frames": int(feats["attention_mask"][k].sum()),
                        "feat_maxdiff": float((f32[k].cpu() - feats["input_features"][k]).abs().max()),
                        "hf_tokens": hf_tok, "my_tokens": mine[k]["tokens"].tolist(),
                        "hf_ts": [round(x, 2) for x in hf_ts], "my_ts": [round(float(x), 2) for x in mine[k]["token_timestamps"]],
                        "hf_nseg": len(out["segments"][k]), "my_nseg": len(mine[k]["segments"]), "stats": dict(st)})
    report[name] = rep
    for r in rep:
        same = r["hf_tokens"] == r["my_tokens"]
        print(name, "chunk", r["chunk"], "tokens equal:", same, "ts equal:", r["hf_ts"] == r["my_ts"], "nseg", r["hf_nseg"], r["my_nseg"],
              "len", len(r["hf_tokens"]), len(r["my_tokens"]), "featdiff %.2e" % r["feat_maxdiff"])
        if not same:
            n = min(len(r["hf_tokens"]), len(r["my_tokens"]))
            d = [i for i in range(n) if r["hf_tokens"][i] != r["my_tokens"][i]]
            print("   first diff at", d[:1], "hf", r["hf_tokens"][:40], "\n   my", r["my_tokens"][:40])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(ROOT, "gpurun_out", "debug_e2e.json"), "w"))
