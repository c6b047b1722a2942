"""Per-kernel SASS opcode histogram of libcrisper.so (evidence of which kernels use TMA / tcgen05 / legacy HMMA):
    python tools/sass_hist.py > profiles/r02_sass_opcodes.md"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "crisperwhisper_b200/libcrisper.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["UBLKCP", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HMMA", "LDGSTS", "SYNCS", "LDSM", "MUFU", "ATOMG", "RED", "CCTL", "MEMBAR", "BAR"]
print("# SASS opcode evidence, libcrisper.so (sm_100a), round 2\n")
print("`cuobjdump -sass crisperwhisper_b200/libcrisper.so`, opcode counts per kernel (static instruction counts). UBLKCP = `cp.async.bulk` "
      "(1-D TMA), UTMALDG = `cp.async.bulk.tensor` (tiled TMA), UTCHMMA = `tcgen05.mma`, LDTM = `tcgen05.ld`, SYNCS = mbarrier ops, "
      "HMMA = legacy `mma.sync`.\n")
print("| kernel | total | " + " | ".join(KEY) + " |")
print("|---|---:|" + "---:|" * len(KEY))
for k, c in hist.items():
    if sum(c.values()) < 40:
        continue
    print(f"| `{k}` | {sum(c.values())} | " + " | ".join(str(c.get(x, 0)) for x in KEY) + " |")
print("\nTop opcodes of the decode step kernel:\n")
for k, c in hist.items():
    if "decode_stream_kernel" in k:
        print("```\n" + "\n".join(f"{n:6d} {op}" for op, n in c.most_common(30)) + "\n```")
