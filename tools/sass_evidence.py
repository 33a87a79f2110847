"""Counts the SASS mnemonics that prove the tensor-core / TMA path per kernel of the built library (no GPU needed):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, HMMA = legacy mma.sync.

    python tools/sass_evidence.py [profiles/r01_end_sass_evidence.json]
"""
import collections, json, os, re, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(R, "online-neural-cdes_b200", "csrc", "build", "solve.o")
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
pat = {"tcgen05.mma (UTC*MMA)": r"\bUTC[A-Z]*MMA", "tcgen05.ld (LDTM)": r"\bLDTM", "tcgen05.st (STTM)": r"\bSTTM",
       "TMA load (UTMALDG)": r"\bUTMALDG", "TMA store (UTMASTG)": r"\bUTMASTG", "bulk copy (UBLKCP)": r"\bUBLKCP",
       "tanh (MUFU.TANH)": r"MUFU\.TANH", "legacy mma.sync (HMMA)": r"\bHMMA", "fp32 FFMA": r"\bFFMA"}
res, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("void ", "").replace("ncde::", "")
        res[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pat.items():
            if re.search(p, line):
                res[cur][k] += 1
keep = {k: dict(v) for k, v in res.items() if k.startswith("tc_") or k.startswith("field_") or k.startswith("hidden_")}
for k, v in keep.items():
    print("%-34s %s" % (k, ", ".join("%s: %d" % kv for kv in v.items())))
if len(sys.argv) > 1:
    json.dump(keep, open(sys.argv[1], "w"), indent=1)
