"""Source-level hot spots of one kernel from an .ncu-rep (needs -lineinfo + --import-source on)."""
import collections, csv, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 18
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + kernel],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg, text = collections.Counter(), {}
fname, si = "", None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        si = None
        continue
    if len(r) > 5 and any(c.startswith("# Samples") for c in r):
        si = [i for i, c in enumerate(r) if c.startswith("# Samples")][0]
        continue
    if si is None or len(r) <= si or not r[0].strip():
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    key = (fname, r[0])
    agg[key] += n
    text.setdefault(key, r[1][:105])
tot = sum(agg.values()) or 1
for k, n in agg.most_common(top):
    print("%5.1f%% | %-18s %4s | %s" % (100 * n / tot, k[0][:18], k[1], text[k]))
