"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel -> JSON + table."""
import collections, csv, json, re, sys
src, dst = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
lines = [l for l in open(src) if not l.startswith("==")]
agg, total = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns; total += ns
out = [{"kernel": k, "launches": n, "total_ms": round(ns / 1e6, 3), "avg_us": round(ns / n / 1e3, 2),
        "share": round(ns / total, 4)} for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
print("total ms %.3f" % (total / 1e6))
for o in out[:12]:
    print("%-60s n=%5d total=%9.3f ms avg=%8.2f us share=%.3f" % (o["kernel"][:60], o["launches"], o["total_ms"], o["avg_us"], o["share"]))
if dst:
    json.dump({"total_ms": total / 1e6, "kernels": out}, open(dst, "w"), indent=1)
