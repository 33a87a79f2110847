"""Per-CTA view of one (stage, tile) of the backward pass from an NCDE_PS_TRACE dump: who is late, and in which phase."""
import sys
import statistics as st
rows = []
for l in open(sys.argv[1]):
    if l.startswith("pstrace_all bwd"):
        f = l.split()
        rows.append((int(f[2][2:]), list(map(int, f[3:]))))
gs = sorted(set(g for g, _ in rows))
rows = rows[-len(gs):]      # last call
base = min(r[1] for _, r in rows)
rows.sort(key=lambda r: r[1][6])
for g, r in rows[:4] + rows[len(rows) // 2 - 2:len(rows) // 2 + 2] + rows[-8:]:
    print("g=%3d sm=%3d  A_in=%6d  pre=%5d  epi1=%5d  dgwait=%5d  epi2=%5d  signal=%6d" % (g, r[7], r[1] - base, r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - base))
for k, (i, j) in {"pre": (1, 2), "epi1": (2, 3), "dgwait": (3, 4), "epi2": (4, 5), "sig": (5, 6)}.items():
    v = [r[j] - r[i] for _, r in rows]
    print(k, "min %d med %d max %d" % (min(v), st.median(v), max(v)))
v = [r[6] - base for _, r in rows]
print("signal spread: min %d med %d max %d" % (min(v), st.median(v), max(v)))
late = [(g, r[7]) for g, r in rows if r[6] - base > st.median(v) + 10000]
print("late CTAs (g, sm):", sorted(late))
print("all SMs used by field CTAs:", sorted(r[7] for _, r in rows))
