"""Summarise an NCDE_PS_TRACE dump: period per stage and the timeline of one whole stage (all traced tiles) of the traced h-group."""
import sys, statistics as st
NAMES = {"fwd": ["flag", "A_in", "mma", "epi_end", "signal", "h_cnt", "z_in", "L0mma", "L0epi", "Llast_mma", "Llast_epi", "stored", "released"],
         "bwd": ["pre_issued", "A_in", "epi1_start", "epi1_end", "epi2_start", "epi2_end", "signal", "h_cnt", "top", "Lmma", "Lepi", "L0mma", "dz",
                 "released", "bias_end", "wgrad_done", "gk_flag", "gk_ready", "blk0", "blk1", "dg_issued", "wg_issued", "e22", "e23"] + ["w%d_e1end" % w for w in range(8)] + ["w%d_start" % w for w in range(8)]}
for kind in ("fwd", "bwd"):
    recs = {}
    for l in open(sys.argv[1]):
        if not l.startswith("pstrace " + kind): continue
        f = l.split()
        q, t = int(f[2][2:]), int(f[3][2:])
        recs[(q, t)] = list(map(int, f[4:]))
    if not recs: continue
    nq = max(q for q, _ in recs) + 1
    qs = [q for q in range(nq // 2 + 4, nq - 1)] if nq > 20 else list(range(2, nq - 1))
    ev = 1
    per = [recs[(q + 1, 0)][ev] - recs[(q, 0)][ev] for q in qs if (q + 1, 0) in recs and recs[(q, 0)][ev] > 0]
    if not per:
        continue
    print(kind, "period %.0f ns (median of %d stages)" % (st.median(per), len(per)))
    q = qs[len(qs) // 2]
    base = min(v for (qq, t), r in recs.items() if qq == q for v in r if v > 0)
    for t in sorted(t for (qq, t) in recs if qq == q):
        r = recs[(q, t)]
        names = NAMES[kind]
        print("  q=%d t=%d:" % (q, t), " ".join("%s=%d" % (n, v - base) for n, v in sorted(zip(names, r), key=lambda x: x[1]) if v > 0))
