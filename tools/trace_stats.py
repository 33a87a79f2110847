import sys, statistics as st
for kind, names in (("fwd", ["flag", "A_in", "mma", "epi_end", "signal", "h_cnt", "z_in", "L0mma", "L0epi", "L1mma", "L1epi", "stored", "released"]),
                    ("bwd", ["deps", "A_in", "epi1_start", "epi1_end", "epi2_start", "epi2_end", "signal", "h_cnt", "top", "Lmma", "Lepi", "L0mma", "dz", "released"])):
    rows = [list(map(int, l.split()[3:])) for l in open(sys.argv[1]) if l.startswith("pstrace " + kind)]
    rows = rows[len(rows) // 2 + 8:] if len(rows) > 60 else rows[8:]    # second call's stages (warm)
    if not rows: continue
    print(kind, "period %.0f ns" % st.median([rows[i + 1][1] - rows[i][1] for i in range(len(rows) - 1)]))
    r = rows[len(rows) // 2]
    base = min(v for v in r if v > 0)
    print("  one stage:", " ".join("%s=%d" % (n, v - base) for n, v in sorted(zip(names, r), key=lambda x: x[1])))
