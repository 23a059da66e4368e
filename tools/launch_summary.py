#!/usr/bin/env python
"""Condense an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) into kernel-time shares.

    python tools/launch_summary.py X.csv STEPS "header comment" > X_summary.txt
"""
import csv
import re
import sys


def main(path, steps, comment):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = {}
    n = 0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("ocp::<unnamed>::", "").replace("void ", "").replace("<unnamed>::", "")
        t = float(r[mv].replace(",", ""))
        t = t / 1000.0 if r[mu] in ("ns", "nsecond") else t
        a = agg.setdefault(name, [0.0, 0])
        a[0] += t
        a[1] += 1
        n += 1
    tot = sum(a[0] for a in agg.values())
    print(f"# {comment}")
    print(f"# total {tot:.1f} us over {n} launches = {tot / steps:.1f} us per GD iteration (per-launch times are serialised: compare SHARES)")
    print("share%  total_us  launches  kernel")
    for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * t / tot:6.2f} {t:10.1f} {c:7d}  {name}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "")
