"""Print one training step's kernels from an ncu `--metrics gpu__time_duration.sum --csv` log: python scripts/launch_list.py <csv>"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
seq = []
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    if r[ui] in ("ns", "nsecond"): v /= 1000.0
    seq.append((r[ki].split("(")[0][-48:], v))
ends = [i for i, (n, _) in enumerate(seq) if "finalize_kernel" in n]      # last kernel of a step
a, b = (ends[0] + 1, ends[1] + 1) if len(ends) >= 2 else (0, len(seq))
tot = sum(v for _, v in seq[a:b])
for n, v in seq[a:b]:
    print("%-50s %8.2f us  %5.1f%%" % (n, v, 100 * v / tot))
print("%-50s %8.2f us" % ("one step (serialised, cold caches)", tot))
