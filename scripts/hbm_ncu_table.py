"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log of scripts/bench_standalone.py:
per kernel launches, average DRAM bytes, duration and DRAM GB/s (development aid).  usage: hbm_ncu_table.py <csv> [peak GB/s]"""
import collections, csv, sys
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6553.9
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
iid = hdr.index("ID")
acc = collections.OrderedDict()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
for r in rows[1:]:
    name = r[ik].split("(")[0].replace("void ", "")
    if not name.startswith("cnf::"):
        continue
    d = acc.setdefault(name, {})
    d.setdefault(r[im], []).append(float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0))
print("%-40s %8s %12s %12s %10s %10s %8s" % ("kernel", "launches", "dram_rd_MB", "dram_wr_MB", "time_us", "dram_GB/s", "of_peak"))
for name, d in acc.items():
    n = len(d["gpu__time_duration.sum"])
    rd = sum(d["dram__bytes_read.sum"]) / n; wr = sum(d["dram__bytes_write.sum"]) / n; t = sum(d["gpu__time_duration.sum"]) / n
    gbs = (rd + wr) / t / 1e3
    print("%-40s %8d %12.2f %12.2f %10.1f %10.1f %8.3f" % (name, n, rd / 1e6, wr / 1e6, t, gbs, gbs / peak))
