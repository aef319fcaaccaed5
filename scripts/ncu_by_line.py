"""Join an ncu SASS source page (CSV) with nvdisasm line info: instructions executed / stall samples per CUDA source line.
usage: ncu_by_line.py <ncu source csv> <nvdisasm --print-line-info of the kernel> [top_n]"""
import csv, re, sys, collections
src_csv, dis = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
line_of = {}
cur = None
for ln in open(dis):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);', ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
inst = collections.Counter(); samp = collections.Counter()
for r in rows[2:]:
    if len(r) <= isamp: continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None: base = a
    key = line_of.get(a - base, ("?", 0))
    inst[key] += int(r[ii] or 0); samp[key] += int(r[isamp] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp instructions %d, samples %d" % (ti, ts))
print("%-22s %6s %8s %8s" % ("file", "line", "inst%", "samp%"))
for key, v in sorted(inst.items(), key=lambda kv: -samp[kv[0]])[:topn]:
    print("%-22s %6d %8.2f %8.2f" % (key[0], key[1], 100.0 * v / ti, 100.0 * samp[key] / ts))
# per-file-line dump for range aggregation
import json
json.dump({"%s:%d" % k: [inst[k], samp[k]] for k in inst}, open("/tmp/by_line.json", "w"))
