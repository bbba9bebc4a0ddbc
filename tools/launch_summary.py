"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.

With --last-prove only the launches of the last prove in the file are counted (a prove starts at its
first k_spmv launch: R1CS evaluation is the first thing a prove does on the main stream in the serial
schedule; the witness-only MSMs that precede it on the same stream in that schedule are attributed by
walking back to the previous prove's end)."""
import collections
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
rows = rows[1:]
if "--last-prove" in sys.argv:
    # a prove = everything between two h2d-separated groups; we split on the witness MSM's first k_digits<0>
    # that follows a k_fold_parts/k_segment_bits of the H MSM, i.e. on groups of 3 consecutive k_spmv
    # only the library's prove kernels (no torch fill kernels of the L2 flush, no pipe probe after the last prove)
    rows = [r for r in rows if "fb::" in r[ki] and "k_probe" not in r[ki]] or rows
    spmv = [i for i, r in enumerate(rows) if "k_spmv" in r[ki].split("(")[0]]
    groups = [spmv[i] for i in range(0, len(spmv), 3)]
    if len(groups) >= 2:
        per = groups[-1] - groups[-2]
        end = len(rows)
        # the prove's launches: same count as the distance between two proves, ending at the file end
        rows = rows[end - per:]
agg = collections.OrderedDict()
tot = 0.0
for r in rows:
    k = r[ki].replace("fb::", "")[:70]
    v = float(r[vi].replace(",", ""))
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
    tot += v
print(f"total {tot/1e6:.3f} ms over {len(rows)} launches")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{c:5d} {t/1e6:12.3f} ms {100*t/tot:6.1f}%  {k}")
