"""profiles/r02_k_accumulate_traffic.json from an ncu metrics pass:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:k_accumulate --csv --log-file gpurun_out/X.csv  python tools/prove_once.py      (SERIAL=1 LOG=24 REPS=1)
    python tools/ncu_traffic.py gpurun_out/X.csv 24
bench.py reads roofline.traffic (DRAM read + write bytes of one G1 accumulate launch, mean over the launches of the
prove) from that file instead of a constant."""
import collections
import csv
import json
import os
import sys

path, log_rows = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ID, K, M, V = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[ID], r[K]), {})[r[M]] = float(r[V].replace(",", ""))
out = {}
for name, tag in (("FqCfg", "g1"), ("Fq2", "g2")):
    ls = [m for (i, k), m in per.items() if "k_accumulate" in k and name in k]
    if not ls:
        continue
    out[tag] = {"launches": len(ls),
                "dram_bytes_per_launch": sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in ls) / len(ls),
                "read": sum(m["dram__bytes_read.sum"] for m in ls) / len(ls),
                "write": sum(m["dram__bytes_write.sum"] for m in ls) / len(ls),
                "ms_per_launch_under_ncu": sum(m["gpu__time_duration.sum"] for m in ls) / len(ls) / 1e6}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(root, "profiles", "r02_k_accumulate_traffic.json")
cur = json.load(open(dst)) if os.path.exists(dst) else {}
cur[str(log_rows)] = {"dram_bytes_per_launch": out["g1"]["dram_bytes_per_launch"], "g1": out.get("g1"), "g2": out.get("g2"),
                      "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the k_accumulate<Fq> launches of one "
                                f"serial-schedule prove of 2^{log_rows} rows ({os.path.basename(path)})"}
json.dump(cur, open(dst, "w"), indent=1)
print(json.dumps(cur[str(log_rows)], indent=1))
