"""Launch shares of our kernels from an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list.
usage: python tools/launch_shares.py launches.csv [command description]"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0].replace("void ", "")
    if not any(k in name for k in ("gemm_bf16", "area_", "attention", "head_kernel", "head_mma", "vpass", "layernorm", "topk", "sim_", "text_embed",
                                   "hpass", "zero_pad", "rescore", "nv12")):
        continue            # torch's own glue kernels (fill, copy) are not ours
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print("launch shares of OUR kernels, ncu launch list of `%s` (cold-cache serialised times: compare shares)" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{100 * v / s:6.2f} %  {v:9.3f} ms  {cnt[name]:5d} launches  {name}")
