#!/usr/bin/env python
"""DRAM traffic per launch of the GEMM kernels from an `ncu --set full` capture of
`bench.py --graph 0 --extras 0` (eager order per step: forward, input gradient, weight gradient, reduce)
-> profiles/r02_traffic.json, keyed by workload, as bench.py's roofline.traffic expects.
usage: tools/ncu_traffic.py gpurun_out/r2b_full.ncu-rep submconv3d_k3_c64_fp16_100k_kitti"""
import csv, io, json, os, subprocess, sys

rep, workload = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
unit_r, unit_w = rows[1][ir], rows[1][iw]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
seq = []
for r in rows[2:]:
    name = r[ik]
    b = float(r[ir].replace(",", "")) * scale.get(unit_r, 1) + float(r[iw].replace(",", "")) * scale.get(unit_w, 1)
    seq.append((name, b))
gg = [b for n, b in seq if "tc_gather_gemm" in n]
wg = [b for n, b in seq if "tc_wgrad" in n]
rd = [b for n, b in seq if "wgrad_reduce" in n]
res = {"fwd": int(sum(gg[0::2]) / max(len(gg[0::2]), 1)), "dgrad": int(sum(gg[1::2]) / max(len(gg[1::2]), 1)),
       "wgrad": int(sum(wg) / max(len(wg), 1)),
       # ncu flushes the caches before every kernel: the reduce kernel then re-reads the fp32 partials from DRAM,
       # which it finds in L2 in a real run -- listed for completeness, not part of the roofline traffic
       "wgrad_reduce_under_ncu_cache_flush": int(sum(rd) / max(len(rd), 1)),
       "source": os.path.basename(rep), "launches": {"gather_gemm": len(gg), "wgrad": len(wg), "reduce": len(rd)}}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[workload] = res
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(res))
