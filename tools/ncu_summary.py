#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel headline metrics as CSV + the
top stall-sample SASS lines.   usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx"""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def source(rep, kernel_regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel_regex}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not his:
        return []
    hi = his[0]
    hdr = rows[hi]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for idx, r in enumerate(rows[hi + 1:]):
        try:
            data.append((int(r[isamp] or 0), int(r[iex] or 0), r[isrc].strip(), idx))
        except Exception:
            break
    return data


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    hdr, units, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    keep = [k for k in KEEP if k in idx]
    with open(prefix + "_kernels.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[idx[k]] for k in keep])
        for r in rows:
            w.writerow([r[idx[k]] for k in keep])
    for r in rows:
        print(" | ".join(f"{k.split('.')[0][-28:]}={r[idx[k]]}" for k in keep[:2] + keep[6:8] + keep[11:15]))
    def base(n):
        n = n.split("(")[0].split("<")[0].strip()
        if n.startswith("void "):
            n = n[5:]
        return n.split("::")[-1].strip()
    names = sorted(set(base(r[idx["Kernel Name"]]) for r in rows))
    with open(prefix + "_stalls.txt", "w") as f:
        for name in names:
            data = source(rep, name)
            tot = sum(d[0] for d in data) or 1
            f.write(f"==== {name}: {tot} samples, {len(data)} SASS instructions (first launch in the report)\n")
            for s, e, src, i in sorted(data, key=lambda d: -d[0])[:25]:
                f.write(f"{s:6d} {100.0 * s / tot:5.1f}%  exec={e:9d}  #{i:5d}  {src[:110]}\n")
    print("wrote", prefix + "_kernels.csv", prefix + "_stalls.txt")


if __name__ == "__main__":
    main()
