"""Per-kernel GPU times of the LAST bench step in an ncu `gpu__time_duration.sum` launch list (csv)."""
import csv, re, sys

def main(path, anchor="subm_insert"):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        name = re.sub(r"^void ", "", r[ki])
        name = re.sub(r"\(.*", "", name).split("<")[0].replace("spx::", "")
        seq.append((name, v))
    starts = [i for i, (n, _) in enumerate(seq) if anchor in n]
    last = seq[starts[-1]:] if starts else seq
    tot = sum(v for _, v in last)
    for n, v in last:
        print(f"{n:45s} {v:8.2f} us  {100 * v / tot:5.1f} %")
    print(f"{'sum':45s} {tot:8.2f} us")

if __name__ == "__main__":
    main(*sys.argv[1:])
