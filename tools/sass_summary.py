"""SASS mnemonic counts of the built library (no GPU needed): python tools/sass_summary.py > profiles/r02_sass_summary.txt
The mnemonics that prove the Blackwell paths: UTCHMMA / UTCIMMA = tcgen05.mma (f16, tf32 / i8), LDTM = tcgen05.ld,
UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier, UTCBAR = tcgen05.commit."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spconv_b200", "lib", "libspconv_b200.so")
KEYS = ["UTCHMMA", "UTCIMMA", "LDTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "UTCBAR", "UTCATOMSWS", "ELECT", "ATOMG", "ATOMS",
        "REDG", "MEMBAR", "LDG", "STG", "MATCH", "VOTE"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    per, cur, order, idx = {}, None, [], 0
    for ln in sass.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = re.sub(r"\(.*", "", names[idx]); idx += 1
            per.setdefault(cur, collections.Counter()); order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    per[cur][k] += 1
                    break
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("# SASS instruction counts of spconv_b200/lib/libspconv_b200.so (cuobjdump -sass; sm_100a), tools/sass_summary.py")
    print("# UTCHMMA/UTCIMMA = tcgen05.mma (f16,tf32 / i8), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,")
    print("# LDGSTS = cp.async, SYNCS = mbarrier, UTCBAR = tcgen05.commit, MEMBAR / REDG: fences and reductions (peer exchange, sorts)\n")
    print("total: " + ", ".join(f"{k} {total[k]}" for k in KEYS if total[k]) + "\n")
    seen = set()
    for n in order:
        if n in seen:
            continue
        seen.add(n)
        c = per[n]
        if c:
            print(f"{n[:100]:100s} " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))


if __name__ == "__main__":
    sys.exit(main())
