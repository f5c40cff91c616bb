"""SASS opcode histogram of libruniab200.so per kernel (`cuobjdump -sass`): the mnemonics that prove the
Blackwell-native paths (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAPF = TMA load / store / prefetch, UTCBAR =
tcgen05.commit, LDTM / STTM = tcgen05.ld / st, SYNCS = mbarrier) plus the pipes the bandwidth-bound kernels lean on.
Runs on the build container (no GPU needed).  Writes profiles/r2_sass_histogram.json."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "runia_core_b200", "libruniab200.so")
WATCH = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "FMNMX", "FMNMX3",
         "VIMNMX", "FFMA", "FFMA2", "FADD2", "DFMA", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "REDUX", "CREDUX",
         "SHFL", "BAR", "MEMBAR", "CCTL")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = {}
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("runia::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op, mods = m.group(1), m.group(2)
            cur["_total"] += 1
            cur[op] += 1
            if op == "UTCHMMA" and ".2CTA" in mods:
                cur["UTCHMMA.2CTA"] += 1
            if op == "UTCBAR" and "MULTICAST" in mods:
                cur["UTCBAR.MULTICAST"] += 1
    out = {}
    for k, c in sorted(kernels.items()):
        row = {w: c[w] for w in WATCH + ("UTCHMMA.2CTA", "UTCBAR.MULTICAST") if c[w]}
        row["instructions"] = c["_total"]
        out[k] = row
    path = os.path.join(ROOT, "profiles", "r2_sass_histogram.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    tot = collections.Counter()
    for row in out.values():
        tot.update({k: v for k, v in row.items() if k != "instructions"})
    print(json.dumps({"kernels": len(out), "totals": {k: tot[k] for k in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG",
                                                                       "UTMAPF", "UTCBAR", "LDTM", "SYNCS")}}))
    tc = {k: {w: v[w] for w in ("UTCHMMA", "UTMALDG", "LDTM") if w in v} for k, v in out.items() if "UTCHMMA" in v}
    print(json.dumps(tc, indent=1))


if __name__ == "__main__":
    sys.exit(main())
