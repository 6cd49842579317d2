#!/usr/bin/env python
"""Per-kernel SASS census of libunetb200.so: how many tcgen05 MMA (UTCHMMA / UTCQMMA ...), TMA load / store (UTMALDG / UTMASTG),
TMEM load (LDTM), TMEM alloc (UTCATOMSWS) and mbarrier (SYNCS) instructions each kernel contains -- the evidence that the conv /
deconv / wgrad kernels are tcgen05 + TMEM + TMA code and that the memory-bound kernels are not.
  python tools/sass_census.py [path/to/lib.so] > profiles/r02_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "semantic-segmentation-unet_b200", "libunetb200.so")
PAT = ["UTCHMMA", "2CTA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCATOMSWS", "UTCBAR", "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS", "REDG", "ATOM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    counts = []
    cur = None
    it = iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            counts.append((next(it, m.group(1)), cur))
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for p in PAT:
                if op == p or op.startswith(p + "."):
                    cur[p] += 1
            if ".2CTA" in op:
                cur["2CTA"] += 1
    print("# SASS census of libunetb200.so (sm_100a), `cuobjdump -sass` opcode counts per kernel\n")
    print("`UTCHMMA` = tcgen05.mma (kind::f16), `UTMALDG`/`UTMASTG` = TMA tensor load/store, `LDTM` = tcgen05.ld (TMEM -> registers), "
          "`UTCATOMSWS` = tcgen05.alloc/dealloc, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `2CTA` = instructions of the cta_group::2 forms (`UTCHMMA.2CTA`, `UTMALDG.*.2CTA`, `UTCBAR.2CTA.MULTICAST`: the CTA-pair conv kernel).  `HMMA` (mma.sync) must be 0 everywhere.\n")
    print("| kernel | instr | " + " | ".join(PAT) + " |")
    print("|---|---:|" + "---:|" * len(PAT))
    tot = collections.Counter()
    for name, c in sorted(counts, key=lambda kv: (-kv[1]["UTCHMMA"], kv[0])):
        short = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", short)          # drop the argument list, keep the template arguments
        short = re.sub(r"\((?:int|bool)\)", "", short).replace("void ", "")
        print(f"| `{short[:110]}` | {c['_total']} | " + " | ".join(str(c[p]) if c[p] else "" for p in PAT) + " |")
        tot.update(c)
    print(f"| **total ({len(counts)} kernels)** | {tot['_total']} | " + " | ".join(str(tot[p]) for p in PAT) + " |")


if __name__ == "__main__":
    main()
