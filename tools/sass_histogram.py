#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libradvlm_b200.so (cuobjdump -sass): the tcgen05 / TMEM / TMA / mbarrier
mnemonics that prove which hardware paths a kernel uses (B200_PROFILING.md).  Runs in the container (no GPU).
  python tools/sass_histogram.py [lib] > profiles/<round>_sass_opcodes.csv"""
import collections
import re
import subprocess
import sys

WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UTCATOMSWS", "SYNCS",
         "UCGABAR", "MUFU.EX2", "MUFU.TANH", "RED", "ATOM", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "HMMA", "FFMA"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "radvlm_b200/libradvlm_b200.so"
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                    cur[w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("kernel,total," + ",".join(WATCH))
    for (name, c), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm).replace("void ", "").replace("rv::", "")
        print('"%s",%d,%s' % (short, c["total"], ",".join(str(c[w]) for w in WATCH)))


if __name__ == "__main__":
    main()
