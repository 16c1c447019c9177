"""SASS opcode summary of libwvd.so: per kernel, the counts of the mnemonics that prove which hardware path it uses
(tcgen05 MMA = UTCHMMA / UTCQMMA..., TMEM ld/st = LDTM / STTM, TMA = UTMALDG / UTMASTG, mbarrier = SYNCS, cluster
barriers = UCGABAR, MUFU.EX2 ...).      python tools/sass_summary.py > profiles/r2_sass_summary.txt
Runs on the build container (cuobjdump, no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video_styler_b200", "libwvd.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UCGABAR", "MUFU.EX2", "MUFU.TANH", "MUFU.RCP",
         "HMMA", "FFMA2", "FFMA", "HFMA2", "HADD2", "HMUL2", "F2FP", "FMNMX3", "LDS", "STS", "LDG", "STG", "BAR.SYNC", "BAR.ARV", "ATOM", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if cur and m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
                    break
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode summary of video_styler_b200/libwvd.so ({os.path.getsize(LIB)} bytes), cuobjdump -sass, sm_100a.")
    print("# per kernel: total instructions, then the counts of the watched mnemonics (static counts in the binary, not executed counts)")
    tot = collections.Counter()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("void ", "")
        items = " ".join(f"{k}={v}" for k, v in c.items() if k != "_total")
        print(f"{short:100s} total={c['_total']:6d}  {items}")
        tot.update(c)
    print("# whole library: " + " ".join(f"{k}={v}" for k, v in tot.items()))


if __name__ == "__main__":
    sys.exit(main())
