"""Per-kernel SASS mnemonic counts of libtxh.so (cuobjdump -sass), written as a table: the evidence that the
kernels are sm_100a code using FP64 tensor-core MMA (DMMA), asynchronous copies (LDGSTS = cp.async, UBLKCP /
UTMALDG = cp.async.bulk / TMA) and the barrier / memory instructions the design relies on.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tx_fast_hydrology_b200", "libtxh.so")
WATCH = ["DMMA", "DFMA", "DADD", "DMUL", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "LDG", "STG", "LDS", "STS", "BAR",
         "ATOMG", "SHFL", "NANOSLEEP", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); kernels[cur] = Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"libtxh.so  arch: {', '.join(arch)}   kernels: {len(kernels)}")
    print("columns: total SASS instructions, then counts of the mnemonics " + " ".join(WATCH))
    for (name, c), pretty in zip(kernels.items(), demangle):
        short = pretty.replace("(anonymous namespace)::", "").replace("txh::", "").replace("void ", "")
        short = re.sub(r"\((?!int|bool).*", "", short).replace("(int)", "").replace("(bool)", "")
        print(f"{short[:60]:<60} {c['_total']:>6} " + " ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
    tot = Counter()
    for c in kernels.values():
        tot.update(c)
    print("ALL " + " ".join(f"{w}={tot[w]}" for w in WATCH))


if __name__ == "__main__":
    sys.exit(main())
