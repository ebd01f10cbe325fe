"""Per-kernel count of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM
(tcgen05.ld), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce), .IM2COL loads, MULTIMEM (NVLS).
    python tools/sass_table.py > profiles/rNN_sass_table.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal_classification_b200", "libvilbert_b200.so")
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "IM2COL", "UTMASTG", "UTMAREDG", "MULTIMEM", "MUFU.TANH"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        for op in OPS:
            if op in line:
                counts[cur][op] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonics per kernel of libvilbert_b200.so (`cuobjdump -sass`, sm_100a)\n")
    print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, IM2COL = TMA loads in im2col")
    print("mode (implicit-GEMM convolution), MULTIMEM = NVLS multimem.ld_reduce / multimem.st.\n")
    print("| kernel | " + " | ".join(OPS) + " |\n|---|" + "---|" * len(OPS))
    for n, d in zip(names, dem):
        c = counts[n]
        if not any(c[o] for o in OPS):
            continue
        short = re.sub(r"\(.*", "", d).replace("void ", "")
        print(f"| `{short}` | " + " | ".join(str(c[o]) for o in OPS) + " |")


if __name__ == "__main__":
    main()
