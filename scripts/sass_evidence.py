"""Blackwell-native evidence from the built library's SASS (B200_PROFILING.md "What proves a Blackwell-native
kernel"): counts of UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG / UTMAPF (TMA load / store / prefetch),
UTCBAR (tcgen05.commit) and legacy HMMA (mma.sync — must be 0) per kernel.

    python scripts/sass_evidence.py [lib.so]        # prints the table (profiles/r01/sass_evidence.txt)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "HMMA"]


def sass_counts(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        c = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", f, re.M):
            c[m.group(1)] += 1
        out[name] = c
    return out


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
        ROOT, "selectivenet_for_semantic_segmentation_binary_b200", "libsunet_b200.so")
    counts = sass_counts(lib)
    print("kernel".ljust(64), *[k.rjust(8) for k in KEYS])
    for name, c in sorted(counts.items()):
        if any(c[k] for k in KEYS):
            short = re.sub(r"^_ZN5sunet\d+", "", name)
            short = re.sub(r"E?v?14CUtensorMap_st.*", "", short)
            print(short[:62].ljust(64), *[str(c[k]).rjust(8) for k in KEYS])


if __name__ == "__main__":
    main()
