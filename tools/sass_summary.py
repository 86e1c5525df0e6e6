"""Per-kernel counts of the SASS mnemonics that prove Blackwell-native code paths (B200_PROFILING.md evidence table):
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA tensor copies -> UTMALDG/UTMASTG, 1-D bulk copies -> UBLKCP,
stmatrix -> STSM, cp.async -> LDGSTS, legacy tensor path -> HMMA (must be 0).
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "temporal_inverse_kinematics_b200", "libtik.so")
PATS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "STSM", "LDSM", "LDGSTS", "SYNCS", "HMMA", "FFMA", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("tik::", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["instructions"] += 1
            for p in PATS:
                if op.startswith(p) and not (p == "HMMA" and op.startswith("UTCHMMA")):
                    counts[cur][p] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{p:>7s}" for p in PATS))
    tot = collections.Counter()
    for k, c in counts.items():
        print(f"{k[:78]:78s} {c['instructions']:7d} " + " ".join(f"{c[p]:7d}" for p in PATS))
        tot.update(c)
    print(f"{'TOTAL':78s} {tot['instructions']:7d} " + " ".join(f"{tot[p]:7d}" for p in PATS))


if __name__ == "__main__":
    main()
