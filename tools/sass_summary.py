"""Per-kernel SASS evidence of the Blackwell-native instructions in libdiffsci_b200.so -> profiles/sass_summary.txt

    python tools/sass_summary.py            (cuobjdump -sass on the in-tree library; needs no GPU)

Counts, per kernel: UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), LDTM (tcgen05.ld), UTMALDG (TMA tensor loads),
UTCBAR (tcgen05.commit, .MULTICAST), SYNCS (mbarrier), and HMMA / legacy mma.sync (must be zero)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffsci_b200", "libdiffsci_b200.so")
PAT = {"UTCHMMA.2CTA": r"\bUTCHMMA\.2CTA", "UTCHMMA": r"\bUTCHMMA(?!\.2CTA)", "LDTM": r"\bLDTM", "UTMALDG": r"\bUTMALDG",
       "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS", "HMMA(legacy)": r"\bHMMA", "FFMA": r"\bFFMA"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            counts.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        for k, p in PAT.items():
            if re.search(p, line):
                counts[cur][k] += 1
    arch = re.findall(r"arch = (sm_\w+)", sass)
    out = [f"libdiffsci_b200.so: {len(counts)} kernels, arch {sorted(set(arch))}",
           f"{'kernel':78s} " + " ".join(f"{k:>13s}" for k in PAT)]
    tot = collections.Counter()
    for name, c in counts.items():
        tot.update(c)
        if any(c[k] for k in PAT if k not in ("FFMA", "SYNCS")):
            out.append(f"{name[:78]:78s} " + " ".join(f"{c[k]:13d}" for k in PAT))
    out.append(f"{'TOTAL (all kernels)':78s} " + " ".join(f"{tot[k]:13d}" for k in PAT))
    text = "\n".join(out) + "\n"
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(path, "w").write(text)
    print(text)


if __name__ == "__main__":
    sys.exit(main())
