"""Per-kernel counts of the SASS mnemonics that prove what each kernel is made of (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, dp2a/dp4a -> IDP, three-input max -> VIMNMX3, barriers, atomics.
    python tools/sass_counts.py [lib.so] > profiles/sass_counts.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "pixelbox_b200/lib/libpixelbox_b200.so"
PAT = [("UTCIMMA", r"\bUTCIMMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"),
       ("IDP.2A", r"\bIDP\.2A"), ("IDP.4A", r"\bIDP\.4A"), ("VIMNMX3", r"\bVIMNMX3"), ("SYNCS", r"\bSYNCS"), ("ATOMG/RED", r"\b(ATOMG|RED)\b"),
       ("LDG.128", r"\bLDG\.E\.128"), ("SHFL", r"\bSHFL"), ("HMMA/IMMA (legacy mma.sync)", r"\b(HMMA|IMMA)")]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("pbx::", "")
        counts[name] = collections.Counter()
        continue
    if name is None or "/*" not in line:
        continue
    counts[name]["instructions"] += 1
    for label, pat in PAT:
        if re.search(pat, line):
            counts[name][label] += 1
print(f"# SASS mnemonic counts per kernel of {so} (cuobjdump -sass); 0 columns omitted")
for k, c in counts.items():
    extra = "  ".join(f"{lab}={c[lab]}" for lab, _ in PAT if c[lab])
    print(f"{k:60s} instr={c['instructions']:6d}  {extra}")
