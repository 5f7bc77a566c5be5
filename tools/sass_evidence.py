"""Regenerates profiles/<tag>_sass_evidence.txt: counts of the data-movement / synchronisation / arithmetic SASS
mnemonics per kernel family of huygens_b200/lib/libhuygens_cuda.so (cuobjdump -sass; no GPU needed).

    python tools/sass_evidence.py r2
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "huygens_b200", "lib", "libhuygens_cuda.so")
FAMILIES = ["fb_exact_ws_kernel", "fb_exact_ru_kernel", "fb_exact_kernel", "osc_kernel", "rms_ws_kernel"]
KEEP = re.compile(r"^(FMUL|FADD|FFMA|DFMA|DMUL|DADD|LDS\.128|STS\.128|SYNCS\.\S+|LDG\.E\.\d+\.STRONG\.SYS|STG\.E\.\d+\.STRONG\.SYS|"
                  r"ATOMG\.\S+|UBLKCP\.\S+|MUFU\.RCP\S*|STAS\S*|UCGABAR\S*)$")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = {f: collections.Counter() for f in FAMILIES}
    archs, fam = set(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fam = next((f for f in FAMILIES if re.search(r"\d+" + f + r"I", m.group(1))), None)
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            archs.add(m.group(1))
        if fam is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and KEEP.match(m.group(1)):
            counts[fam][m.group(1)] += 1
    out = [f"# SASS evidence (cuobjdump -sass huygens_b200/lib/libhuygens_cuda.so, sm_100a), round {tag[1:]} — tools/sass_evidence.py",
           "# counts of the Blackwell-era data-movement / synchronisation instructions per kernel family (all template instances)", ""]
    for f in FAMILIES:
        out.append(f"== {f}")
        out += [f"{n:7d} {k}" for k, n in counts[f].most_common()]
        out.append("")
    out += ["== arch", "  " + " ".join(sorted(archs))]
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_evidence.txt"), "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("\n".join(out[:24]))


if __name__ == "__main__":
    main()
