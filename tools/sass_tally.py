"""Per-kernel tally of the Blackwell-specific SASS opcodes in libvt_b200.so (B200_PROFILING.md "What proves a Blackwell-native
kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP/UTMAPF, cluster barriers, DSMEM.

    python tools/sass_tally.py > profiles/r02_sass_opcodes.txt          (CPU box: cuobjdump only)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "vla_touch_b200", "lib", "libvt_b200.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "UCGABAR",
            "SYNCS", "FENCE.VIEW.ASYNC", "LDG.E.STRONG", "REDG", "CCTL", "MUFU", "HMMA", "FFMA2", "FADD2", "FMUL2"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kern, tally = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*$", "", kern)
            tally[kern] = collections.Counter()
            continue
        if kern is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        tally[kern]["instructions"] += 1
        for p in PATTERNS:
            if op.startswith(p):
                tally[kern][p + (".2CTA" if ".2CTA" in op else "")] += 1
    print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)} ({os.path.getsize(SO)} bytes): Blackwell opcodes per kernel")
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UTMAPF", "UBLKPF", "UTCBAR", "UTCBAR.2CTA", "UCGABAR", "SYNCS",
            "FENCE.VIEW.ASYNC", "LDG.E.STRONG", "REDG", "CCTL", "MUFU", "FFMA2", "HMMA"]
    tot = collections.Counter()
    for k, c in tally.items():
        if not any(c[p] for p in cols if p not in ("MUFU", "FFMA2", "SYNCS")):
            continue
        print(f"\n{k}   [{c['instructions']} instructions]")
        print("   " + ", ".join(f"{p} {c[p]}" for p in cols if c[p]))
        tot.update(c)
    print("\nTOTAL over tensor-core / TMA kernels: " + ", ".join(f"{p} {tot[p]}" for p in cols if tot[p]))
    print(f"kernels in the library: {len(tally)}; legacy HMMA (mma.sync / wmma) instructions anywhere: {sum(c['HMMA'] for c in tally.values())}")


if __name__ == "__main__":
    main()
