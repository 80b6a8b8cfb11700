"""Per-kernel SASS evidence that the hot kernels are Blackwell-native: counts of tcgen05 MMA (UTCHMMA, .2CTA for CTA
pairs), TMEM loads (LDTM), TMA loads / stores (UTMALDG / UTMASTG), mbarrier waits (SYNCS), fp32 atomics (ATOM / RED .F32
-- there must be none in the training reductions) per kernel of libctk.so.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "torch-unet_b200", "ctk", "libctk.so")
PATTERNS = collections.OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"),
    ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"), ("UTCBAR", r"\bUTCBAR"), ("HMMA(legacy)", r"\bHMMA"),
    ("ATOM/RED.F32", r"\b(ATOM|RED|ATOMG|ATOMS)\S*\.F32"), ("ATOM/RED.F64", r"\b(ATOM|RED|ATOMG)\S*\.F64"),
    ("FFMA", r"\bFFMA"), ("DADD/DFMA", r"\b(DADD|DFMA)")])


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS.items():
            if re.search(pat, line):
                cur[name] += 1
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            cur["instructions"] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    names = list(PATTERNS)
    print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)}: {len(kernels)} kernels; instruction-mnemonic counts per kernel")
    print("# " + " | ".join(["kernel"] + names + ["instructions"]))
    tot = collections.Counter()
    for (mangled, c), pretty in zip(kernels.items(), dem):
        short = re.sub(r"\(anonymous namespace\)::", "", pretty)
        short = re.sub(r"\(.*", "", short)[:70]
        print(f"{short:70s} " + " ".join(f"{c[n]:6d}" for n in names) + f" {c['instructions']:7d}")
        tot.update(c)
    print(f"{'TOTAL':70s} " + " ".join(f"{tot[n]:6d}" for n in names) + f" {tot['instructions']:7d}")


if __name__ == "__main__":
    sys.exit(main())
