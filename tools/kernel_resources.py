"""Per-kernel register / stack / static shared memory table of libctk.so (cuobjdump --dump-resource-usage)."""
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "torch-unet_b200/ctk/libctk.so"
txt = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", txt):
    d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    d = d.replace("(anonymous namespace)::", "")
    d = re.sub(r"\(.*", "", d)
    print(d[:80].ljust(82), "REG", m.group(2).rjust(3), "STACK", m.group(3).rjust(3), "SMEM", m.group(4))
