"""Build libctk from the sources of another git revision, for same-box A/B runs of a kernel change:

    python tools/build_alt_lib.py <git-ref> <name>     ->  torch-unet_b200/ctk/libctk_<name>.so
    CTK_LIB=torch-unet_b200/ctk/libctk_<name>.so python bench.py ...

(The C ABI has to match the Python binding of the working tree; the .so is git-ignored and travels with gpurun.)
"""
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-unet_b200"))
import build as B  # noqa: E402

ref, name = sys.argv[1], sys.argv[2]
out = os.path.join(ROOT, "torch-unet_b200", "ctk", f"libctk_{name}.so")
with tempfile.TemporaryDirectory() as tmp:
    tar = subprocess.run(["git", "-C", ROOT, "archive", ref, "torch-unet_b200/csrc", "include"], capture_output=True, check=True).stdout
    subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
    csrc = os.path.join(tmp, "torch-unet_b200", "csrc")
    srcs = [f for f in sorted(os.listdir(csrc)) if f.endswith(".cu")]

    def one(src):
        obj = os.path.join(tmp, src.replace(".cu", ".o"))
        r = subprocess.run([B.NVCC, *[f for f in B.FLAGS if f not in ("-Xptxas", "-v")], "-c", os.path.join(csrc, src), "-o", obj],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, srcs))
    subprocess.run([B.NVCC, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
print(out)
