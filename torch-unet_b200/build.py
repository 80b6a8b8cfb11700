"""Build libctk.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python torch-unet_b200/build.py [--force]

The .so lands in torch-unet_b200/ctk/ so it travels to the GPU box with the snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "ctk", "libctk.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xptxas", "-v", "--expt-relaxed-constexpr"]
SOURCES = ["ctk_common.cu", "pearson.cu", "prep.cu", "conv_first.cu", "conv_tc.cu", "gemm_tc.cu", "head.cu", "optim.cu", "bn_train.cu", "head_train.cu", "wgrad_tc.cu", "first_train.cu", "tile_metrics.cu", "prep_input.cu", "ssim.cu", "reduce.cu", "f32_train.cu"]


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/ctk.h"]:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        r = subprocess.run([NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = []
    for src, obj, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(f"--- {src}\n{r.stderr}\n")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    r = subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr)
        raise RuntimeError("link failed")
    open(stamp, "w").write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
