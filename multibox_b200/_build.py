"""Builds libmultibox_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

No torch headers are involved: the library's boundary is the plain C ABI in
include/multibox_b200.h.  The built .so is git-ignored but travels with the
repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["mbx_api.cu", "mbx_match.cu", "mbx_match_reg.cu", "mbx_detect.cu"]
LIB = os.path.join(HERE, "libmultibox_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--fmad=false",            # fp32/fp64 cost arithmetic must not be contracted; FMAs are explicit
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--shared",
    "-cudart", "shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
        [os.path.join(HERE, "..", "include", "multibox_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "csrc", "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
