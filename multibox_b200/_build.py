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
SOURCES = ["mbx_api.cu", "mbx_match.cu", "mbx_match_reg.cu", "mbx_match_reg_w1.cu", "mbx_match_reg_w2.cu",
           "mbx_match_reg_w4.cu", "mbx_match_reg_w8.cu", "mbx_match_reg_w16.cu", "mbx_detect.cu"]
LIB = os.path.join(HERE, "libmultibox_b200.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--fmad=false",            # fp32/fp64 cost arithmetic must not be contracted; FMAs are explicit
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]
LINK_FLAGS = ["--shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
        [os.path.join(HERE, "..", "include", "multibox_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    src, obj, extra = args
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return " ".join(cmd) + "\n" + proc.stdout, proc.returncode


def build(force=False, verbose=False, extra_flags=(), lib=None):
    """Compiles every translation unit (in parallel) for sm_100a and links the shared library."""
    lib = lib or LIB
    if not force and lib == LIB and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = OBJ_DIR + ("_" + "".join(c for c in "".join(extra_flags) if c.isalnum()) if extra_flags else "")
    os.makedirs(obj_dir, exist_ok=True)
    jobs = [(os.path.join(CSRC, s), os.path.join(obj_dir, s[:-3] + ".o"), tuple(extra_flags)) for s in SOURCES]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(_compile_one, jobs))
    text = "".join(r[0] for r in results)
    rc = 1 if any(r[1] != 0 for r in results) else 0      # (a compiler killed by a signal returns < 0)
    if rc == 0:
        cmd = [nvcc_path()] + LINK_FLAGS + ["-o", lib] + [j[1] for j in jobs]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        text += " ".join(cmd) + "\n" + proc.stdout
        rc = proc.returncode
    log = os.path.join(HERE, "csrc", "build.log")
    with open(log, "w") as f:
        f.write(text)
    if verbose or rc != 0:
        print(text)
    if rc != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return lib


if __name__ == "__main__":
    print(build(force=True, verbose=True))
