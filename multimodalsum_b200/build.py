"""In-tree build of libmmsum_b200.so (nvcc, sm_100a only).

`python -m multimodalsum_b200.build` or `__graft_entry__.build()`.  The .so is git-ignored but travels
to the GPU box with the repo snapshot; nothing is JIT-compiled at run time.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmmsum_b200.so")
SOURCES = ["gemm_sm100.cu", "attention_sm100.cu", "decode_sm100.cu", "beam_sm100.cu", "rowwise.cu", "optim.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr",
]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "mmsum_b200.h")]
    if not force and os.path.exists(LIB) and not any(_newer(d, LIB) for d in deps if os.path.exists(d)):
        return LIB
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and %s is missing or stale" % LIB)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(CSRC, os.path.basename(s).replace(".cu", ".o"))
        objs.append(o)
        extra = ["-DMMSUM_ATTN_TRACE"] if os.environ.get("MMSUM_TRACE") else []
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
