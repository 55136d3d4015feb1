"""Build recipe for libbnn_b200.so (hand-written sm_100a CUDA + the C ABI).

    python -m bnn_b200.build          # or __graft_entry__.build()

nvcc cross-compiles without a GPU.  The library is built IN-TREE
(bnn_b200/libbnn_b200.so) so that it travels with the repo snapshot; it links
the CUDA runtime statically and has no torch / Python dependency.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbnn_b200.so")
SOURCES = ["abi.cu", "bbb_train.cu", "dropout_train.cu", "forward_simt.cu", "forward_tc.cu", "forward_tcl.cu", "moments.cu", "sampling.cu", "sgld_step.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, trace=False):
    """trace=True builds libbnn_b200_trace.so: the same library with the in-kernel timeline probes of the tensor-core
    forward compiled in (-DBNN_TC_TRACE_BUILD; debugging only, see tools/tc_trace.py)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tag = "trace" if trace else os.environ.get("BNN_BUILD_TAG", "")     # BNN_BUILD_TAG=x: experiment build libbnn_b200_x.so
    objdir = os.path.join(HERE, "build_" + tag if tag else "build")
    lib = LIB.replace(".so", "_" + tag + ".so") if tag else LIB
    flags = NVCC_FLAGS + (["-DBNN_TC_TRACE_BUILD"] if trace else [])
    for knob in ("BNN_SG_MI", "BNN_TC_SETS", "BNN_TC_REG_E1", "BNN_TC_REG_E2", "BNN_TC_REG_MMA", "BNN_TC_REG_SLACK", "BNN_DT_THREADS"):   # experiment knobs (compile-time)
        if os.environ.get(knob):
            flags = flags + [f"-D{knob}=" + os.environ[knob]]
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "bnn_b200.h"))
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            subprocess.run(cmd, check=True)
    if force or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-cudart", "static", "-o", lib] + objs + ["-ldl"]
        subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
