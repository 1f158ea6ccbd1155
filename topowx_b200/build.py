"""Build libtwxi.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m topowx_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtwxi.so")
SOURCES = ["api.cu", "knn.cu", "setup.cu", "ked.cu", "ked_warp.cu", "ked_rl.cu", "gwr.cu", "fixer.cu", "peak.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--use_fast_math=false"]
NVCC_FLAGS = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]   # never fast-math: FP64 parity


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "twxi.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False, extra=(), out=None, tag=""):  # noqa: C901
    """extra/out/tag: instrumented developer builds (e.g. extra=["-DTWXI_KED_PROFILE"], out="libtwxi_prof.so")."""
    target = os.path.join(HERE, out) if out else LIB
    if not force and target == LIB and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", tag + ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [_nvcc(), "-shared", "-o", target] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    if "--prof" in sys.argv:
        print(build_lib(force=True, extra=["-DTWXI_KED_PROFILE"], out="libtwxi_prof.so", tag="_prof"))
    else:
        print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
