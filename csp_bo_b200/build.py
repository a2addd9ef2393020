"""Builds the native libraries in-tree (csp_bo_b200/lib/*.so) with nvcc for sm_100a.

    python -m csp_bo_b200.build [--force]

  libcspbo_b200.so        CUDA kernels + C ABI (include/cspbo_b200.h)
  libdot_kernel_b200.so   reference-named shim for cspbo.kernels._dot_kernel  (libdot_builder.py:5-9)
  librbf_kernel_b200.so   reference-named shim for cspbo.kernels._rbf_kernel  (librbf_builder.py:5-12)

nvcc cross-compiles without a GPU; the built .so files are git-ignored but travel to the GPU box.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
INC = os.path.join(ROOT, "include")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = ["api.cu", "pack.cu", "tiles.cu", "gp.cu", "dense.cu", "chol.cu", "so3.cu"]
HEADERS = [os.path.join(SRC, "common.h"), os.path.join(SRC, "shim_common.h"), os.path.join(INC, "cspbo_b200.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build(force=False, verbose_ptxas=False):
    os.makedirs(LIB, exist_ok=True)
    objs, cmds = [], []
    for s in CU_SOURCES:
        src = os.path.join(SRC, s)
        obj = os.path.join(LIB, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + HEADERS):
            cmd = [NVCC] + ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I", INC, "-I", SRC,
                                   "-c", src, "-o", obj]
            if os.environ.get("CSPBO_TUNE"):      # e.g. CSPBO_TUNE="6,2,48": warps per CTA, CTAs per SM, KB per stage
                w, c, kb = os.environ["CSPBO_TUNE"].split(",")
                cmd += ["-DCSPBO_WARPS=" + w, "-DCSPBO_CTAS=" + c, "-DCSPBO_STAGE_KB=" + kb]
            if os.environ.get("CSPBO_DEFS"):      # extra -D switches for experiments, e.g. CSPBO_DEFS="CSPBO_LIBM_EXP"
                cmd += ["-D" + d for d in os.environ["CSPBO_DEFS"].split(",")]
            if verbose_ptxas:
                cmd += ["-Xptxas", "-v"]
            cmds.append(cmd)
    if cmds:
        # the translation units are independent: compile them side by side (tiles.cu, 156 kernel instantiations, takes a
        # minute on its own; the others finish in its shadow)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as pool:
            list(pool.map(_run, cmds))
    main = os.path.join(LIB, "libcspbo_b200.so")
    if force or _stale(main, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", main] + objs +
             ["-L", os.path.join(CUDA_HOME, "lib64"), "-lcusolver", "-lcublas",
              "-Xlinker", "-rpath," + os.path.join(CUDA_HOME, "lib64")])
    for name in ("dot", "rbf"):
        src = os.path.join(SRC, "shim_%s.c" % name)
        out = os.path.join(LIB, "lib%s_kernel_b200.so" % name)
        if force or _stale(out, [src, main] + HEADERS):
            _run(["gcc", "-O2", "-fPIC", "-shared", "-I", INC, "-I", SRC, "-o", out, src,
                  "-L", LIB, "-lcspbo_b200", "-Wl,-rpath,$ORIGIN"])
    return main


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
