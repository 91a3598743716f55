"""Builds ``locotouch_b200/liblocotouch_b200.so`` (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m locotouch_b200.csrc.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

CSRC = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(CSRC)
ROOT = os.path.dirname(PKG)
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(PKG, "liblocotouch_b200.so")

# Kernels whose masks must be bit-exact against the torch expression order are built without FMA contraction
# (torch evaluates a*b+c as two rounded ops).
PER_FILE_FLAGS = {"mdp_step.cu": ["-fmad=false"], "taxel.cu": ["-fmad=false"], "gae.cu": ["-fmad=false"],
                  "commands.cu": ["-fmad=false"]}

def cutlass_include_dirs() -> list[str]:
    """CUTLASS / CuTe header trees for the tcgen05 GEMM with the fused bias + ELU epilogue (gemm_fused.cu).  The image vendors
    CUTLASS 4.x under flashinfer's data directory; LT_CUTLASS_DIR overrides.  Empty list: the file is built as a stub that
    reports LT_ERR_UNSUPPORTED and the callers keep the cuBLAS + elementwise path."""
    roots = [os.environ.get("LT_CUTLASS_DIR")]
    try:
        import importlib.util
        spec = importlib.util.find_spec("flashinfer")
        if spec is not None and spec.origin:
            roots.append(os.path.join(os.path.dirname(spec.origin), "data", "cutlass"))
    except Exception:
        pass
    for r in roots:
        if r and os.path.exists(os.path.join(r, "include", "cutlass", "cutlass.h")):
            return [os.path.join(r, "include"), os.path.join(r, "tools", "util", "include")]
    return []


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=default",
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps() -> list[str]:
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    variant = os.environ.get("LT_LIB_VARIANT")  # tuning builds: liblocotouch_b200.<variant>.so with LT_MDP_* knobs, next to the product
    if variant:
        return _build(True, verbose, variant)
    if not force and not needs_build():
        return LIB_PATH
    return _build(force, verbose, None)


def _build(force: bool, verbose: bool, variant: str | None) -> str:
    lib_path = LIB_PATH if not variant else LIB_PATH[:-3] + f".{variant}.so"
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    objs = []
    build_dir = os.path.join(CSRC, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in sources():
        knob_files = os.environ.get("LT_VARIANT_FILES", "mdp_step.cu").split(",")  # files the LT_MDP_* / LT_TAXEL_* knobs touch
        knobbed = variant and os.path.basename(src) in knob_files
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + (f".{variant}.o" if knobbed else ".o"))
        objs.append(obj)
        if (not force or (variant and not knobbed)) and os.path.exists(obj) and all(os.path.getmtime(obj) > os.path.getmtime(d) for d in [src] + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))):
            continue
        extra = [f"-D{k}={v}" for k, v in os.environ.items() if k.startswith(("LT_MDP_", "LT_TAXEL_"))] if (knobbed or not variant) else []  # tuning knobs
        if os.path.basename(src) in ("gemm_fused.cu", "wgrad_tc.cu", "mlp_fused.cu"):
            dirs = cutlass_include_dirs()
            extra += ["--expt-extended-lambda", "-DLT_HAVE_CUTLASS=1"] + [x for d in dirs for x in ("-I", d)] if dirs else ["-DLT_HAVE_CUTLASS=0"]
        cmd = [nvcc, *NVCC_FLAGS, *PER_FILE_FLAGS.get(os.path.basename(src), []), *extra, "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path, *objs, "-cudart", "static"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    return lib_path


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
