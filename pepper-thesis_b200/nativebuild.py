"""In-tree native build (no JIT cache): nvcc for sm_100a, gcc/g++ for host pieces.

Artefacts (all git-ignored, all travel to the GPU box with the gpurun snapshot):
  pepper-thesis_b200/libpepper_b200.so          C-ABI library (CUDA kernels), include/pepper_b200.h
  pepper-thesis_b200/libpv_synth.so             synthetic pileup generator (tests/bench input only)
  pepper-thesis_b200/libpv_ingest.so            BAM/FASTA ingest without htslib (host only), include/pepper_ingest.h
  pepper-thesis_b200/build/PEPPER_VARIANT*.so   pybind11 drop-in for the reference's PEPPER_VARIANT module
  oracle/libpv_oracle_port.so, oracle/_ref/*    test oracles (see oracle/Makefile)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

LIB = os.path.join(PKG, "libpepper_b200.so")
SYNTH = os.path.join(PKG, "libpv_synth.so")
INGEST = os.path.join(PKG, "libpv_ingest.so")
PYMOD_DIR = os.path.join(PKG, "build")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout[-4000:], r.stderr[-8000:]))
    if verbose and r.stderr.strip():
        print(r.stderr[-4000:], file=sys.stderr)


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    # fast_inflate.h belongs to the host-only ingest library (build_ingest lists it)
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh")) and f != "fast_inflate.h"]
    return hs + [os.path.join(INC, "pepper_b200.h")]


HOST_SOURCES = ["host_pack.cpp"]      # g++ pieces of libpepper_b200.so (target attributes / immintrin: not for nvcc)


def build_lib(force=False, verbose=False):
    srcs = cuda_sources()
    host_srcs = [os.path.join(CSRC, f) for f in HOST_SOURCES]
    if not force and not _newer(LIB, srcs + host_srcs + headers()):
        return LIB
    objs = []
    for s in host_srcs:
        o = s[:-4] + ".o"
        if force or _newer(o, [s] + headers()):
            _run(["g++", "-O3", "-fPIC", "-std=c++17", "-Wall", "-I", INC, "-I", CSRC, "-c", s, "-o", o], verbose)
        objs.append(o)
    for s in srcs:
        o = os.path.join(CSRC, os.path.basename(s)[:-3] + ".o")
        if force or _newer(o, [s] + headers()):
            extra = os.environ.get("PV_NVCC_FLAGS", "").split()
            _run([NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", *extra,
                  "-I", INC, "-I", CSRC, "-c", s, "-o", o], verbose)
        objs.append(o)
    _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"], verbose)
    return LIB


def build_synth(force=False, verbose=False):
    src = os.path.join(CSRC, "synth_reads.c")
    if force or _newer(SYNTH, [src]):
        _run(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-o", SYNTH, src, "-lm", "-lpthread"], verbose)
    return SYNTH


def build_ingest(force=False, verbose=False):
    """Host-only BAM/FASTA ingest (zlib), include/pepper_ingest.h."""
    src = os.path.join(CSRC, "ingest.cpp")
    if force or _newer(INGEST, [src, os.path.join(CSRC, "fast_inflate.h"), os.path.join(INC, "pepper_ingest.h"), os.path.join(INC, "pepper_b200.h")]):
        _run(["g++", "-O3", "-fPIC", "-shared", "-std=c++17", "-Wall", "-I", INC, src, "-o", INGEST, "-lz", "-lpthread"], verbose)
    return INGEST


def build_pymod(force=False, verbose=False):
    import pybind11
    src = os.path.join(CSRC, "pybind_module.cpp")
    if not os.path.exists(src):
        return None
    os.makedirs(PYMOD_DIR, exist_ok=True)
    init = os.path.join(PYMOD_DIR, "__init__.py")
    if not os.path.exists(init):
        with open(init, "w") as f:
            f.write("# mirrors `from pepper_variant.build import PEPPER_VARIANT` (reference setup.py:74-84)\n")
    out = os.path.join(PYMOD_DIR, "PEPPER_VARIANT" + sysconfig.get_config_var("EXT_SUFFIX"))
    if force or _newer(out, [src, LIB, INGEST, os.path.join(INC, "pepper_ingest.h")] + headers()):
        _run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I", INC, "-I", sysconfig.get_paths()["include"],
              "-I", pybind11.get_include(), src, "-o", out, "-L", PKG, "-lpepper_b200", "-lpv_ingest",
              "-Wl,-rpath,$ORIGIN/.."], verbose)
    return out


def build_oracle(force=False, verbose=False):
    odir = os.path.join(ROOT, "oracle")
    if force:
        subprocess.run(["make", "-C", odir, "clean"], capture_output=True)
    port = os.path.join(odir, "libpv_oracle_port.so")
    if _newer(port, [os.path.join(odir, "region_summary_port.c")]):
        _run(["make", "-C", odir, "port"], verbose)
    ref_dir = os.path.join(odir, "_ref")
    if os.path.isdir("/root/reference/pepper_variant/modules/cpp"):
        have = os.listdir(ref_dir) if os.path.isdir(ref_dir) else []
        shims = {"pv_ref_oracle": "ref_shim.cpp", "pv_ref_polisher": "ref_shim_polisher.cpp", "pv_ref_bam": "ref_shim_bam.cpp",
                 "pv_ref_legacy": "ref_shim_legacy.cpp"}
        stale = False
        for mod, shim in shims.items():
            so = [os.path.join(ref_dir, f) for f in have if f.startswith(mod)]
            if not so or _newer(so[0], [os.path.join(odir, shim), os.path.join(odir, "Makefile")]):
                stale = True
        if stale:
            _run(["make", "-C", odir, "ref"], verbose)


def build_all(force=False, verbose=False):
    build_synth(force, verbose)
    build_ingest(force, verbose)
    build_lib(force, verbose)
    build_pymod(force, verbose)
    build_oracle(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("ok")
