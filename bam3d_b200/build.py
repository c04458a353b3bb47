"""Build the CUDA shared library (C ABI in include/bam3d_gpu.h) in-tree for sm_100a with nvcc.

`-fmad=false` is part of the numerics contract (bit-parity with the reference's scalar f32 arithmetic);
division and sqrt keep nvcc's IEEE defaults (-prec-div=true -prec-sqrt=true). No GPU is needed to build.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbam3d_gpu.so")
SOURCES = ["narrow.cu", "broad.cu", "capi.cu", "comm.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "bam3d_gpu.h"), __file__]
    if not force and _newer(LIB, deps):
        return LIB
    objs, procs = [], []
    for src in SOURCES:  # the translation units compile side by side
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stderr=subprocess.PIPE, text=True)))
        objs.append(obj)
    for src, p in procs:
        err = p.communicate()[1]
        if p.returncode != 0 or verbose:
            sys.stderr.write(err)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}")
    subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"], check=True)
    return LIB


def build_variant(name, extra_flags=(), csrc=CSRC):
    """Development: build the same sources with extra nvcc flags (e.g. -DB3_...) into build/variants/lib_<name>.so,
    for A/B timing with tests/tools/ab_time.py. The product library is untouched."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out_dir = os.path.join(HERE, "..", "build", "variants")
    obj_dir = os.path.join(out_dir, "obj_" + name)
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        procs.append(subprocess.Popen([nvcc] + NVCC_FLAGS + list(extra_flags) + ["-I", os.path.join(HERE, "..", "include"), "-c",
                                                                                 os.path.join(csrc, src), "-o", obj]))
        objs.append(obj)
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed for variant " + name)
    lib = os.path.abspath(os.path.join(out_dir, f"lib_{name}.so"))
    subprocess.run([nvcc, "-shared", "-o", lib] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"], check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
