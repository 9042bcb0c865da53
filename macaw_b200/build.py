"""Builds macaw_b200/libmacaw_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

  python -m macaw_b200.build [--force]

Flags that matter: -gencode arch=compute_100a,code=sm_100a (Blackwell only), -lineinfo (ncu source
pages), -fmad=false and -Xcompiler -ffp-contract=off (no FMA contraction on device or host so FP64
arithmetic is bit-identical to the CPU oracle).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmacaw_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
         "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-Xptxas", "-v"]
# (source, object, extra flags): the kernels are compiled twice, without and with the rarely used branches
# (csrc/mcw_kernels.h: namespaces mcw_r0 / mcw_r1)
SOURCES = [("mcw_kernels.cu", "mcw_kernels_r0.o", ["-DMCW_RARE=0"]), ("mcw_kernels.cu", "mcw_kernels.o", ["-DMCW_RARE=1"]),
           ("mcw_api.cu", "mcw_api.o", [])]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def needs_build():
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "macaw_b200.h")]
    return _newest(deps) > os.path.getmtime(LIB)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs = []

    def compile_one(item):
        src, oname, extra = item
        obj = os.path.join(HERE, "build", oname)
        cmd = [NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log[-4000:]}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


def build_variant(name, defines=()):
    """Kernel A/B experiments: the same sources with extra -D flags into build/variants/<name>.so.
    Selected at run time with MCW_LIB_PATH (macaw_b200/lib.py); never used by tests or the bench line."""
    out_dir = os.path.join(HERE, "variants")
    os.makedirs(os.path.join(out_dir, name), exist_ok=True)
    def compile_one(item):
        src, oname, extra = item
        obj = os.path.join(out_dir, name, oname)
        cmd = [NVCC] + FLAGS + extra + [f"-D{d}" for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        open(obj + ".log", "w").write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError((r.stdout + r.stderr)[-4000:])
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    lib = os.path.join(out_dir, f"libmacaw_b200_{name}.so")
    subprocess.run([NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs,
                   check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
