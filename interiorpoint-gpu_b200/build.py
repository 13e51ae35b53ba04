"""Build libipm_b200.so (sm_100a only) in-tree with nvcc.  No JIT cache: the .so travels with the repo
snapshot to the GPU box."""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libipm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """Default: lib/libipm_b200.so.  `variant` = NAME builds lib/variants/libNAME.so with the extra -D `defines` (same-box
    A/B runs select it with IPM_B200_LIB=...; e.g. --variant dagtiming -DIPM_DAG_TIMING for tools/potrf_dag_timing.py)."""
    lib, obj_dir = LIB, LIB_DIR
    if variant:
        obj_dir = os.path.join(LIB_DIR, "variants", variant + ".obj")
        lib = os.path.join(LIB_DIR, "variants", f"lib{variant}.so")
        os.makedirs(obj_dir, exist_ok=True)
    elif not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, *defines, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    name = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=name,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
