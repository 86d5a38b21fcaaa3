"""
In-tree build of libpfx_b200.so (nvcc, sm_100a only) and of the oracle/test helpers.

`python -m phasefieldx_b200.build` or `__graft_entry__.build()`.  The shared object lands next
to this file (git-ignored, but shipped to the GPU box by gpurun).  nvcc cross-compiles without
a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpfx_b200.so")
HOSTCHECK = os.path.join(ROOT, "tests", "native", "libpfx_hostcheck.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "550"]


def _nvcc():
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found: libpfx_b200.so cannot be built (there is no CPU fallback)")
    return nv


def _metis():
    """METIS 5 static library shipped with the CUDA toolkit (graph partitioner of pfx_partition_graph_cell_rank)."""
    for d in (os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "lib64"), "/usr/local/cuda/lib64",
              "/usr/local/cuda/targets/x86_64-linux/lib"):
        p = os.path.join(d, "libmetis_static.a")
        if os.path.exists(p):
            return p
    raise RuntimeError("libmetis_static.a (CUDA toolkit) not found")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "pfx_b200.h")]


def build(force=False, verbose=False, ptxas_log=None):
    """Compile the CUDA library.  Returns the path of the shared object."""
    if not force and not _newer(LIB, sources()):
        return LIB
    nv = _nvcc()
    bdir = os.path.join(ROOT, "build")
    os.makedirs(bdir, exist_ok=True)
    objs, procs = [], []
    for src in ("pfx_engine.cu", "pfx_blocks.cpp", "pfx_format.cpp"):
        obj = os.path.join(bdir, src.rsplit(".", 1)[0] + ".o")
        cmd = [nv] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_log else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = ""
    for src, p in procs:
        out, _ = p.communicate()
        log += out
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    if ptxas_log:
        with open(ptxas_log, "w") as fh:
            fh.write(log)
    cmd = [nv, "-shared", "-o", LIB] + objs + [_metis(), "-ldl", "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


def build_hostcheck(force=False):
    """Test-only host build of the element headers + block builder (g++; see tests/native)."""
    srcs = [os.path.join(ROOT, "tests", "native", "hostcheck.cpp"), os.path.join(CSRC, "pfx_blocks.cpp")]
    deps = srcs + sources()
    if not force and not _newer(HOSTCHECK, deps):
        return HOSTCHECK
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", HOSTCHECK] + srcs + [_metis()]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed:\n{r.stdout}")
    return HOSTCHECK


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_log=os.path.join(ROOT, "build", "ptxas.log")))
    print(build_hostcheck(force="--force" in sys.argv))
