"""Builds dzahui_b200/libdzahui_b200.so (the C-ABI library of include/dzahui_b200.h) in-tree with nvcc for sm_100a.

`python -m dzahui_b200.build` or `dzahui_b200.build.build()`.  nvcc cross-compiles without a GPU.
-fmad=false: the FAITHFUL kernels must keep the reference's separate multiply/add roundings; the FAST kernels
call fma() explicitly where they want it."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdzahui_b200.so")
SOURCES = ["dzb_api.cu", "gl_quadrature.cpp", "mesh_ingest.cpp", "csv_writer.cpp"]
HEADERS = ["assembly_kernels.cuh", "tridiag_kernels.cuh", "bigsys_kernels.cuh", "onelaunch_kernels.cuh", "ol_layout.h", "comm_kernels.cuh", "dzb_comm.inl", "euler_kernels.cuh", "mesh2d_kernels.cuh", "gl_quadrature.h", "mesh_ingest.h",
           "gl_tables_generated.h", os.path.join("..", "..", "include", "dzahui_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


STAMP = OUT + ".srchash"


def _source_hash():
    import hashlib
    h = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS) + [os.path.abspath(__file__)]:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if os.path.exists(path):
            h.update(f.encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    """True when the library is missing or was built from other sources than the ones in the tree (by content, not by
    mtime: a copy of the tree — the GPU box's snapshot — keeps contents but not time stamps)."""
    if not os.path.exists(OUT) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # this image's $CXX wrapper lacks parts of the toolchain; the distro g++ is the host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode:
        print(r.stdout)
    if r.returncode:
        raise RuntimeError(f"nvcc failed (see {log})")
    with open(STAMP, "w") as f:
        f.write(_source_hash() + "\n")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
