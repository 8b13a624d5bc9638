"""The C++ host mirror of the reference API (include/dzahui_b200.hpp) over the C ABI: tests/cpp/host_mirror.cpp restates
the reference's own hot-path unit tests (builders, error vocabulary, solver windows).  Compiled here with g++ against the
in-tree library; the --cpu half needs no GPU, the --gpu half runs on the B200 box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror")


def _build(dz):
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("dzahui_b200.hpp", "dzahui_b200.h")]
    so = os.path.join(ROOT, "dzahui_b200", "libdzahui_b200.so")
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(f) for f in [src, so] + hdrs):
        return
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                           "-L", os.path.join(ROOT, "dzahui_b200"), "-ldzahui_b200", "-Wl,-rpath," + os.path.join(ROOT, "dzahui_b200")])


def test_cpp_host_mirror_cpu(dz, tmp_path):
    _build(dz)
    r = subprocess.run([EXE, "--cpu", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host mirror ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_gpu(dz, assets_dir):
    _build(dz)
    r = subprocess.run([EXE, "--gpu", assets_dir], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host mirror ok" in r.stdout, r.stdout + r.stderr
