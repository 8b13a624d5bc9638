"""CPU-only: the C-ABI library loads and exports every symbol include/dzahui_b200.h declares, the host-side logic
(quadrature table, params builders, ingestion parser errors, synthetic workloads) is right, and — there being no
GPU here — compute entry points fail loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "dzahui_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dzb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(dz):
    from dzahui_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 35
    lib = C.CDLL(_lib.SO_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dzahui_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dzb_[a-z0-9_]+)", out))
    assert set(syms) <= exported


def test_library_is_sm100a_only(dz):
    from dzahui_b200 import _lib
    r = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", r.stdout))
    assert archs == {"100a"}, archs


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "dzahui_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in txt and "libdzahui_oracle" not in txt and "dzahui_oracle" not in txt, f


def test_quadrature_matches_oracle_bit_for_bit(dz, oracle, golden):
    for G in (2, 3, 7, 50, 99, 100, 101, 150, 350, 600, 1001):
        x, w = dz.quad_table(G)
        xo, wo = oracle.quad_table(G)
        assert np.array_equal(x, xo) and np.array_equal(w, wo), G
        if str(G) in golden["quad"]:
            assert np.array_equal(x, golden["quad"][str(G)]["x"]) and np.array_equal(w, golden["quad"][str(G)]["w"])
    for n, k in [(5, 1), (100, 99), (101, 51), (101, 52), (350, 349), (350, 1)]:
        assert dz.quad_pair(n, k) == oracle.quad_pair(n, k)
    x, w = dz.quad_table(1)
    assert x.size == 0                                 # `for j in 1..1` is empty: no error
    x, w = dz.quad_table(0)
    assert x.size == 0


def test_quad_pair_misuse_is_integration_error(dz):
    with pytest.raises(dz.Integration) as e:
        dz.quad_pair(150, 150)
    assert "k should be smaller than n" in str(e.value)
    with pytest.raises(dz.Integration):
        dz.quad_pair(10, 11)


def test_param_builders_mirror_reference(dz):
    p = dz.StokesParams.static_pressure().hydrostatic_pressure(100.0).density(1.0).force_function(lambda _x: -10.0).build()
    assert (p.rho, p.hydrostatic_pressure, p.force_function(3.0)) == (1.0, 100.0, -10.0)
    assert "f(0) -> -10.0" in repr(p)
    with pytest.raises(dz.BuilderPanic, match="Params lack 'pressure' term!"):
        dz.StokesParams.normal_1d().density(1.0).force_function(lambda x: x).build()
    with pytest.raises(dz.BuilderPanic, match="Params lack 'density' term!"):
        dz.StokesParams.normal_1d().hydrostatic_pressure(1.0).force_function(lambda x: x).build()
    with pytest.raises(dz.BuilderPanic, match="Params lack force_function!"):
        dz.StokesParams.normal_1d().hydrostatic_pressure(1.0).density(1.0).build()
    q = dz.DiffussionParams.time_independent().b(1.0).mu(2.0).boundary_conditions(0.0, 1.0).build()
    assert (q.mu, q.b, q.boundary_conditions) == (2.0, 1.0, (0.0, 1.0))
    with pytest.raises(dz.BuilderPanic, match="Params lack 'mu' term!"):
        dz.DiffussionParams.time_independent().b(1.0).boundary_conditions(0, 1).build()
    with pytest.raises(dz.BuilderPanic, match="Params lack boundary conditions!"):
        dz.DiffussionParams.time_independent().b(1.0).mu(1.0).build()
    t = dz.DiffussionParams.time_dependent().b(1.0).mu(1.0).boundary_conditions(1.0, 0.0).initial_conditions([15.0] * 1).build()
    assert t.initial_conditions == [15.0]
    with pytest.raises(dz.BuilderPanic, match="Params lack initial conditions!"):
        dz.DiffussionParams.time_dependent().b(1.0).mu(1.0).boundary_conditions(1.0, 0.0).build()
    base = dz.DiffussionParams.time_dependent().mu(1.0)
    assert base.b(2.0) is not base and base._b is None  # setters return new builders, like `Self { .., ..self }`


def test_no_cpu_fallback(dz):
    """Without a GPU every compute entry point must raise CudaError; with one this test has nothing to check."""
    if dz.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(dz.CudaError):
        dz.Mesh.from_nodes([0.0, 0.5, 1.0])
    with pytest.raises(dz.CudaError):
        dz.EnsembleTimeDependent(np.array([[0.0, 0.5, 1.0]]), 1.0, 1.0, (0.0, 1.0), np.array([[0.0]]), 150)


def test_ingest_parse_errors_precede_cuda(dz, tmp_path, assets_dir):
    """Parsing is host code: reference error classes surface even without a device."""
    with pytest.raises(dz.Io):
        dz.Mesh.builder(str(tmp_path / "nope.obj")).build_mesh_1d()
    with pytest.raises(dz.MeshParse, match="Only coordinates over a line"):
        dz.Mesh.builder(assets_dir + "/big_mesh.obj").build_mesh_1d()
    p = tmp_path / "bad.obj"
    p.write_text("v 0.0 0.0 1.0\nv 0.0 0.0\n")
    with pytest.raises(dz.MeshParse, match="3 elements"):
        dz.Mesh.builder(str(p)).build_mesh_1d()
    p.write_text("v 0.0 0.0 0x10\n")
    with pytest.raises(dz.MeshParse):
        dz.Mesh.builder(str(p)).build_mesh_1d()


def test_synthetic_workloads_are_deterministic(dz):
    from dzahui_b200 import synthetic
    m = synthetic.ensemble_meshes([0, 1, 65535], 1024)
    assert m.shape == (3, 1024) and np.all(m[:, 0] == 0.0) and np.all(m[:, -1] == 1.0)
    assert np.all(np.diff(m, axis=1) > 0)
    h = np.diff(m, axis=1)
    assert h.min() >= 0.8 / 1023 - 1e-15 and h.max() <= 1.2 / 1023 + 1e-15
    again = synthetic.ensemble_meshes([65535], 1024)
    assert np.array_equal(again[0], m[2])
    # splitmix64 known answer: first output of the stream seeded with 0
    assert int(synthetic.splitmix64(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF
    ic = synthetic.ensemble_initial_conditions(m, [0, 1, 65535])
    assert ic.shape == (3, 1022)
    x = synthetic.regular_bar(11)
    assert x[0] == 0.0 and x[-1] == 1.0 and x[3] == 3 / 10


def test_csv_writer_matches_rust_formatting(dz, tmp_path):
    """Writer::write (writer.rs:99-146): header, one value per line for 1 variable, Rust Display for f64"""
    cases = [(1.0, "1"), (-2.0, "-2"), (0.5, "0.5"), (0.1 + 0.2, "0.30000000000000004"), (1e21, "1000000000000000000000"),
             (1e-7, "0.0000001"), (-0.0, "-0"), (0.0, "0"), (100.9937043619243, "100.9937043619243"),
             (123456789.125, "123456789.125"), (5e-324, "0." + "0" * 323 + "5"), (float("inf"), "inf"), (float("-inf"), "-inf"),
             (1.7976931348623157e308, "17976931348623157" + "0" * 292), (15.0, "15"),
             (0.9485126528155906, "0.9485126528155906"), (2.5e-5, "0.000025"), (1234.5, "1234.5")]
    for v, want in cases:
        assert dz.format_f64(v) == want, (v, dz.format_f64(v))
    assert dz.format_f64(float("nan")) == "NaN"
    for v in np.random.default_rng(0).normal(size=200) * 10.0 ** np.random.default_rng(1).integers(-12, 12, 200):
        s = dz.format_f64(v)
        assert "e" not in s and float(s) == v                         # positional and round-trips
    (tmp_path / "old.csv").write_text("x")
    (tmp_path / "sub").mkdir()
    w = dz.Writer(str(tmp_path), "pressure_", ["p"], erase_prev_dir=True)
    assert not (tmp_path / "old.csv").exists() and (tmp_path / "sub").exists()
    path = w.write(12.5, [109.96852572142232, 100.0, -0.0])
    assert path.endswith("pressure_12.5.csv")
    assert open(path).read() == "p\n109.96852572142232\n100\n-0\n"
    w2 = dz.Writer(str(tmp_path), "v", ["v_x", "v_y"])
    assert open(w2.write(3.0, [1.0, 2.0, 3.0, 4.0, 5.0])).read() == "v_x,v_y\n1,2\n3,4\n5\n"   # chunks(): short last line
    with pytest.raises(dz.Io):
        dz.Writer(str(tmp_path / "missing"), "a", ["p"])


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU port of the reference on the host cores) honours the bench contract"""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-bars", "32"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_exact_division_sequence_is_correctly_rounded():
    """exact_div() of tridiag_kernels.cuh (quotient by a precomputed correctly rounded reciprocal + two corrections, used
    by the warp-per-bar FAITHFUL sweeps) replayed in exact rational arithmetic: the bits of a / den, every time"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("check_exact_div", os.path.join(ROOT, "tools", "check_exact_div.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    bad, inexact = mod.run(seed=2026, cases=30000)
    assert bad == 0 and inexact == 0


def test_header_is_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: include/dzahui_b200.h must compile as C99 (no C++-isms, no torch types)"""
    src = tmp_path / "consumer.c"
    src.write_text('#include "dzahui_b200.h"\nint main(void) { dzb_options o; dzb_mesh* m = 0; (void)o; (void)m; return DZB_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                           str(src)])


def test_rust_shim_matches_header():
    """INTEGRATION.md's `extern "C"` block is uncompiled (no rustc in the image): keep it honest mechanically.  Every
    `pub fn dzb_*` it declares must exist in include/dzahui_b200.h with the same number of parameters, and the Rust
    `DzbOptions` must list the header's `dzb_options` fields in the header's order."""
    import re
    hdr = open(os.path.join(ROOT, "include", "dzahui_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    rust = doc[doc.index('extern "C" {'):]
    rust = rust[:rust.index("\n}\n")]
    rust = re.sub(r"//[^\n]*", "", rust)

    def split_args(text):           # top-level commas only (a Rust fn-pointer parameter has commas of its own)
        out, depth, cur = [], 0, ""
        for ch in text:
            depth += ch == "("
            depth -= ch == ")"
            if ch == "," and depth == 0:
                out.append(cur)
                cur = ""
            else:
                cur += ch
        return [a for a in out + [cur] if a.strip() and a.strip() != "void"]

    c_fns = {m.group(1): split_args(m.group(2)) for m in re.finditer(r"\b(dzb_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)}
    r_fns = {m.group(1): split_args(m.group(2)) for m in re.finditer(r"pub fn (dzb_\w+)\s*\((.*?)\)\s*(?:->\s*[\w:* ]+)?;", rust, flags=re.S)}
    assert len(r_fns) >= 30
    for name, args in r_fns.items():
        assert name in c_fns, f"{name} is not in the header"
        assert len(args) == len(c_fns[name]), f"{name}: {len(args)} parameters in the Rust block, {len(c_fns[name])} in the header"
    c_opt = re.search(r"typedef struct dzb_options \{(.*?)\} dzb_options;", hdr, flags=re.S).group(1)
    c_fields = re.findall(r"int\s+(\w+)\s*;", c_opt)
    r_opt = re.search(r"pub struct DzbOptions \{(.*?)\}", doc, flags=re.S).group(1)
    r_fields = re.findall(r"pub (\w+): c_int", r_opt)
    assert c_fields == r_fields and len(c_fields) == 4, (c_fields, r_fields)
    # the status codes
    for name, val in re.findall(r"pub const (DZB_\w+): c_int = (\d+);", doc):
        assert re.search(rf"\b{name} = {val}\b", hdr), name
