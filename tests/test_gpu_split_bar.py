"""GPU tests of the split bar (BASELINE config 4 at N > 1) through the library's own communicator: `dzb_comm_*` and the
collective `dzb_solver_block_solve_device` (include/dzahui_b200.h).  Reference semantics: ONE solve_by_thomas over the
whole system (matrix_solver/mod.rs:17-48) after the assembly of time_independent.rs:91-182.
A one-rank communicator runs everywhere; the two-rank runs need two GPUs (torchrun, NCCL + NVLink peer mailboxes) and are
skipped on a one-GPU box."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def gpu(dz):
    assert dz.device_count() >= 1, "GPU tests need a CUDA device (there is no CPU fallback)"
    dz.set_device(0)
    return dz


def jittered(n, seed):
    rng = np.random.default_rng(seed)
    return np.concatenate([[0.0], np.cumsum(rng.uniform(0.996, 1.004, n - 1))]) / n


@pytest.mark.parametrize("n", [4096, 300_007, 2048 * 296 + 1, 1_000_003])
def test_one_rank_comm_matches_the_single_gpu_solver(gpu, oracle, n):
    """world = 1: the block is the whole (closed) bar; the collective call must give the bits of DiffussionSolverTimeIndependent"""
    from dzahui_b200.distributed import NativeComm, SplitBarTimeIndependent
    mesh = jittered(n, n)
    p = gpu.DiffussionParams.time_independent().mu(1.0).b(0.5).boundary_conditions(-2.0, 7.0).build()
    comm = NativeComm(None, 0, 1)
    assert comm.world == 1
    for amode in (gpu.ASSEMBLY_FAST, gpu.ASSEMBLY_AUTO):
        opt = gpu.make_options(amode, gpu.SOLVE_PARALLEL, 0)
        bar = SplitBarTimeIndependent(p, mesh, 150, comm, opt)
        u = bar.solve()
        s = gpu.DiffussionSolverTimeIndependent(p, mesh, 150, opt)
        us = s.solve(0.0)
        if amode == gpu.ASSEMBLY_FAST and n % 2048 != 1:
            assert "k_tri_onelaunch" in bar.path()
        lo, di, up, rhs = bar.block.bands()
        slo, sdi, sup, srhs = s.bands()
        assert all(np.array_equal(a, b) for a, b in zip((lo, di, up, rhs), (slo, sdi, sup, srhs)))
        truth = oracle.thomas(lo, di, up, rhs, kind="f128")
        assert rel_l2(u, truth) <= 1e-10
        assert rel_l2(u, us) <= 1e-12
        if "k_tri_onelaunch" in bar.path() and "k_tri_onelaunch" in s.path():
            assert u.tobytes() == us.tobytes()
        # a step = rebuild + solve, with new coordinates in between
        mesh2 = jittered(n, n + 1)
        bar.rebuild(mesh2)
        u2 = bar.solve()
        lo, di, up, rhs = bar.block.bands()
        assert rel_l2(u2, oracle.thomas(lo, di, up, rhs, kind="f128")) <= 1e-10
        bar.rebuild(mesh)
        assert bar.solve().tobytes() == u.tobytes()
        bar.close()
        s.close()
    comm.close()


def test_one_rank_refinement(gpu, oracle):
    from dzahui_b200.distributed import NativeComm, SplitBarTimeIndependent
    n = 1_000_003
    mesh = np.linspace(0.0, 1.0, n)
    p = gpu.DiffussionParams.time_independent().mu(1.0).b(1.0).boundary_conditions(1.0, 15.0).build()
    comm = NativeComm(None, 0, 1)
    bar = SplitBarTimeIndependent(p, mesh, 150, comm, gpu.make_options(gpu.ASSEMBLY_AUTO, gpu.SOLVE_PARALLEL, 1))
    u = bar.solve()
    lo, di, up, rhs = bar.block.bands()
    assert rel_l2(u, oracle.thomas(lo, di, up, rhs, kind="f128")) <= 1e-15
    bar.close()
    comm.close()


def test_comm_argument_checks(gpu):
    from dzahui_b200.distributed import NativeComm
    with pytest.raises(gpu.Invalid):
        NativeComm(None, 0, 2)       # more than one rank needs the id
    with pytest.raises(gpu.Invalid):
        NativeComm(None, 3, 1)
    with pytest.raises(gpu.Invalid):
        NativeComm(b"\0" * 128, 0, 9)  # one box: at most 8 ranks


def _torchrun(nproc, extra_env, *args):
    env = dict(os.environ, **extra_env)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tools", "split_bar_check.py"), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]


@pytest.mark.parametrize("no_peer", [False, True])
def test_two_ranks_against_the_oracle(gpu, no_peer):
    """two GPUs: one bar split in two, both transports (peer mailboxes / ncclAllGather), both assemblies; the script asserts
    parity with the oracle (bands bit-identical to the reference arithmetic, nodal values within the three-way rule)"""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    lines = _torchrun(2, {"DZB_COMM_NO_PEER": "1"} if no_peer else {}, "--nodes", "300007", "1000003", "--steps", "5")
    assert len(lines) == 4
    for l in lines:
        assert l["world"] == 2 and l["relL2_gpu_vs_f128_of_its_bands"] <= 1e-10
        if no_peer:
            assert not l["peer_mailboxes"]


def test_second_device_in_one_process(gpu, oracle):
    """dzb_set_device switches devices inside one process: every device-side one-off (shared-memory attributes, occupancy,
    scratch flags, the copy stream of the pipelined read-back) is kept per device, so the same work on device 1 after
    device 0 gives the same bits.  Needs two GPUs."""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    from dzahui_b200 import synthetic
    p = gpu.DiffussionParams.time_independent().mu(1.0).b(0.5).boundary_conditions(-2.0, 7.0).build()
    mesh = jittered(300_007, 3)
    ids = np.arange(2100)
    meshes = synthetic.ensemble_meshes(ids, 300)
    ic = synthetic.ensemble_initial_conditions(meshes, ids)
    results = []
    try:
        for dev in (0, 1, 0):
            gpu.set_device(dev)
            out = []
            for amode in (gpu.ASSEMBLY_FAST, gpu.ASSEMBLY_AUTO):
                s = gpu.DiffussionSolverTimeIndependent(p, mesh, 150, gpu.make_options(amode, gpu.SOLVE_PARALLEL, 0))
                out.append(s.solve(0.0))
                s.close()
            e = gpu.EnsembleTimeDependent(meshes, 1.0, 1.0, (1.0, 15.0), ic, 150)
            for _ in range(3):
                st = e.solve(1e-8)          # the third call takes the pipelined (copy-stream) read-back
            out.append(st.copy())
            e.close()
            results.append(out)
    finally:
        gpu.set_device(0)
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert a.tobytes() == b.tobytes()


def test_collective_call_checks_its_arguments(gpu):
    """a block whose open ends do not match its rank in the communicator, a solver that is not a block, the library's own
    all-gather on a one-rank communicator"""
    import ctypes as C
    import torch
    from dzahui_b200 import _lib
    from dzahui_b200.distributed import BarBlock, NativeComm
    lib = _lib.load()
    p = gpu.DiffussionParams.time_independent().mu(1.0).b(0.5).boundary_conditions(-2.0, 7.0).build()
    comm = NativeComm(None, 0, 1)
    x = np.linspace(0.0, 1.0, 5001)
    open_left = BarBlock(p, x, True, False, 150, gpu.make_options(gpu.ASSEMBLY_AUTO, gpu.SOLVE_PARALLEL, 0))   # a middle/last block
    ptr = C.c_void_p()
    assert lib.dzb_solver_block_solve_device(open_left._h, comm._h, C.byref(ptr)) == 7      # DZB_INVALID: rank 0 of 1 has no neighbour
    assert b"open ends" in lib.dzb_last_error()
    open_left.close()
    s = gpu.DiffussionSolverTimeIndependent(p, x, 150)
    assert lib.dzb_solver_block_solve_device(s._h, comm._h, C.byref(ptr)) == 7               # not a block solver
    assert lib.dzb_solver_block_rebuild(s._h, None) == 7
    s.close()
    src = torch.arange(6, dtype=torch.float64, device="cuda")
    dst = torch.zeros(6, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    comm.all_gather(src.data_ptr(), dst.data_ptr(), 6)
    gpu.synchronize()
    assert torch.equal(src, dst)
    comm.close()
