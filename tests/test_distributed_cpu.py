"""CPU-only, world_size 2 over gloo: the host side of the multi-GPU paths.

The block kernels need a GPU, so here every rank's block is computed by `OracleBlock` (numpy + the oracle's __float128
Thomas: exact elimination of the block interior), and what is under test is the product's host logic: row / bar
partitioning, halo slicing, the all-gather of interface rows through torch.distributed, the interface solve
(`dzb_interface_solve`, host code of the C-ABI library) and the refinement exchange of edge values."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


class OracleBlock:
    """same interface as dzahui_b200.distributed.BarBlock, exact arithmetic on the CPU (test infrastructure)"""

    def __init__(self, lo, di, up, rhs, has_left, has_right):
        import oracle_lib as o
        self.o, self.lo, self.di, self.up, self.rhs = o, lo, di, up, rhs
        self.hl, self.hr, self.n = has_left, has_right, len(di)
        self.u = np.zeros(self.n)
        self.resid = np.zeros(self.n)

    def _interior(self, f):
        o, m = self.o, self.n
        T = (self.lo[1:m - 1], self.di[1:m - 1], self.up[1:m - 1])
        e0 = np.zeros(m - 2)
        e0[0] = self.lo[1]
        e1 = np.zeros(m - 2)
        e1[-1] = self.up[m - 2]
        return [o.thomas(*T, r, kind="f128") for r in (f[1:m - 1], e0, e1)]

    def block_reduce(self, which_rhs=0):
        f = self.resid if which_rhs else self.rhs
        m = self.n
        y, v, w = self._interior(f)
        a0 = self.lo[0] if self.hl else 0.0
        b0, c0, f0 = self.di[0] - self.up[0] * v[0], -self.up[0] * w[0], f[0] - self.up[0] * y[0]
        c1 = self.up[m - 1] if self.hr else 0.0
        a1, b1, f1 = -self.lo[m - 1] * v[-1], self.di[m - 1] - self.lo[m - 1] * w[-1], f[m - 1] - self.lo[m - 1] * y[-1]
        return np.array([a0, c0, a0 + b0 + c0, f0, a1, c1, a1 + b1 + c1, f1])

    def block_backsolve(self, which_rhs, us, ue, accumulate=False):
        f = self.resid if which_rhs else self.rhs
        y, v, w = self._interior(f)
        x = np.concatenate([[us], y - v * us - w * ue, [ue]])
        self.u = self.u + x if accumulate else x

    def block_edges(self):
        return np.array([self.u[0], self.u[-1]])

    def block_residual(self, u_left, u_right):
        L = np.longdouble
        ext = np.concatenate([[u_left], self.u, [u_right]]).astype(L)
        lo = self.lo.astype(L).copy()
        up = self.up.astype(L).copy()
        if not self.hl:
            lo[0] = 0
        if not self.hr:
            up[-1] = 0
        self.resid = (self.rhs.astype(L) - lo * ext[:-2] - self.di.astype(L) * ext[1:-1] - up * ext[2:]).astype(np.float64)

    def block_read(self):
        return self.u.copy()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, result_dir):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    import oracle_lib as o
    from dzahui_b200 import distributed as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        comm = D.TorchComm()
        assert (comm.rank, comm.world) == (rank, world)
        # ---- ensembles: bars shard with no overlap and no gap
        first, last = D.shard_bars(65536 + 3, rank, world)
        spans = comm.all_gather(np.array([[first, last]], dtype=np.float64))
        assert spans[0][0] == 0 and spans[-1][1] == 65536 + 3
        assert all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
        # ---- one bar split in row blocks, 2 blocks per process -> 4 blocks in total
        rng = np.random.default_rng(2)
        n = 403
        mesh = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, n - 1))]) / n
        lo, di, up, rhs = o.ti_assemble(mesh, 1.0, 1.0, (1.0, 15.0), 150)
        want = o.thomas(lo, di, up, rhs, kind="f128")
        for per in (1, 2):
            parts = D.partition_rows(n, world * per)
            blocks = []
            for j in range(per):
                a, b = parts[rank * per + j]
                nodes, hl, hr = D.block_nodes(mesh, a, b)
                assert len(nodes) == (b - a) + hl + hr and nodes[int(hl)] == mesh[a]
                blocks.append(OracleBlock(lo[a:b], di[a:b], up[a:b], rhs[a:b], hl, hr))
            for refine in (0, 1):
                got = np.concatenate(D.spike_solve(blocks, comm, refine_steps=refine))
                a, b = parts[rank * per][0], parts[rank * per + per - 1][1]
                err = np.linalg.norm(got - want[a:b]) / np.linalg.norm(want[a:b])
                assert err <= 1e-12, (per, refine, err)
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_partition_helpers():
    from dzahui_b200 import distributed as D
    assert [D.shard_bars(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert D.partition_rows(10_000_000, 8)[-1] == (8_750_000, 10_000_000)
    with pytest.raises(ValueError):
        D.partition_rows(5, 4)
    mesh = np.arange(10.0)
    nodes, hl, hr = D.block_nodes(mesh, 0, 4)
    assert (list(nodes), hl, hr) == ([0, 1, 2, 3, 4], False, True)
    nodes, hl, hr = D.block_nodes(mesh, 4, 10)
    assert (list(nodes), hl, hr) == ([3, 4, 5, 6, 7, 8, 9], True, False)


def test_interface_solve_is_host_code(dz):
    """dzb_interface_solve runs without a GPU: two 1-unknown-wide 'blocks' of a 4-row system"""
    from dzahui_b200 import distributed as D
    # u = (1, 2, 3, 4); rows (a, c, s, f) with diagonal = s - a - c
    A = np.array([[4., -1, 0, 0], [-1, 4, -1, 0], [0, -1, 4, -1], [0, 0, -1, 4]])
    u = np.array([1., 2, 3, 4])
    f = A @ u
    flat = []
    for i in range(4):
        a = A[i, i - 1] if i > 0 else 0.0
        c = A[i, i + 1] if i < 3 else 0.0
        flat += [a, c, a + A[i, i] + c, f[i]]
    ends = D.interface_solve(np.array(flat).reshape(2, 8))
    assert np.allclose(ends.ravel(), u, rtol=0, atol=1e-14)


def test_spike_over_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
