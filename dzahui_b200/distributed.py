"""Multi-GPU paths of the hot path, one process per GPU.

* Ensembles shard by bars: `shard_bars` gives a rank its contiguous slice; nothing is exchanged while stepping.
* One long bar (BASELINE config 4) is split SPIKE-style into contiguous row blocks, one per rank.  Each rank
  eliminates everything inside its block on its own GPU (the same chunk-reduction kernels as the single-GPU solver,
  with the block's end couplings kept), ONE all-gather moves 8 doubles per rank (the two interface rows), every rank
  solves the 2P-row interface system redundantly on the host, and back-substitutes locally.  Optional refinement sweeps
  add one all-gather of the block edge values (the residual's halo) and one more of interface rows per sweep.

Two drivers.  `SplitBarTimeIndependent` + `NativeComm` is the production one: the communicator lives inside the C library
(dzb_comm_*: NCCL bound at run time plus NVLink peer mailboxes), a solve is ONE collective C call per rank
(dzb_solver_block_solve_device) and, with closed-form assembly, one kernel launch per rank.  The older step-by-step driver
(`spike_solve`, `DeviceSpike`) moves the interface rows through `torch.distributed` (NCCL on the GPU box, gloo in the CPU
tests); `LocalComm` runs all blocks in one process — what a single-GPU test uses to exercise the block kernels."""
import ctypes as C

import numpy as np

from . import _lib
from .errors import check
from .mesh import Mesh


# ---------------------------------------------------------------------------------------------- partitioning
def shard_bars(n_bars, rank, world):
    """contiguous slice [first, last) of an ensemble's bars owned by `rank`; sizes differ by at most one"""
    base, rem = divmod(int(n_bars), int(world))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def partition_rows(n, world):
    """row blocks [(first, last), ...] of one bar; every block keeps at least 2 rows"""
    world = int(world)
    if n < 2 * world:
        raise ValueError("a bar of %d nodes cannot be split into %d blocks of at least 2 rows" % (n, world))
    return [shard_bars(n, r, world) for r in range(world)]


def block_nodes(mesh, first, last):
    """the node coordinates block [first, last) needs: its own plus one halo node per existing neighbour"""
    n = len(mesh)
    has_left, has_right = first > 0, last < n
    return np.ascontiguousarray(mesh[first - (1 if has_left else 0): last + (1 if has_right else 0)]), has_left, has_right


# ---------------------------------------------------------------------------------------------- communicators
class LocalComm:
    """all blocks live in this process (single-GPU emulation of P ranks, run one after another)"""
    rank, world = 0, 1

    def all_gather(self, local):            # local: [k_local, width] -> [k_total, width]
        return np.asarray(local, dtype=np.float64)


class TorchComm:
    """torch.distributed process group: NCCL (CUDA staging tensors) or gloo (CPU tensors)"""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist, self._group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        backend = dist.get_backend(group)
        self._device = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                                          if backend == "nccl" else torch.device("cpu"))

    def all_gather(self, local):
        torch = self._torch
        loc = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(self._device)
        out = torch.empty((self.world,) + tuple(loc.shape), dtype=torch.float64, device=self._device)
        self._dist.all_gather_into_tensor(out.view(-1), loc.view(-1), group=self._group)
        return out.cpu().numpy().reshape(self.world * loc.shape[0], *loc.shape[1:])


# ---------------------------------------------------------------------------------------------- one block on one GPU
class BarBlock:
    """DiffussionSolverTimeIndependent for the rows of one block (dzb_diffusion_ti_create_block)."""

    def __init__(self, params, nodes_with_halo, has_left, has_right, gauss_step, options=None):
        self._lib = _lib.load()
        self._h = None
        self._mesh = Mesh.from_nodes(nodes_with_halo)
        h = C.c_void_p()
        check(self._lib.dzb_diffusion_ti_create_block(self._mesh._h, int(bool(has_left)), int(bool(has_right)), params.mu, params.b,
                                                      params.boundary_conditions[0], params.boundary_conditions[1],
                                                      int(gauss_step), C.byref(options) if options is not None else None,
                                                      C.byref(h)), self._lib)
        self._h = h
        self.n = len(nodes_with_halo) - int(bool(has_left)) - int(bool(has_right))

    def block_reduce(self, which_rhs=0):
        top = np.empty(8)
        check(self._lib.dzb_solver_block_reduce(self._h, int(which_rhs), top.ctypes.data_as(_lib.c_double_p)), self._lib)
        return top

    def block_backsolve(self, which_rhs, us, ue, accumulate=False):
        check(self._lib.dzb_solver_block_backsolve(self._h, int(which_rhs), float(us), float(ue), int(bool(accumulate))), self._lib)

    def block_edges(self):
        e = np.empty(2)
        check(self._lib.dzb_solver_block_edges(self._h, e.ctypes.data_as(_lib.c_double_p)), self._lib)
        return e

    def block_residual(self, u_left, u_right):
        check(self._lib.dzb_solver_block_residual(self._h, float(u_left), float(u_right)), self._lib)

    # -- the same steps on device buffers (raw pointers, e.g. torch tensors' data_ptr()), no host synchronisation
    def block_reduce_device(self, which_rhs, d_top8):
        check(self._lib.dzb_solver_block_reduce_device(self._h, int(which_rhs), C.c_void_p(int(d_top8))), self._lib)

    def block_backsolve_device(self, which_rhs, d_ends2, accumulate=False):
        check(self._lib.dzb_solver_block_backsolve_device(self._h, int(which_rhs), C.c_void_p(int(d_ends2)), int(bool(accumulate))),
              self._lib)

    def block_edges_device(self, d_edges2):
        check(self._lib.dzb_solver_block_edges_device(self._h, C.c_void_p(int(d_edges2))), self._lib)

    def block_residual_device(self, d_left, d_right):
        check(self._lib.dzb_solver_block_residual_device(self._h, C.c_void_p(int(d_left)) if d_left else None,
                                                         C.c_void_p(int(d_right)) if d_right else None), self._lib)

    def device_solution_ptr(self):
        p = C.c_void_p()
        check(self._lib.dzb_solver_block_device_solution(self._h, C.byref(p)), self._lib)
        return p.value

    def block_read(self):
        out = np.empty(self.n)
        check(self._lib.dzb_solver_block_read(self._h, out.ctypes.data_as(_lib.c_double_p), self.n), self._lib)
        return out

    def block_read_into(self, host_ptr):
        """the block's nodal values into a caller's host buffer of self.n doubles (pinned memory: PCIe speed)"""
        check(self._lib.dzb_solver_block_read(self._h, C.cast(C.c_void_p(int(host_ptr)), _lib.c_double_p), self.n), self._lib)

    def bands(self):
        lo, di, up, rhs = (np.empty(self.n) for _ in range(4))
        check(self._lib.dzb_solver_bands(self._h, 0, lo.ctypes.data_as(_lib.c_double_p), di.ctypes.data_as(_lib.c_double_p),
                                         up.ctypes.data_as(_lib.c_double_p), rhs.ctypes.data_as(_lib.c_double_p), self.n), self._lib)
        return lo, di, up, rhs

    def close(self):
        if self._h:
            self._lib.dzb_solver_destroy(self._h)
            self._h = None
        if self._mesh is not None:
            self._mesh.close()
            self._mesh = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def interface_solve(rows):
    """rows [P, 8] (block r's s-row and e-row as (a, c, s, f)) -> ends [P, 2]; host side, long double (C ABI)"""
    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 8)
    ends = np.empty((rows.shape[0], 2))
    check(lib.dzb_interface_solve(rows.ctypes.data_as(_lib.c_double_p), rows.shape[0], ends.ctypes.data_as(_lib.c_double_p)), lib)
    return ends


# ---------------------------------------------------------------------------------------------- the SPIKE driver
def spike_solve(blocks, comm, refine_steps=0, solve_interface=interface_solve):
    """Solves the bar whose consecutive row blocks are `blocks` on this process (one in production) and the other ranks
    of `comm`.  Returns the list of this process' block solutions.  Exactly one all-gather of 8 doubles per block per
    solve, plus one of 2 doubles per refinement sweep."""
    k = len(blocks)
    first_global = comm.rank * k  # every process owns the same number of consecutive blocks

    def one_solve(which_rhs, accumulate):
        rows = comm.all_gather(np.stack([b.block_reduce(which_rhs) for b in blocks]))
        ends = solve_interface(rows)
        for j, b in enumerate(blocks):
            us, ue = ends[first_global + j]
            b.block_backsolve(which_rhs, us, ue, accumulate)
        return rows.shape[0]

    total = one_solve(0, False)
    for _ in range(int(refine_steps)):
        edges = comm.all_gather(np.stack([b.block_edges() for b in blocks]))
        for j, b in enumerate(blocks):
            g = first_global + j
            b.block_residual(edges[g - 1][1] if g > 0 else 0.0, edges[g + 1][0] if g + 1 < total else 0.0)
        one_solve(1, True)
    return [b.block_read() for b in blocks]


class DeviceSpike:
    """The SPIKE solve with every operand in device memory (torch CUDA tensors as buffers): block reductions, ONE
    `all_gather_into_tensor` of 8 doubles per block over NCCL, the interface solve as a one-thread kernel run redundantly
    on every rank, back-substitution — all queued on the current stream, nothing waits for the host."""

    def __init__(self, blocks, group=None, world=None, rank=None):
        import torch
        self._torch, self.blocks, self.group = torch, blocks, group
        if world is None:
            import torch.distributed as dist
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world, self.rank, self.k = world, rank, len(blocks)
        dev = torch.device("cuda", torch.cuda.current_device())
        P = world * self.k
        self.top = torch.zeros((self.k, 8), dtype=torch.float64, device=dev)
        self.rows = torch.zeros((P, 8), dtype=torch.float64, device=dev)
        self.ends = torch.zeros((P, 2), dtype=torch.float64, device=dev)
        self.work = torch.zeros((P, 4), dtype=torch.float64, device=dev)
        self.edges = torch.zeros((self.k, 2), dtype=torch.float64, device=dev)
        self.alledges = torch.zeros((P, 2), dtype=torch.float64, device=dev)
        self._lib = _lib.load()

    def _gather(self, dst, src):
        if self.world == 1:
            dst.copy_(src)
        else:
            import torch.distributed as dist
            dist.all_gather_into_tensor(dst.view(-1), src.view(-1), group=self.group)

    def _one(self, which_rhs, accumulate):
        for j, b in enumerate(self.blocks):
            b.block_reduce_device(which_rhs, self.top[j].data_ptr())
        self._gather(self.rows, self.top)
        check(self._lib.dzb_interface_solve_device(C.c_void_p(self.rows.data_ptr()), self.rows.shape[0],
                                                   C.c_void_p(self.ends.data_ptr()), C.c_void_p(self.work.data_ptr())), self._lib)
        for j, b in enumerate(self.blocks):
            b.block_backsolve_device(which_rhs, self.ends[self.rank * self.k + j].data_ptr(), accumulate)

    def solve(self, refine_steps=0):
        """leaves every block's solution in its device buffer (BarBlock.device_solution_ptr / block_read)"""
        self._one(0, False)
        P = self.world * self.k
        for _ in range(int(refine_steps)):
            for j, b in enumerate(self.blocks):
                b.block_edges_device(self.edges[j].data_ptr())
            self._gather(self.alledges, self.edges)
            for j, b in enumerate(self.blocks):
                g = self.rank * self.k + j
                left = self.alledges[g - 1, 1:].data_ptr() if g > 0 else 0
                right = self.alledges[g + 1, 0:].data_ptr() if g + 1 < P else 0
                b.block_residual_device(left, right)
            self._one(1, True)


class DistributedBarTimeIndependent:
    """`DiffussionSolverTimeIndependent` (time_independent.rs:46-53) for one bar split across the ranks of `comm`.
    `mesh` is the whole bar's node array (every rank slices its own block + halo from it) or, with `local=True`,
    already this rank's `block_nodes`."""

    def __init__(self, params, mesh, gauss_step, comm=None, options=None, refine_steps=0, blocks_per_process=1):
        self.comm = comm if comm is not None else LocalComm()
        mesh = np.asarray(mesh, dtype=np.float64)
        total_blocks = self.comm.world * blocks_per_process
        self.partition = partition_rows(len(mesh), total_blocks)
        self.refine_steps = int(refine_steps)
        self.blocks = []
        for j in range(blocks_per_process):
            first, last = self.partition[self.comm.rank * blocks_per_process + j]
            nodes, hl, hr = block_nodes(mesh, first, last)
            self.blocks.append(BarBlock(params, nodes, hl, hr, gauss_step, options))
        self.rows = (self.partition[self.comm.rank * blocks_per_process][0],
                     self.partition[(self.comm.rank + 1) * blocks_per_process - 1][1])

    def solve(self, _time_step=0.0):
        """nodal values of this process' rows [self.rows[0], self.rows[1])"""
        return np.concatenate(spike_solve(self.blocks, self.comm, self.refine_steps))

    def solve_device(self, group=None):
        """the same solve with device-resident operands (DeviceSpike); results stay on the device until block_read()"""
        if getattr(self, "_dev", None) is None:
            self._dev = DeviceSpike(self.blocks, group, self.comm.world, self.comm.rank)
        self._dev.solve(self.refine_steps)

    def read(self):
        return np.concatenate([b.block_read() for b in self.blocks])

    def close(self):
        for b in self.blocks:
            b.close()
        self.blocks = []


# ---------------------------------------------------------------------------------------------- native communicator
COMM_ID_BYTES = 128


class NativeComm:
    """dzb_comm: the library's own communicator (include/dzahui_b200.h).  `NativeComm.from_torch()` takes rank / world from
    the default torch.distributed group and ships rank 0's 128-byte id through it — the one thing the host has to do;
    `NativeComm(id_bytes, rank, world)` is the same for a host that moved the id some other way (a Rust host: MPI, a
    file).  world == 1 needs no id and no NCCL."""

    def __init__(self, unique_id, rank, world):
        self._lib = _lib.load()
        self._h = None
        h = C.c_void_p()
        check(self._lib.dzb_comm_init(unique_id, int(rank), int(world), C.byref(h)), self._lib)
        self._h = h
        self.rank, self.world = int(rank), int(world)
        pm = C.c_int()
        check(self._lib.dzb_comm_info(h, None, None, C.byref(pm)), self._lib)
        self.peer_mailboxes = bool(pm.value)

    @staticmethod
    def unique_id():
        lib = _lib.load()
        buf = C.create_string_buffer(COMM_ID_BYTES)
        check(lib.dzb_comm_unique_id(buf), lib)
        return buf.raw

    @classmethod
    def from_torch(cls, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            return cls(None, 0, 1)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(box[0], rank, world)

    def all_gather(self, d_send, d_recv, count):
        """ncclAllGather of `count` doubles per rank between device pointers, on the library's stream"""
        check(self._lib.dzb_comm_all_gather(self._h, C.c_void_p(int(d_send)), C.c_void_p(int(d_recv)), int(count)), self._lib)

    def close(self):
        if self._h:
            self._lib.dzb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SplitBarTimeIndependent:
    """`DiffussionSolverTimeIndependent` (time_independent.rs:46-53) for ONE bar split into row blocks, one per rank of a
    `NativeComm`.  `mesh` is the whole bar's node array (every rank slices its block + halo nodes from it).
    `rebuild()` + `solve_device()` is one assemble + solve step; nothing in them waits for the host."""

    def __init__(self, params, mesh, gauss_step, comm, options=None):
        self.comm = comm
        mesh = np.asarray(mesh, dtype=np.float64)
        self.n_total = len(mesh)
        self.partition = partition_rows(self.n_total, comm.world)
        self.rows = self.partition[comm.rank]
        nodes, hl, hr = block_nodes(mesh, *self.rows)
        self.block = BarBlock(params, nodes, hl, hr, gauss_step, options)
        self._lib = self.block._lib

    def rebuild(self, nodes_with_halo=None):
        """XSolver::new again into the block's buffers; `nodes_with_halo`: new coordinates of the block (own + halo nodes)"""
        b = self.block
        if nodes_with_halo is not None:
            b._mesh.set_nodes(nodes_with_halo)
        check(self._lib.dzb_solver_block_rebuild(b._h, b._mesh._h), self._lib)

    def solve_device(self):
        """collective: every rank calls it; returns the device pointer of this block's nodal values"""
        p = C.c_void_p()
        check(self._lib.dzb_solver_block_solve_device(self.block._h, self.comm._h, C.byref(p)), self._lib)
        return p.value

    def status(self):
        """synchronises; raises if a collective solve since the last call went wrong on this rank"""
        check(self._lib.dzb_solver_block_status(self.block._h, self.comm._h), self._lib)

    def solve(self, _time_step=0.0):
        """nodal values of this rank's rows [self.rows[0], self.rows[1])"""
        self.solve_device()
        self.status()
        return self.block.block_read()

    def path(self):
        buf = C.create_string_buffer(160)
        check(self._lib.dzb_solver_path(self.block._h, buf, 160), self._lib)
        return buf.value.decode()

    def close(self):
        if self.block is not None:
            self.block.close()
            self.block = None
