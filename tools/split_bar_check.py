#!/usr/bin/env python3
"""Run under torchrun on N GPUs: ONE regular bar split into row blocks, one per rank, through the library's own communicator
(dzb_comm_*: NCCL bound at run time + NVLink peer mailboxes) and the collective `dzb_solver_block_solve_device`.  Every
size / assembly is checked on rank 0 against the CPU oracle (the reference's quadrature-loop bands, f64 Thomas and
__float128 Thomas), then assemble + solve and the solve alone are timed on the device (CUDA events, max over ranks).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/split_bar_check.py --nodes 1000003
DZB_COMM_NO_PEER=1 forces the ncclAllGather transport."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, nargs="+", default=[1_000_003])
    ap.add_argument("--assembly", nargs="+", default=["fast", "auto"])
    ap.add_argument("--refine", type=int, default=0)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--jitter", action="store_true", help="irregular (0.4 %% jitter) mesh instead of the regular bar")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import dzahui_b200 as dz
    from dzahui_b200 import distributed as D, synthetic
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dz.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dz.set_stream(stream.cuda_stream)
    comm = D.NativeComm.from_torch()
    mu, b, bc, G = 1.0, 1.0, (1.0, 15.0), 150
    p = dz.DiffussionParams.time_independent().mu(mu).b(b).boundary_conditions(*bc).build()
    rel = lambda a, c: float(np.linalg.norm(a - c) / np.linalg.norm(c))

    def timed(fn):
        if world > 1:
            dist.barrier()   # rank 0 may come out of a long CPU check: the mailboxes' time-out is not meant to cover that
        for _ in range(5):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            fn()
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    results = []
    for n in args.nodes:
        if args.jitter:
            rng = np.random.default_rng(n)
            mesh = np.concatenate([[0.0], np.cumsum(rng.uniform(0.996, 1.004, n - 1))]) / n
        else:
            mesh = synthetic.regular_bar(n)
        truth = {}
        for mode in args.assembly:
            amode = {"auto": dz.ASSEMBLY_AUTO, "faithful": dz.ASSEMBLY_FAITHFUL, "fast": dz.ASSEMBLY_FAST}[mode]
            bar = D.SplitBarTimeIndependent(p, mesh, G, comm, dz.make_options(amode, dz.SOLVE_PARALLEL, args.refine))
            u = bar.solve()
            out = {"n": n, "world": world, "assembly": mode, "refine": args.refine, "peer_mailboxes": comm.peer_mailboxes,
                   "path": bar.path()}
            if not args.no_check:
                bands = bar.block.bands()
                if world > 1:
                    parts = [None] * world if rank == 0 else None
                    dist.gather_object((u, bands), parts, dst=0)
                else:
                    parts = [(u, bands)]
                if rank == 0:
                    import oracle_lib as o
                    full = np.concatenate([q[0] for q in parts])
                    lo, di, up, rhs = (np.concatenate([q[1][k] for q in parts]) for k in range(4))
                    own = o.thomas(lo, di, up, rhs, kind="f128")     # exact solution of the system the GPUs assembled
                    out["relL2_gpu_vs_f128_of_its_bands"] = rel(full, own)
                    if mode != "fast":
                        if not truth:
                            thr = max(1, len(os.sched_getaffinity(0)))
                            rb = o.ti_assemble(mesh, mu, b, bc, G, mode=0, threads=thr)
                            truth["bands_equal"] = all(np.array_equal(x, y) for x, y in zip(rb, (lo, di, up, rhs)))
                            truth["ref"] = o.thomas(*rb)
                            truth["f128"] = o.thomas(*rb, kind="f128")
                        out.update(bands_bit_identical_to_reference=truth["bands_equal"], relL2_gpu_vs_reference_f64=rel(full, truth["ref"]),
                                   relL2_reference_f64_vs_f128=rel(truth["ref"], truth["f128"]), relL2_gpu_vs_f128=rel(full, truth["f128"]))
                    assert full[0] == bc[0] and full[-1] == bc[1], (full[0], full[-1])
                    assert out["relL2_gpu_vs_f128_of_its_bands"] <= 1e-10, out
                    if mode != "fast":
                        assert truth["bands_equal"], "block bands differ from the reference arithmetic"
                        assert out["relL2_gpu_vs_f128"] <= max(1e-10, out["relL2_reference_f64_vs_f128"]), out

            def step():
                bar.rebuild()
                bar.solve_device()
            out["ms_assemble_plus_solve"] = timed(step)
            out["ms_solve"] = timed(bar.solve_device)
            out["ms_assemble"] = timed(bar.rebuild)
            bar.status()
            out["dof_per_s"] = n / (out["ms_assemble_plus_solve"] * 1e-3)
            u2 = bar.solve()
            assert u2.tobytes() == u.tobytes(), "the solve is not repeatable bit for bit"
            bar.close()
            results.append(out)
            if rank == 0:
                print(json.dumps(out), flush=True)
    comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
