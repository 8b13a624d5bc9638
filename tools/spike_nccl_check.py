#!/usr/bin/env python3
"""Run under torchrun on N GPUs: one regular bar split SPIKE-style across the ranks (NCCL all-gather of the interface
rows), checked on rank 0 against the CPU oracle (f64 Thomas and __float128 Thomas of the same system), then timed.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/spike_nccl_check.py --nodes 1000003"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1_000_003)
    ap.add_argument("--refine", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import dzahui_b200 as dz
    from dzahui_b200 import distributed as D, synthetic
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dz.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = D.TorchComm()
    n = args.nodes
    mesh = synthetic.regular_bar(n)
    p = dz.DiffussionParams.time_independent().mu(1.0).b(1.0).boundary_conditions(1.0, 15.0).build()
    bar = D.DistributedBarTimeIndependent(p, mesh, 150, comm=comm, refine_steps=args.refine,
                                          options=dz.make_options(dz.ASSEMBLY_AUTO, dz.SOLVE_PARALLEL))
    u = bar.solve()
    # gather the pieces on every rank (check only)
    sizes = [b - a for a, b in bar.partition]
    pad = np.zeros(max(sizes))
    pad[:len(u)] = u
    allu = comm.all_gather(pad[None, :])
    full = np.concatenate([allu[r][:sizes[r]] for r in range(world)])
    out = {"n": n, "world": world, "refine": args.refine}
    if rank == 0 and not args.no_check:
        import oracle_lib as o
        thr = max(1, len(os.sched_getaffinity(0)))
        lo, di, up, rhs = o.ti_assemble(mesh, 1.0, 1.0, (1.0, 15.0), 150, mode=0, threads=thr)
        ref = o.thomas(lo, di, up, rhs)
        truth = o.thomas(lo, di, up, rhs, kind="f128")
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        out.update(relL2_gpu_truth=rel(full, truth), relL2_ref_truth=rel(ref, truth), relL2_gpu_ref=rel(full, ref))
    # timing: solve only (bands resident).  (a) host-mediated exchange, (b) device-resident pipeline on one stream
    def timed(fn):
        dz.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def host_path():
        rows = comm.all_gather(np.stack([b.block_reduce(0) for b in bar.blocks]))
        ends = D.interface_solve(rows)
        bar.blocks[0].block_backsolve(0, ends[rank][0], ends[rank][1], False)

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dz.set_stream(stream.cuda_stream)
    bar.solve_device()
    u_dev = bar.read()
    sizes_ok = len(u_dev) == len(u)
    err_dev_vs_host = float(np.linalg.norm(u_dev - u) / np.linalg.norm(u))
    t_host = timed(host_path)
    for _ in range(3):
        bar.solve_device()
    t_dev = timed(bar.solve_device)
    out.update(ms_per_solve_host_exchange=t_host * 1e3, ms_per_solve=t_dev * 1e3, dof_per_s=n / t_dev,
               relL2_device_vs_host_path=err_dev_vs_host, refine_in_timing=args.refine)
    assert sizes_ok and err_dev_vs_host <= 1e-12, err_dev_vs_host
    if rank == 0:
        print(json.dumps(out), flush=True)
        if "relL2_gpu_truth" in out:
            assert out["relL2_gpu_truth"] <= max(1e-10, out["relL2_ref_truth"]), out
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
