#!/usr/bin/env python3
"""How fast can the HOST take results?  Every rank copies `--mb` MB from its GPU into pinned host memory `--reps` times with a
bare cudaMemcpyAsync (torch), all ranks at once — no kernels, no library.  The aggregate GB/s is the ceiling of bench.py's end
to end leg (`e2e`: every step ships every bar's state, 0.54 GB per rank, to the host), whatever the kernels do.
    python tools/d2h_probe.py                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_probe.py
Prints one JSON line (rank 0): per-rank and aggregate device-to-host GB/s, and the same for f32-sized (half) transfers."""
import argparse
import json
import os
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=536.870912)   # the C5 state of one rank: 65536 x 1024 f64
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    try:  # same placement as bench.py: the CPUs next to this GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world, "reps": args.reps}
    for tag, nbytes in (("f64_state", int(args.mb * 1e6)), ("f32_state", int(args.mb * 1e6) // 2)):
        dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            host.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t.item())
        out[tag] = {"bytes_per_copy": nbytes, "ms_per_copy_slowest_rank": worst / args.reps * 1e3,
                    "GBps_per_rank": nbytes * args.reps / worst / 1e9, "GBps_aggregate": world * nbytes * args.reps / worst / 1e9,
                    "this_rank_GBps": nbytes * args.reps / sec / 1e9}
    try:
        out["cpu_affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
