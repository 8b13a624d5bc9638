#!/usr/bin/env python3
"""bench.py — FEM DOF·steps/s of the Dzahui 1D hot path on B200 (BASELINE.json metric), one process per GPU.

Default workload (N = 1): BASELINE config 5 — an ensemble of 65536 independent 1024-node bars, time-dependent
diffusion, one `solve(dt)` on every bar per step (diffusion_solver/time_dependent.rs:257-280).  N > 1: every rank owns
its own 65536-bar slice (global bar ids rank*65536 ...), no data-path communication -> "scaling": "weak".
`--workload c4` times BASELINE config 4 instead (one 10^7-node bar: static diffusion assemble + solve per step).

  value     device-resident throughput: K steps timed with CUDA events on the launching stream, max over ranks
  e2e       the same steps through the reference-facing call `solve(dt) -> Vec<f64>` (dzb_ensemble_solve /
            dzb_solver_solve) with a pinned HOST result buffer: the D2H copy of every step's result is inside the region
  roofline  dominant kernel, algorithmic bytes / CUDA-event time per launch, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/, port of the reference algorithm) on a bounded sample, rank 0, N = 1
`--impl reference` runs only that CPU port with every host thread on the same config/metric."""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "FEM DOF-steps/sec (assemble+tridiag solve)"
UNIT = "DOF-steps/s"
MU, BADV, BC, G = 1.0, 1.0, (1.0, 15.0), 150
C5_N, C5_DT = 1024, 1e-8
ALG_BYTES_TD_STEP = 64       # u in/out 16 B + 3 M bands + 3 S bands 48 B per DOF-step; k_td_step_lean moves 40
ALG_BYTES_STATIC = 80        # assembly 40 B/DOF + solve 40 B/DOF


def c5_workload(bars):
    return (f"C5: ensemble of {bars} x {C5_N}-node bars per GPU, time-dependent diffusion, one solve(dt) per bar per step, "
            f"dt={C5_DT}, G={G}, mu=b=1, bc=(1,15)")


def c4_workload(nodes):
    return f"C4: one regular bar of {nodes} nodes per GPU, static diffusion assemble + solve per step, G={G}, mu=b=1, bc=(1,15)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dzahui_b200", choices=["dzahui_b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c4"])
    ap.add_argument("--bars", type=int, default=65536, help="bars per GPU (c5)")
    ap.add_argument("--nodes", type=int, default=10_000_000, help="nodes of the single bar (c4)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 100)")
    ap.add_argument("--cpu-sample-bars", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stored-operators", action="store_true",
                    help="c5: step from pre-scaled operator arrays (k_td_step_fast, 56 B/DOF-step) instead of k_td_step_lean")
    ap.add_argument("--assembly", default="auto", choices=["auto", "faithful", "fast"],
                    help="c4: auto = the reference's arithmetic (bit-identical bands); fast = closed form (HBM-bound)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- clocks during the run
def ramp_clocks(step, index, budget_s=0.4):
    """More untimed warm-up: a few warm-up steps last ~2 ms, less than the SM clock needs to leave its idle state, and a
    timed region that starts during the ramp reports clocks below max with no throttle reason.  Steps until NVML shows the
    SM clock within 3 % of its maximum, or `budget_s` has passed."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        top = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < budget_s:
            for _ in range(10):
                step()
            torch.cuda.synchronize()
            if pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM) >= 0.97 * top:
                break
    except Exception:
        pass


class ClockSampler(threading.Thread):
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        while not self.stop_flag:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# ---------------------------------------------------------------------------------------------- CPU legs (oracle port)
class CpuTdSample:
    """the oracle's DiffussionSolverTimeDependent::solve on `sample_bars` bars of the C5 ensemble"""

    def __init__(self, sample_bars, threads, rank_offset=0):
        import numpy as np
        import oracle_lib as o
        from dzahui_b200 import synthetic
        ids = np.arange(sample_bars) + rank_offset
        meshes = synthetic.ensemble_meshes(ids, C5_N)
        ic = synthetic.ensemble_initial_conditions(meshes, ids)
        self.o, self.threads, self.bars = o, threads, sample_bars
        self.bands = o.ensemble_td_assemble(meshes, MU, BADV, G, threads=threads)
        self.st = np.stack([o.td_initial_state(ic[k], C5_N, BC) for k in range(sample_bars)])
        self.run(1)  # warm

    def run(self, steps):
        """-> (DOF-steps/s, seconds)"""
        t0 = time.perf_counter()
        self.st = self.o.ensemble_td_run(self.bands, self.st, C5_DT, BC, steps, threads=self.threads)
        dt = time.perf_counter() - t0
        return self.bars * C5_N * steps / dt, dt


def cpu_td_sample(sample_bars, steps, threads, rank_offset=0):
    return CpuTdSample(sample_bars, threads, rank_offset).run(steps)


def cpu_static_sample(n, threads):
    """the oracle's assemble (reference quadrature loop, table hoisted) + solve_by_thomas on an n-node regular bar"""
    import oracle_lib as o
    from dzahui_b200 import synthetic
    mesh = synthetic.regular_bar(n)
    t0 = time.perf_counter()
    lo, di, up, rhs = o.ti_assemble(mesh, MU, BADV, BC, G, mode=0, threads=threads)
    o.thomas(lo, di, up, rhs)
    dt = time.perf_counter() - t0
    return n / dt, dt


def host_threads():
    """cores this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what the reference arm wants)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args, rank):
    """the reference's own CPU algorithm (oracle port; the Rust crate cannot be built in this image) with all host threads"""
    if rank != 0:
        return
    threads = host_threads()
    vals, secs = [], []
    if args.workload == "c5":
        sample = f"{args.cpu_sample_bars} of {args.bars} bars x {C5_N} nodes, 10 solve(dt) per bench step"
        leg = CpuTdSample(args.cpu_sample_bars, threads)
        for k in range(args.warmup + args.steps):
            v, sec = leg.run(10)
            if k >= args.warmup:
                vals.append(v)
                secs.append(sec)
        cfg = {"workload": c5_workload(args.bars)}
    else:
        nn = min(args.nodes, 1_000_000)
        sample = f"{nn}-node regular bar (of {args.nodes}), one assemble + Thomas per bench step"
        for k in range(args.warmup + args.steps):
            v, sec = cpu_static_sample(nn, threads)
            if k >= args.warmup:
                vals.append(v)
                secs.append(sec)
        cfg = {"workload": c4_workload(args.nodes)}
    value = statistics.mean(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": statistics.mean(secs) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------- GPU arm
_JSON_OUT = None


def protect_stdout():
    """stdout carries exactly one JSON line.  Libraries write to fd 1 behind Python's back (NCCL prints its version banner
    there whenever NCCL_DEBUG is VERSION or louder, whatever NCCL_DEBUG_FILE says — the GPU boxes export NCCL_DEBUG=VERSION):
    keep a private duplicate of the real stdout for the line and point fd 1 at stderr for everybody else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    protect_stdout()
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import dzahui_b200 as dz
    from dzahui_b200 import synthetic

    if not torch.cuda.is_available() or dz.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: dzahui_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dz.set_device(local_rank)
    try:  # run (and first-touch the pinned result buffer) on the CPUs next to this GPU: the e2e leg is a 0.5 GB D2H per step
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dz.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = peaks()
    extra = {}
    if args.workload == "c5":
        B, n = args.bars, C5_N
        ids = np.arange(B, dtype=np.int64) + rank * B
        meshes = synthetic.ensemble_meshes(ids, n)
        ic = synthetic.ensemble_initial_conditions(meshes, ids)
        t0 = time.perf_counter()
        solver = dz.EnsembleTimeDependent(meshes, MU, BADV, BC, ic, G,   # H2D + assembly + factorisation
                                          dz.make_options(stored_operators=True) if args.stored_operators else None)
        dz.synchronize()
        extra["create_s"] = time.perf_counter() - t0
        dof = B * n
        step_dev = lambda: solver.step(C5_DT, 1)
        host_out = torch.empty(dof, dtype=torch.float64, pin_memory=True)
        step_e2e = lambda: solver.solve_into(C5_DT, host_out.data_ptr())
        alg_bytes = ALG_BYTES_TD_STEP * dof
        kernel = "k_td_step_fast<32,4>" if args.stored_operators else "k_td_step_lean<8,4>"
        workload = c5_workload(B)
        d2h = dof * 8
    else:
        n = args.nodes
        mesh = synthetic.regular_bar(n)
        params = dz.DiffussionParams.time_independent().mu(MU).b(BADV).boundary_conditions(*BC).build()
        dmesh = dz.Mesh.from_nodes(mesh)
        t0 = time.perf_counter()
        amode = {"auto": dz.ASSEMBLY_AUTO, "faithful": dz.ASSEMBLY_FAITHFUL, "fast": dz.ASSEMBLY_FAST}[args.assembly]
        solver = dz.DiffussionSolverTimeIndependent(params, dmesh, G, dz.make_options(amode, dz.SOLVE_AUTO))
        extra["assembly"] = args.assembly
        dz.synchronize()
        extra["create_s"] = time.perf_counter() - t0
        dof = n
        host_out = torch.empty(dof, dtype=torch.float64, pin_memory=True)
        host_mesh = torch.from_numpy(mesh).pin_memory()

        def step_dev():                       # assemble (XSolver::new into the same buffers) + solve, all on the device
            solver.rebuild()
            solver.solve_device(0.0)

        def step_e2e():                       # host coordinates in, host nodal values out
            dmesh.set_nodes(host_mesh.numpy())
            solver.rebuild(dmesh)
            solver.solve_into(0.0, host_out.data_ptr())
        alg_bytes = ALG_BYTES_STATIC * dof
        kernel = "assemble + tridiagonal solve kernels"
        workload = c4_workload(n)
        d2h = dof * 8

    # ---- device-resident timing
    for _ in range(max(args.warmup, 3)):
        step_dev()
    ramp_clocks(step_dev, local_rank)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dz.launch_count(reset=True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for k in range(args.steps):
        step_dev()
        evs[k + 1].record(stream)
    barrier()
    launches = dz.launch_count()
    clocks = sampler.result()
    total_ms = max_over_ranks(evs[0].elapsed_time(evs[-1]))
    per_step = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    ms_per_step = total_ms / args.steps
    value = dof * world * args.steps / (total_ms * 1e-3)
    kern_ms = statistics.mean(per_step)
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9

    # ---- temporal blocking, reported separately: one launch advances every bar by 64 steps with operator and state on chip
    if args.workload == "c5" and not args.stored_operators:
        blk, reps = 64, 5
        solver.step(C5_DT, blk)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            solver.step(C5_DT, blk)
        e1.record(stream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        extra["multi_step"] = {"call": f"dzb_ensemble_step(dt, {blk})", "value": dof * world * blk * reps / (ms * 1e-3), "unit": UNIT,
                               "ms_per_step": ms / (blk * reps),
                               "note": "intermediate states are not written back; not the solve(dt) contract, hence not `value`"}

    # ---- end to end through the host-facing call
    e2e_steps = args.e2e_steps or min(args.steps, 100)
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = dof * world * e2e_steps / e2e_s
    h2d = 0 if args.workload == "c5" else dof * 8
    if "create_s" in extra:   # the same e2e steps with the one-off construction (H2D of meshes + IC, assembly, factors) charged too
        extra["e2e_including_create"] = {"value": dof * world * e2e_steps / (e2e_s + extra["create_s"]), "unit": UNIT,
                                         "steps": e2e_steps, "create_s": extra["create_s"]}

    # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture
    # (profiles/r01_c5_td_step_lean_full.txt), per launch; None where no capture of this exact kernel/size exists
    traffic = None
    if args.workload == "c5" and args.bars == 65536:
        traffic = 3.735e9 if args.stored_operators else 2.655e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload, "dof_per_gpu": dof, "l2": "inputs larger than L2 (operators + state > 3 GB per step)"
                   if dof * 56 > 2.6e8 else "working set may fit L2", "e2e_steps": e2e_steps,
                   "e2e_call": "dzb_ensemble_solve(dt, host_out)" if args.workload == "c5" else
                   "dzb_mesh_set_nodes(host_x) + dzb_solver_rebuild + dzb_solver_solve(host_out)",
                   "e2e_inputs": ("solve(dt) takes one scalar: the bars' state and operators live on the device by the API's nature, so a "
                                  "step has no H2D; the one-off H2D of meshes + initial conditions (1.07 GB) and the assembly are charged "
                                  "in extra.e2e_including_create") if args.workload == "c5" else
                                 "node coordinates H2D (pinned) + assembly + solve + nodal values D2H (pinned) every step",
                   "parallelism": f"{world} independent slices, no collective"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / e2e_steps * 1e3},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms, "peak_source": peak_src},
        "extra": extra,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        if args.workload == "c5":
            v_all, t_all = cpu_td_sample(args.cpu_sample_bars, 2000, threads)
            v_one, t_one = cpu_td_sample(max(args.cpu_sample_bars // 8, 64), 1000, 1)
            sample = (f"{args.cpu_sample_bars} of {B} bars x {n} nodes, 2000 solve(dt) on {threads} threads ({t_all:.1f} s); "
                      f"1-thread figure: {max(args.cpu_sample_bars // 8, 64)} bars, 1000 solve(dt) ({t_one:.1f} s)")
        else:
            nn = min(n, 1_000_000)
            v_all, t_all = cpu_static_sample(nn, threads)
            v_one, _ = cpu_static_sample(nn, 1)
            sample = f"{nn}-node regular bar: quadrature-loop assembly + Thomas ({t_all:.1f} s)"
        line["cpu_baseline"] = {"value": v_all, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                "single_thread_value": v_one}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
