#!/usr/bin/env python3
"""bench.py — FEM DOF·steps/s of the Dzahui 1D hot path on B200 (BASELINE.json metric), one process per GPU.

Headline workload: BASELINE config 5 — an ensemble of 65536 independent 1024-node bars, time-dependent diffusion, one
`solve(dt)` on every bar per step (diffusion_solver/time_dependent.rs:257-280).  N > 1: every rank owns its own
65536-bar slice (global bar ids rank*65536 ...), no data-path communication -> "scaling": "weak".
The same default invocation also measures BASELINE config 4 (one 10^7-node bar: static diffusion assemble + solve,
time_independent.rs:57-196) with both assemblies and reports it in `extra.c4`; at N > 1 that ONE bar is split across the
ranks (SPIKE; strong scaling) and the C5 ensemble is additionally timed as BASELINE words it (65536 bars in total,
sharded: `extra.strong`).  `--workload c4` makes config 4 the headline instead.

  value     device-resident throughput: K steps timed with CUDA events on the launching stream, max over ranks
  e2e       the same steps through the reference-facing call `solve(dt) -> Vec<f64>` (dzb_ensemble_solve /
            dzb_solver_solve) with a pinned HOST result buffer: the D2H copy of every step's result is inside the region
  roofline  dominant kernel: bytes its design must move / CUDA-event time per launch, against MEASURED_PEAKS.json;
            `frac_survey_bytes` is the same with SURVEY §8(d)'s accounting, `traffic` the ncu-measured DRAM bytes per
            launch from profiles/ (null unless the kernel's source is unchanged since that capture)
  cpu_baseline  the CPU oracle (oracle/, port of the reference algorithm) on a bounded sample, rank 0, N = 1
`--impl reference` runs only that CPU port with every host thread on the same config/metric."""
import argparse
import hashlib
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
C5_N, C5_DT, C5_BARS = 1024, 1e-8, 65536
# bytes per DOF-step.  SURVEY §8(d): u in/out 16 + 3 M bands + 3 S bands 48 = 64.  The default step kernel (k_td_step_lean)
# is DESIGNED to move fewer: u in/out 16 + element length 8 + 1/den 8 + unscaled diagonal 8 = 40 (the six off-diagonal
# entries are rebuilt on chip from the element length); the stored-operator kernel moves u 16 + A' 24 + alpha 8 + c 8 = 56.
SURVEY_BYTES_TD_STEP = 64
DESIGN_BYTES_TD_STEP = {"lean": 40, "stored": 56}
SURVEY_BYTES_STATIC = 80      # assembly 40 B/DOF + solve 40 B/DOF
FP64_OPS_PER_NODE_TI = 27     # k_assemble_faithful<TI>, lean row: 3 integrals x 9 separately rounded operations per quadrature node
FP64_LANES_PER_SM, SM_COUNT = 64, 148


def c5_workload(bars):
    return (f"C5: ensemble of {bars} x {C5_N}-node bars per GPU, time-dependent diffusion, one solve(dt) per bar per step, "
            f"dt={C5_DT}, G={G}, mu=b=1, bc=(1,15)")


def c4_workload(nodes, world=1):
    split = "" if world == 1 else f" split into {world} row blocks, one per GPU (SPIKE),"
    return f"C4: one regular bar of {nodes} nodes,{split} static diffusion assemble + solve per step, G={G}, mu=b=1, bc=(1,15)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dzahui_b200", choices=["dzahui_b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c4"])
    ap.add_argument("--bars", type=int, default=C5_BARS, help="bars per GPU (c5)")
    ap.add_argument("--nodes", type=int, default=10_000_000, help="nodes of the single bar (c4)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 100)")
    ap.add_argument("--cpu-sample-bars", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (no extra.c4 / extra.strong)")
    ap.add_argument("--stored-operators", action="store_true",
                    help="c5: step from pre-scaled operator arrays (k_td_step_fast, 56 B/DOF-step) instead of k_td_step_lean")
    ap.add_argument("--assembly", default="auto", choices=["auto", "faithful", "fast"],
                    help="c4 headline: auto = the reference's arithmetic (bit-identical bands); fast = closed form (HBM-bound)")
    return ap.parse_args()


def config_for(args, world):
    """the `config` object; both arms (GPU and --impl reference) print exactly this for the same command line"""
    e2e_steps = args.e2e_steps or min(args.steps, 100)
    if args.workload == "c5":
        dof = args.bars * C5_N
        return {"workload": c5_workload(args.bars), "dof_per_gpu": dof,
                "value_covers": "the per-step hot path (one solve(dt) per bar); assembly + factorisation happen once at construction "
                                "and are charged in extra.create_s / extra.e2e_including_create",
                "l2": "inputs larger than L2 (operators + state > 2.6 GB per step)" if dof * 40 > 2.6e8 else "working set may fit L2",
                "e2e_steps": e2e_steps, "e2e_call": "dzb_ensemble_solve(dt, host_out)",
                "e2e_inputs": ("solve(dt) takes one scalar: the bars' state and operators live on the device by the API's nature, so a "
                               "step has no H2D; the one-off H2D of meshes + initial conditions (1.07 GB) and the assembly are charged "
                               "in extra.e2e_including_create"),
                "parallelism": f"{world} independent slices, no collective"}
    return {"workload": c4_workload(args.nodes, world), "dof_per_gpu": args.nodes // world,
            "value_covers": "assembly (XSolver::new into resident buffers) + tridiagonal solve, every step",
            "l2": "inputs larger than L2 (bands 320 MB)" if args.nodes * 32 > 2.6e8 else "working set may fit L2",
            "e2e_steps": e2e_steps, "e2e_call": ("dzb_mesh_set_nodes(host_x) + dzb_solver_rebuild + dzb_solver_solve(host_out)" if world == 1 else
                         "every rank: dzb_mesh_set_nodes(host block) + dzb_solver_block_rebuild + dzb_solver_block_solve_device + "
                         "dzb_solver_block_read(host_out)"),
            "e2e_inputs": "node coordinates H2D (pinned) + assembly + solve + nodal values D2H (pinned) every step",
            "parallelism": "1 GPU" if world == 1 else (f"one bar split into {world} row blocks; 8 doubles per rank cross GPUs per solve "
                                                        "(NVLink peer mailboxes, else ncclAllGather: extra.transport)")}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def source_sha16(files):
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic(kernel, launch_key):
    """ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed capture named in
    profiles/traffic.json — only if the files that define the kernel are byte-identical to the ones it was captured from."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f)[kernel][launch_key]
        if source_sha16(rec["files"]) != rec["sha16"]:
            return None, {"traffic_source": rec["source"], "traffic_note": "kernel source changed since this capture: not reported"}
        return float(rec["bytes_per_launch"]), {"traffic_source": rec["source"], "traffic_build": rec["sha16"]}
    except Exception:
        return None, {"traffic_source": None}


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
    """The oracle's DiffussionSolverTimeDependent::solve on `sample_bars` bars of the C5 ensemble.  The matrices are built
    once and the states advance in place inside the oracle (oracle_lib.EnsembleSession): a run() is `steps` solve(dt) per
    bar in one OpenMP region, with no per-call band or state copies — the same leg serves `cpu_baseline` and the
    `--impl reference` arm, so the two agree."""

    def __init__(self, sample_bars, threads, rank_offset=0):
        import numpy as np
        import oracle_lib as o
        from dzahui_b200 import synthetic
        ids = np.arange(sample_bars) + rank_offset
        meshes = synthetic.ensemble_meshes(ids, C5_N)
        ic = synthetic.ensemble_initial_conditions(meshes, ids)
        self.threads, self.bars = threads, sample_bars
        bands = o.ensemble_td_assemble(meshes, MU, BADV, G, threads=threads)
        st = np.stack([o.td_initial_state(ic[k], C5_N, BC) for k in range(sample_bars)])
        self.session = o.EnsembleSession(bands, st, BC)
        self.run(2)  # warm (thread pool, page faults)

    def run(self, steps):
        """-> (DOF-steps/s, seconds)"""
        t0 = time.perf_counter()
        self.session.advance(C5_DT, steps, self.threads)
        dt = time.perf_counter() - t0
        return self.bars * C5_N * steps / dt, dt


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


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            return next(l.split(":", 1)[1].strip() for l in f if l.startswith("model name"))
    except Exception:
        return "unknown"


def host_threads():
    """cores this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what the reference arm wants)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


REF_SOLVES_PER_STEP = 100


def run_reference(args, rank, world):
    """the reference's own CPU algorithm (oracle port; the Rust crate cannot be built in this image) with all host threads"""
    if rank != 0:
        return
    threads = host_threads()
    vals, secs = [], []
    if args.workload == "c5":
        sample = (f"{args.cpu_sample_bars} of {args.bars} bars x {C5_N} nodes, {REF_SOLVES_PER_STEP} solve(dt) per bench step, "
                  "matrices built once, states resident")
        leg = CpuTdSample(args.cpu_sample_bars, threads)
        for k in range(args.warmup + args.steps):
            v, sec = leg.run(REF_SOLVES_PER_STEP)
            if k >= args.warmup:
                vals.append(v)
                secs.append(sec)
    else:
        nn = min(args.nodes, 1_000_000)
        sample = f"{nn}-node regular bar (of {args.nodes}), one assemble + Thomas per bench step"
        for k in range(args.warmup + args.steps):
            v, sec = cpu_static_sample(nn, threads)
            if k >= args.warmup:
                vals.append(v)
                secs.append(sec)
    value = sum(vals) / len(vals) if args.workload == "c4" else args.cpu_sample_bars * C5_N * REF_SOLVES_PER_STEP * len(secs) / sum(secs)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": statistics.mean(secs) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_for(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "cpu_model": cpu_model()},
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


class Ctx:
    """what every measurement needs: torch, the library, this rank's stream, barriers and reductions"""

    def __init__(self, torch, dist, dz, stream, rank, local_rank, world):
        self.torch, self.dist, self.dz, self.stream = torch, dist, dz, stream
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.peak, self.peak_src = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, step, steps, warmup):
        """`steps` calls of step() between CUDA events on the launching stream, barrier + synchronize on both sides.
        -> (total ms as the max over ranks, per-step ms on this rank, kernels launched)"""
        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        self.dz.launch_count(reset=True)
        evs = [self.torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record(self.stream)
        for k in range(steps):
            step()
            evs[k + 1].record(self.stream)
        self.barrier()
        launches = self.dz.launch_count()
        total_ms = self.max_over_ranks(evs[0].elapsed_time(evs[-1]))
        return total_ms, [evs[k].elapsed_time(evs[k + 1]) for k in range(steps)], int(launches)


def measure_c4(cx, n, assembly, steps, warmup, parity_ref=None, e2e_steps=0):
    """BASELINE config 4 on ONE GPU: per step `dzb_solver_rebuild` (the assembly of XSolver::new into the handle's buffers)
    + `dzb_solver_solve_device`.  Returns the section printed under extra.c4[assembly] (or used as the headline)."""
    import numpy as np
    dz, torch = cx.dz, cx.torch
    from dzahui_b200 import synthetic
    mesh = synthetic.regular_bar(n)
    params = dz.DiffussionParams.time_independent().mu(MU).b(BADV).boundary_conditions(*BC).build()
    dmesh = dz.Mesh.from_nodes(mesh)
    amode = {"auto": dz.ASSEMBLY_AUTO, "faithful": dz.ASSEMBLY_FAITHFUL, "fast": dz.ASSEMBLY_FAST}[assembly]
    t0 = time.perf_counter()
    solver = dz.DiffussionSolverTimeIndependent(params, dmesh, G, dz.make_options(amode, dz.SOLVE_AUTO))
    dz.synchronize()
    out = {"assembly": assembly, "nodes": n, "create_s": time.perf_counter() - t0}

    def step():                           # assemble + solve, all on the device
        solver.rebuild()
        solver.solve_device(0.0)

    ramp_clocks(step, cx.local_rank)
    total_ms, per_step, launches = cx.time_steps(step, steps, warmup)
    ms = total_ms / steps
    out.update({"ms_per_step": ms, "value": n * steps / (total_ms * 1e-3), "unit": UNIT, "gpu_launches_per_step": launches / steps,
                "steps": steps})
    asm_ms = cx.time_steps(solver.rebuild, steps, 3)[0] / steps
    sol_ms = cx.time_steps(lambda: solver.solve_device(0.0), steps, 3)[0] / steps
    out["assemble_ms"], out["solve_ms"] = asm_ms, sol_ms
    out["path"] = solver.path()
    one_launch = "onelaunch" in out["path"]
    ach = SURVEY_BYTES_STATIC * n / (ms * 1e-3) / 1e9
    out["roofline"] = {"bound": "hbm",
                       "kernel": "k_tri_onelaunch (rows formed on chip, one cooperative launch per assemble + solve)" if one_launch
                       else "assemble + tridiagonal solve kernels (whole step)",
                       "achieved": ach, "peak": cx.peak,
                       "unit": "GB/s", "frac": ach / cx.peak, "algorithmic_bytes_per_launch": SURVEY_BYTES_STATIC * n,
                       "bytes_accounting": "SURVEY 8(d): assembly 40 B/DOF (x 8 in, 4 bands 32 out) + solve 40 B/DOF (4 bands in, u out)",
                       "design_bytes_per_row": ("x in twice 16 + row excess out and in 16 + u out 8 = 40, plus the level-B/C rows and solutions "
                                                "(~22 B/row, partly L2 hits); the bands are never materialised (rebuild is a no-op: "
                                                "assemble_ms ~ 0, everything is in solve_ms)") if one_launch
                       else "x 8 + bands out 32 + bands in twice 64 + u 8 = 112",
                       "traffic": None, "peak_source": cx.peak_src, "frac_of_8TBs_spec": ach / 8000.0}
    if one_launch:  # ncu's DRAM bytes of one launch at this size, while the kernel's sources are the ones it was captured from
        traffic, lbl = measured_traffic("k_tri_onelaunch", str(n))
        out["roofline"].update({"traffic": traffic, **lbl})
        if traffic:
            out["roofline"]["frac_on_measured_traffic"] = traffic / (ms * 1e-3) / 1e9 / cx.peak
    if assembly != "fast":
        # the reference-arithmetic assembly is fp64-pipe bound, not HBM bound: G-1 quadrature nodes x 27 separately rounded
        # operations per row (no FMA: the reference never contracts) against 64 fp64 lanes per SM per clock
        try:
            import pynvml
            pynvml.nvmlInit()
            mhz = pynvml.nvmlDeviceGetMaxClockInfo(pynvml.nvmlDeviceGetHandleByIndex(cx.local_rank), pynvml.NVML_CLOCK_SM)
        except Exception:
            mhz = 1965
        ops = float(n - 2) * (G - 1) * FP64_OPS_PER_NODE_TI
        peak_ops = SM_COUNT * FP64_LANES_PER_SM * mhz * 1e6
        out["assembly_roofline"] = {"bound": "fp64 pipe", "kernel": "k_assemble_faithful<TI>", "achieved": ops / (asm_ms * 1e-3) / 1e12,
                                    "peak": peak_ops / 1e12, "unit": "T op/s (one rounding each, no FMA)", "frac": ops / (asm_ms * 1e-3) / peak_ops,
                                    "ops_per_launch": ops, "kernel_ms": asm_ms,
                                    "peak_source": f"{SM_COUNT} SMs x {FP64_LANES_PER_SM} fp64 lanes/clk x {mhz} MHz (tools/micro/fp64_latency.cu)"}
    if e2e_steps:
        host_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
        host_mesh = torch.from_numpy(mesh).pin_memory()

        def step_e2e():                   # host coordinates in, host nodal values out
            dmesh.set_nodes(host_mesh.numpy())
            solver.rebuild(dmesh)
            solver.solve_into(0.0, host_out.data_ptr())
        for _ in range(2):
            step_e2e()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        out["e2e"] = {"value": n * e2e_steps / sec, "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
                      "ms_per_step": sec / e2e_steps * 1e3, "steps": e2e_steps}
    if parity_ref is not None:
        # SURVEY §7 hard part 1: at 10^7 rows the reference's own f64 Thomas is ~2.6e-7 from the exact solution of its system, so
        # the plain figure relL2(gpu, reference f64) cannot be below that for ANY algorithm; all three figures are printed.
        import oracle_lib as o
        u = solver.solve(0.0)
        if not parity_ref:                # first (AUTO / FAITHFUL: bands bit-identical to the reference arithmetic) call fills the cache
            lo, di, up, rhs = solver.bands()
            parity_ref["ref_f64"] = o.thomas(lo, di, up, rhs)
            parity_ref["f128"] = o.thomas(lo, di, up, rhs, kind="f128")
            # the yardstick for ANY assembly on a bar this long: the reference's system evaluated and solved in __float128 on its
            # f64 inputs, no band entry ever rounded (oracle dzo_ti_intended_f128; a few seconds on the host's cores)
            from dzahui_b200 import synthetic
            parity_ref["intended"] = o.ti_intended_f128(synthetic.regular_bar(n), MU, BADV, BC, G, threads=host_threads())
        ref, truth = parity_ref["ref_f64"], parity_ref["f128"]
        nrm = float(np.linalg.norm(truth))
        out["parity"] = {"rel_l2_gpu_vs_reference_f64": float(np.linalg.norm(u - ref)) / float(np.linalg.norm(ref)),
                         "rel_l2_reference_f64_vs_f128": float(np.linalg.norm(ref - truth)) / nrm,
                         "rel_l2_gpu_vs_f128": float(np.linalg.norm(u - truth)) / nrm,
                         "reference": "oracle solve_by_thomas (f64) and __float128 Thomas on the reference-arithmetic bands"}
        intended = parity_ref["intended"]
        ni = float(np.linalg.norm(intended))
        out["parity"]["rel_l2_gpu_vs_exact_arithmetic_system"] = float(np.linalg.norm(u - intended)) / ni
        out["parity"]["rel_l2_reference_f64_vs_exact_arithmetic_system"] = float(np.linalg.norm(ref - intended)) / ni
        out["parity"]["exact_arithmetic_system"] = ("time_independent.rs:91-182 evaluated and solved in __float128 on the reference's f64 "
                                                    "inputs (no band entry rounded): three-way rule gpu <= max(1e-10, reference)")
        if assembly == "fast":
            lo, di, up, rhs = solver.bands()
            own = o.thomas(lo, di, up, rhs, kind="f128")
            out["parity"]["rel_l2_gpu_vs_f128_of_its_own_bands"] = float(np.linalg.norm(u - own)) / float(np.linalg.norm(own))
            out["parity"]["note"] = ("closed-form bands agree with the reference's to <= 1e-12 per entry, but the solution of a 10^7-row system "
                                     "is sensitive to exactly that rounding noise: the reference's own f64 result is further from the system "
                                     "it means to solve than the closed-form result is (DESIGN 2)")
    solver.close()
    dmesh.close()
    return out


def measure_c4_split(cx, n, assembly, steps, warmup, comm, parity_ref=None, e2e_steps=0):
    """BASELINE config 4 at N > 1: ONE n-node bar split into row blocks, one per rank (strong scaling).  Per step every rank
    calls `dzb_solver_block_rebuild` (assembly of its block) + `dzb_solver_block_solve_device` (block elimination, exchange of
    the 2 interface rows per block, interface solve, back-substitution): the communicator is the library's own (dzb_comm_*)."""
    import numpy as np
    dz, dist = cx.dz, cx.dist
    from dzahui_b200 import synthetic
    from dzahui_b200.distributed import SplitBarTimeIndependent
    mesh = synthetic.regular_bar(n)
    params = dz.DiffussionParams.time_independent().mu(MU).b(BADV).boundary_conditions(*BC).build()
    amode = {"auto": dz.ASSEMBLY_AUTO, "faithful": dz.ASSEMBLY_FAITHFUL, "fast": dz.ASSEMBLY_FAST}[assembly]
    t0 = time.perf_counter()
    bar = SplitBarTimeIndependent(params, mesh, G, comm, dz.make_options(amode, dz.SOLVE_PARALLEL))
    dz.synchronize()
    out = {"assembly": assembly, "nodes": n, "n_gpus": cx.world, "rows_per_gpu": bar.rows[1] - bar.rows[0],
           "create_s": time.perf_counter() - t0, "transport": "peer mailboxes (NVLink stores + flag, no collective launch)"
           if comm.peer_mailboxes else "ncclAllGather of 8 doubles per rank"}

    def step():
        bar.rebuild()
        bar.solve_device()

    cx.barrier()   # rank 0 may come out of the previous mode's CPU parity check: nobody spins on its mailbox meanwhile
    ramp_clocks(step, cx.local_rank)
    total_ms, per_step, launches = cx.time_steps(step, steps, warmup)
    bar.status()
    ms = total_ms / steps
    out.update({"ms_per_step": ms, "value": n * steps / (total_ms * 1e-3), "unit": UNIT, "scaling": "strong",
                "gpu_launches_per_step": launches / steps, "steps": steps, "path": bar.path()})
    out["assemble_ms"] = cx.time_steps(bar.rebuild, steps, 3)[0] / steps
    out["solve_ms"] = cx.time_steps(bar.solve_device, steps, 3)[0] / steps
    bar.status()
    ach = SURVEY_BYTES_STATIC * n / (ms * 1e-3) / 1e9
    out["roofline"] = {"bound": "hbm", "kernel": out["path"], "achieved": ach, "peak": cx.peak * cx.world, "unit": "GB/s",
                       "frac": ach / (cx.peak * cx.world), "algorithmic_bytes_per_launch": SURVEY_BYTES_STATIC * n,
                       "bytes_accounting": "SURVEY 8(d): assembly 40 B/DOF + solve 40 B/DOF, against N x the measured copy peak",
                       "traffic": None, "peak_source": cx.peak_src, "frac_of_8TBs_spec": ach / (8000.0 * cx.world)}
    if e2e_steps:   # every rank: its block's coordinates (own + halo nodes) from pinned host memory, its nodal values back
        torch = cx.torch
        first, last = bar.rows
        nodes = mesh[first - (1 if first > 0 else 0): last + (1 if last < n else 0)]
        host_x = torch.from_numpy(np.ascontiguousarray(nodes)).pin_memory()
        host_u = torch.empty(last - first, dtype=torch.float64, pin_memory=True)

        def step_e2e():
            bar.rebuild(host_x.numpy())
            bar.solve_device()
            bar.block.block_read_into(host_u.data_ptr())
        for _ in range(2):
            step_e2e()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        cx.torch.cuda.synchronize()
        sec = cx.max_over_ranks(time.perf_counter() - t0)
        bar.status()
        out["e2e"] = {"value": n * e2e_steps / sec, "unit": UNIT, "h2d_bytes_per_step": len(nodes) * 8, "d2h_bytes_per_step": (last - first) * 8,
                      "ms_per_step": sec / e2e_steps * 1e3, "steps": e2e_steps, "bytes": "per rank"}
    if parity_ref is not None:
        u_block = bar.solve(0.0)
        parts = [None] * cx.world if cx.rank == 0 else None
        dist.gather_object(u_block, parts, dst=0)
        if cx.rank == 0:
            import oracle_lib as o
            u = np.concatenate(parts)
            if not parity_ref:   # the reference-arithmetic bands of the whole bar, from the CPU port (all host threads)
                lo, di, up, rhs = o.ti_assemble(mesh, MU, BADV, BC, G, mode=0, threads=host_threads())
                parity_ref["ref_f64"] = o.thomas(lo, di, up, rhs)
                parity_ref["f128"] = o.thomas(lo, di, up, rhs, kind="f128")
                parity_ref["intended"] = o.ti_intended_f128(mesh, MU, BADV, BC, G, threads=host_threads())  # see measure_c4
            ref, truth, intended = parity_ref["ref_f64"], parity_ref["f128"], parity_ref["intended"]
            nrm, ni = float(np.linalg.norm(truth)), float(np.linalg.norm(intended))
            out["parity"] = {"rel_l2_gpu_vs_reference_f64": float(np.linalg.norm(u - ref)) / float(np.linalg.norm(ref)),
                             "rel_l2_reference_f64_vs_f128": float(np.linalg.norm(ref - truth)) / nrm,
                             "rel_l2_gpu_vs_f128": float(np.linalg.norm(u - truth)) / nrm,
                             "rel_l2_gpu_vs_exact_arithmetic_system": float(np.linalg.norm(u - intended)) / ni,
                             "rel_l2_reference_f64_vs_exact_arithmetic_system": float(np.linalg.norm(ref - intended)) / ni,
                             "reference": "oracle quadrature-loop assembly + solve_by_thomas (f64), __float128 Thomas of the same bands, and the "
                                          "same system evaluated and solved in __float128 (no band entry rounded)"}
    bar.close()
    return out


def measure_c5_strong(cx, total_bars, steps, warmup, weak_ms):
    """BASELINE's wording of config 5: ONE ensemble of `total_bars` bars sharded across the ranks (strong scaling)."""
    import numpy as np
    dz = cx.dz
    from dzahui_b200 import synthetic
    from dzahui_b200.distributed import shard_bars
    first, last = shard_bars(total_bars, cx.rank, cx.world)
    ids = np.arange(first, last, dtype=np.int64)
    meshes = synthetic.ensemble_meshes(ids, C5_N)
    ic = synthetic.ensemble_initial_conditions(meshes, ids)
    e = dz.EnsembleTimeDependent(meshes, MU, BADV, BC, ic, G)
    total_ms, _, launches = cx.time_steps(lambda: e.step(C5_DT, 1), steps, warmup)
    e.close()
    ms = total_ms / steps
    out = {"bars_total": total_bars, "bars_per_gpu": last - first, "n_gpus": cx.world, "ms_per_step": ms,
           "value": total_bars * C5_N * steps / (total_ms * 1e-3), "unit": UNIT, "scaling": "strong", "gpu_launches_per_step": launches / steps}
    if weak_ms:   # the same kernel on the whole ensemble on one GPU is the headline's ms_per_step
        out["efficiency_vs_one_gpu"] = weak_ms / (cx.world * ms)
        out["one_gpu_ms_per_step"] = weak_ms
    return out


def main():
    protect_stdout()
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
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
    cx = Ctx(torch, dist, dz, stream, rank, local_rank, world)
    barrier, max_over_ranks, peak, peak_src = cx.barrier, cx.max_over_ranks, cx.peak, cx.peak_src

    extra = {}
    config = config_for(args, world)
    e2e_steps = config["e2e_steps"]
    scaling = "weak"
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
        del meshes, ic
        dof = B * n
        step_dev = lambda: solver.step(C5_DT, 1)
        host_out = torch.empty(dof, dtype=torch.float64, pin_memory=True)
        step_e2e = lambda: solver.solve_into(C5_DT, host_out.data_ptr())
        design_bytes = DESIGN_BYTES_TD_STEP["stored" if args.stored_operators else "lean"] * dof
        survey_bytes = SURVEY_BYTES_TD_STEP * dof
        d2h, h2d = dof * 8, 0
        ramp_clocks(step_dev, local_rank)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        total_ms, per_step, launches = cx.time_steps(step_dev, args.steps, args.warmup)
        clocks = sampler.result()
        kernel = solver.step_kernel()      # asked from the library: the kernel this handle's steps dispatch to
        ms_per_step = total_ms / args.steps
        value = dof * world * args.steps / (total_ms * 1e-3)
        kern_ms = statistics.mean(per_step)
        achieved = design_bytes / (kern_ms * 1e-3) / 1e9
        traffic, traffic_lbl = measured_traffic(kernel, f"{B}x{n}")
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "bytes_accounting": ("bytes the kernel's design must move per DOF-step: "
                                         + ("u 16 + A' 24 + alpha 8 + c 8 = 56" if args.stored_operators else
                                            "u in/out 16 + element length 8 + 1/den 8 + unscaled diagonal 8 = 40 (off-diagonals rebuilt on chip)")),
                    "algorithmic_bytes_per_launch": design_bytes,
                    "frac_survey_bytes": survey_bytes / (kern_ms * 1e-3) / 1e9 / peak, "survey_bytes_per_launch": survey_bytes,
                    "traffic": traffic, **traffic_lbl, "kernel_ms": kern_ms, "peak_source": peak_src,
                    "frac_of_8TBs_spec": achieved / 8000.0}

        # ---- temporal blocking, reported separately: one launch advances every bar by 64 steps with operator and state on chip
        if not args.stored_operators:
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
    else:
        scaling = "strong" if world > 1 else "weak"
        n = args.nodes
        dof = n
        sampler = ClockSampler(local_rank)
        sampler.start()
        if world == 1:
            sec = measure_c4(cx, n, args.assembly, args.steps, args.warmup, parity_ref=None, e2e_steps=e2e_steps)
        else:
            from dzahui_b200.distributed import NativeComm
            comm = NativeComm.from_torch()
            sec = measure_c4_split(cx, n, args.assembly, args.steps, args.warmup, comm, parity_ref={}, e2e_steps=e2e_steps)
            comm.close()
        clocks = sampler.result()
        ms_per_step, value, launches = sec["ms_per_step"], sec["value"], int(sec["gpu_launches_per_step"] * args.steps)
        roofline = sec.pop("roofline")
        e2e_sec = sec.pop("e2e", None)
        extra.update(sec)
        d2h, h2d = n * 8 // world, n * 8 // world

    # ---- end to end through the host-facing call
    if args.workload == "c5":
        for _ in range(3):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": dof * world * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_s / e2e_steps * 1e3}
        extra["e2e_including_create"] = {"value": dof * world * e2e_steps / (e2e_s + extra["create_s"]), "unit": UNIT,
                                         "steps": e2e_steps, "create_s": extra["create_s"]}
        # the same call with the states cast to f32 on the device first (what the reference's renderer uploads): half the bytes
        host32 = torch.empty(dof, dtype=torch.float32, pin_memory=True)
        step_f32 = lambda: solver.solve_f32(C5_DT, host32.data_ptr())
        for _ in range(3):
            step_f32()
        barrier()
        k32 = max(5, e2e_steps // 2)
        t0 = time.perf_counter()
        for _ in range(k32):
            step_f32()
        torch.cuda.synchronize()
        s32 = max_over_ranks(time.perf_counter() - t0)
        extra["e2e_f32"] = {"call": "dzb_ensemble_solve_f32(dt, host_out_f32)", "value": dof * world * k32 / s32, "unit": UNIT,
                            "ms_per_step": s32 / k32 * 1e3, "d2h_bytes_per_step": dof * 4, "steps": k32}
        # what bounds both: a bare pinned D2H copy of the same bytes, no kernels, every rank at once (tools/d2h_probe.py)
        dev_probe = torch.empty(dof * 8, dtype=torch.uint8, device="cuda")
        host_view = host_out.view(torch.uint8)
        host_view.copy_(dev_probe, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            host_view.copy_(dev_probe, non_blocking=True)
        torch.cuda.synchronize()
        sp = max_over_ranks(time.perf_counter() - t0) / 5
        extra["e2e_limiter"] = {"what": "bare cudaMemcpyAsync of one step's result (device -> pinned host), all ranks at once, no kernels",
                                "ms_per_copy": sp * 1e3, "GBps_per_rank": dof * 8 / sp / 1e9, "GBps_aggregate": world * dof * 8 / sp / 1e9,
                                "e2e_step_over_bare_copy": e2e["ms_per_step"] / (sp * 1e3)}
        del dev_probe, host32
        solver.close()
        del host_out
    else:
        e2e = e2e_sec or {"value": None, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                          "note": "not measured"}

    # ---- the other BASELINE configuration on which the metric is quoted, in the same run
    if not args.no_extra and args.workload == "c5":
        if world == 1:
            k = min(args.steps, 50)
            cache = {}
            c4 = {"workload": c4_workload(args.nodes)}
            for mode in ("auto", "fast"):   # auto first: its bands are the reference's, bit for bit, and anchor the parity figures
                c4[mode] = measure_c4(cx, args.nodes, mode, k, 3, parity_ref=cache, e2e_steps=min(k, 10))
            extra["c4"] = c4
        else:
            try:
                from dzahui_b200.distributed import NativeComm
                comm = NativeComm.from_torch()
                cache = {}
                c4 = {"workload": c4_workload(args.nodes, world)}
                for mode in ("auto", "fast"):
                    c4[mode] = measure_c4_split(cx, args.nodes, mode, min(args.steps, 50), 3, comm, parity_ref=cache, e2e_steps=10)
                comm.close()
                extra["c4_split"] = c4
            except Exception as ex:  # the headline must not depend on the extra section
                extra["c4_split"] = {"error": repr(ex)}
        if world == 1 and args.bars == C5_BARS:
            extra["strong"] = {"bars_total": C5_BARS, "bars_per_gpu": C5_BARS, "n_gpus": 1, "ms_per_step": ms_per_step, "value": value,
                               "unit": UNIT, "scaling": "strong", "efficiency_vs_one_gpu": 1.0, "note": "N = 1: the headline run itself"}
        else:
            extra["strong"] = measure_c5_strong(cx, C5_BARS, args.steps, args.warmup, ms_per_step if args.bars == C5_BARS else None)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "extra": extra,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        nn = min(args.nodes, 1_000_000)
        if args.workload == "c5":
            leg = CpuTdSample(args.cpu_sample_bars, threads)
            v_all, t_all = leg.run(2000)
            one = CpuTdSample(max(args.cpu_sample_bars // 8, 64), 1)
            v_one, t_one = one.run(1000)
            sample = (f"{args.cpu_sample_bars} of {args.bars} bars x {C5_N} nodes, 2000 solve(dt) on {threads} threads ({t_all:.1f} s); "
                      f"1-thread figure: {max(args.cpu_sample_bars // 8, 64)} bars, 1000 solve(dt) ({t_one:.1f} s)")
            if "c4" in extra:
                c_all, ct = cpu_static_sample(nn, threads)
                c_one, _ = cpu_static_sample(nn, 1)
                extra["c4"]["cpu_baseline"] = {"value": c_all, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": c_one,
                                               "cpu_model": cpu_model(),
                                               "sample": f"{nn}-node regular bar: quadrature-loop assembly + Thomas ({ct:.1f} s)"}
        else:
            v_all, t_all = cpu_static_sample(nn, threads)
            v_one, _ = cpu_static_sample(nn, 1)
            sample = f"{nn}-node regular bar: quadrature-loop assembly + Thomas ({t_all:.1f} s)"
        line["cpu_baseline"] = {"value": v_all, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                "single_thread_value": v_one, "cpu_model": cpu_model()}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
