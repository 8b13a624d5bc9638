"""ms per step of a FAITHFUL (bit-identical) ensemble of B x 1024-node bars; DZB_TW_MAX_BARS moves the switch between the
warp-per-bar and thread-per-bar kernels.  Usage: python tools/time_faithful_ensemble.py B"""
import os, sys, time, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
import dzahui_b200 as dz
from dzahui_b200 import synthetic
B, n = int(sys.argv[1]), 1024
ids = np.arange(B)
meshes = synthetic.ensemble_meshes(ids, n)
ic = synthetic.ensemble_initial_conditions(meshes, ids)
e = dz.EnsembleTimeDependent(meshes, 1.0, 1.0, (1.0, 15.0), ic, 150, dz.make_options(dz.ASSEMBLY_FAST, dz.SOLVE_FAITHFUL))
e.step(1e-8, 2); torch.cuda.synchronize()
t0 = time.perf_counter(); e.step(1e-8, 5); st = e.read_bars(np.array([0])); dt = (time.perf_counter() - t0) / 5
print(json.dumps({"B": B, "knob": os.environ.get("DZB_TW_MAX_BARS"), "ms_per_step": dt * 1e3, "checksum": float(st.sum())}))
