"""Per-rank CPU affinity after nvmlDeviceSetCpuAffinity and the GPU topology (why the end-to-end leg saturates at N > 1)."""
import os, sys
r = int(os.environ.get("LOCAL_RANK", "0"))
before = sorted(os.sched_getaffinity(0))
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(r)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    ok = "set"
except Exception as e:
    ok = f"failed: {e}"
after = sorted(os.sched_getaffinity(0))
def span(a): return f"{len(a)} cpus [{a[0]}..{a[-1]}]" if a else "none"
print(f"rank {r}: affinity before {span(before)} after {span(after)} ({ok})", file=sys.stderr)
