"""Which environment keeps NCCL's version banner off stdout?  (torchrun, 2 ranks; prints to stderr only)"""
import os, sys
mode = sys.argv[1]
if mode == "file":
    os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
elif mode == "none":
    os.environ["NCCL_DEBUG"] = "NONE"
elif mode == "unset":
    os.environ.pop("NCCL_DEBUG", None)
import torch, torch.distributed as dist
r = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(r)
print(f"[probe {mode}] rank {r} NCCL_DEBUG={os.environ.get('NCCL_DEBUG')!r} NCCL_DEBUG_FILE={os.environ.get('NCCL_DEBUG_FILE')!r}", file=sys.stderr)
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
t = torch.ones(1, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
if r == 0: print("STDOUT-PAYLOAD")
dist.destroy_process_group()
