import os, sys
mode = sys.argv[1]
if mode == "exec" and os.environ.get("PROBE_REEXEC") != "1":
    os.environ.update(NCCL_DEBUG="INFO", PROBE_REEXEC="1")
    os.execv(sys.executable, [sys.executable] + sys.argv)
if mode == "execsub" and os.environ.get("PROBE_REEXEC") != "1":
    os.environ.update(NCCL_DEBUG="INFO", NCCL_DEBUG_SUBSYS="INIT", PROBE_REEXEC="1")
    os.execv(sys.executable, [sys.executable] + sys.argv)
if mode == "dup":
    sys.stdout.flush(); keep = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(1, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
dist.destroy_process_group()
