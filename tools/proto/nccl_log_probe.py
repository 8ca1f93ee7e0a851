import os, sys
mode = sys.argv[1]
if mode == "dup":
    sys.stdout.flush(); keep = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
os.environ.setdefault("NCCL_DEBUG", "INFO")
if mode != "nosubsys":
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(1, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
dist.destroy_process_group()
