"""stdout hygiene of bench.py's multi-rank path: NCCL prints "NCCL version ..." to fd 1 when the communicator is created under
NCCL_DEBUG=VERSION (the GPU boxes export it); bench.stdout_to_stderr keeps stdout to the one JSON line.  World size 1, one GPU:
    python tools/nccl_stdout_check.py > out.txt 2> err.txt     # out.txt must hold exactly the JSONLINE line"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

torch.cuda.set_device(0)
with bench.stdout_to_stderr():
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    dist.barrier()
t = torch.ones(4, device="cuda")
dist.all_reduce(t)
torch.cuda.synchronize()
print("JSONLINE", float(t.sum()), flush=True)
dist.destroy_process_group()
