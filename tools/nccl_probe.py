"""Does NCCL work between the GPUs of this box?  Epoch-boundary collective (global argmin gather) over NCCL."""
import datetime, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import shard

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
t0 = time.time()
dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=60), device_id=torch.device("cuda", lr))
x = torch.ones(1 << 20, device="cuda") * (rank + 1)
dist.all_reduce(x)
torch.cuda.synchronize()
print(f"rank {rank}: all_reduce ok {x[0].item()} after {time.time() - t0:.1f}s", flush=True)
ev, who, p, m = shard.global_argmin(1.0 / (rank + 2), np.arange(19, dtype=np.uint8) + rank, np.arange(5, dtype=np.uint32) + rank, dist, device="cuda")
print(f"rank {rank}: global argmin {ev:.4f} from rank {who} parents[0..3]={p[:3]}", flush=True)
dist.destroy_process_group()
