"""torchrun worker: N ranks build a tiled map and compare it with the untiled map that rank 0 computes alone on the same
clouds (bit for bit): gem_b200.tiled.parity_check under torchrun.  TILED_MODE = peer | padded | packed."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gem_b200 import tiled  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = tiled.parity_check(mode=os.environ.get("TILED_MODE", "peer"))
ok = True
if rank == 0:
    print(res)
    ok = res["status"] == "ok"
    print("TILED_CHECK_OK" if ok else "TILED_CHECK_FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
