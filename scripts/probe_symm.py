import os, sys, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
try:
    t = symm_mem.empty((1024,), dtype=torch.int32, device=dev)
    t.fill_(rank + 100)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok", type(hdl).__name__, [hex(p) for p in hdl.buffer_ptrs], "signal pads", len(hdl.signal_pad_ptrs), flush=True)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.int32)
    torch.cuda.synchronize()
    print(rank, "peer value", int(peer[0].item()), "attrs", [a for a in dir(hdl) if not a.startswith("_")], flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "SYMM FAILED", repr(e), flush=True)
dist.barrier(); dist.destroy_process_group()
