import os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, "signal_pad_ptrs", len(hdl.signal_pad_ptrs), flush=True)
    print(rank, [a for a in dir(hdl) if not a.startswith("_")], flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "symm_mem failed:", repr(e), flush=True)
dist.destroy_process_group()
