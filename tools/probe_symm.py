"""2-GPU probe: does torch symmetric memory rendezvous here, and is there a multicast (NVLS) pointer?
torchrun --nproc-per-node 2 tools/probe_symm.py"""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
t = symm.empty(2 * 512 * 512 * 3, dtype=torch.float32, device="cuda")
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "multicast_ptr", hex(h.multicast_ptr), "buffers", [hex(p) for p in h.buffer_ptrs], "signal pad", h.signal_pad_size, flush=True)
t.zero_()
h.barrier(channel=0)
torch.cuda.synchronize()
# time barrier and NCCL all-reduce of the same buffer
x = torch.zeros_like(t)
for name, fn in (("symm barrier", lambda: h.barrier(channel=0)), ("nccl all_reduce 6.3MB", lambda: dist.all_reduce(x))):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(name, e0.elapsed_time(e1) / 50 * 1e3, "us", flush=True)
dist.destroy_process_group()
