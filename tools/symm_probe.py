"""Probe: does torch symmetric memory (peer-mapped buffers across the ranks of one box) work here?"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = sm.empty(1024, dtype=torch.float32, device=dev)
t.zero_()
hdl = sm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal pads", [hex(p) for p in hdl.signal_pad_ptrs], "pad bytes", hdl.signal_pad_size, flush=True)
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
hdl.barrier()
peer.fill_(float(rank + 1))
hdl.barrier()
torch.cuda.synchronize()
print(rank, "my buffer now holds", t[:3].tolist(), "expected", float((rank - 1) % world + 1), flush=True)
dist.destroy_process_group()
