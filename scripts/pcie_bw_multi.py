"""GPU box, torchrun with one rank per GPU: aggregate host<->device copy bandwidth from page-locked memory when k of the
N GPUs copy at once (k = 1, 2, 4, ... N) -- the ceiling of the end-to-end path at N GPUs.  Per active rank and step:
1 GiB host->device and 0.5 GiB device->host on two streams (the e2e path's 2:1 mix), wall clock around a barrier."""
import os, time
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
n_in, n_out = 1 << 30, 1 << 29
h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True); h_out.fill_(2)
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
k = 1
while k <= world:
    for mode in ("h2d", "both"):
        best = None
        for rep in range(3):
            dist.barrier()
            t = time.time()
            if rank < k:
                for _ in range(4):
                    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
                    if mode == "both":
                        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize()
            dist.barrier()
            dt = time.time() - t
            best = dt if best is None or dt < best else best
        if rank == 0:
            gb = 4 * k * (n_in + (n_out if mode == "both" else 0)) / 1e9
            print(f"{k} GPU(s) at once, {mode}: {gb / best:.1f} GB/s aggregate ({gb / best / k:.1f} per GPU)", flush=True)
    k *= 2
dist.destroy_process_group()
