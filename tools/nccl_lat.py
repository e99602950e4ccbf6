"""All-reduce latency of the residual-delta payload over NCCL with a kernel in between calls (as in the
Gibbs step), 1 process per GPU:  python -m torch.distributed.run --nproc-per-node G tools/nccl_lat.py"""
import os, sys, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
y = torch.zeros(1 << 20, dtype=torch.float64, device="cuda")
big = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
for dt, n in ((torch.float64, 530432), (torch.float64, 265216), (torch.float64, 131072), (torch.float32, 1060864), (torch.float32, 530432)):
    x = torch.ones(n, dtype=dt, device="cuda")
    for _ in range(20): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = []
    for mode in ("b2b", "small-kernel", "150us-kernel"):
        a.record(); t0 = time.perf_counter()
        for _ in range(100):
            if mode == "small-kernel": y.add_(1.0)
            if mode == "150us-kernel": big.add_(1.0)
            dist.all_reduce(x)
        host = (time.perf_counter() - t0) / 100 * 1e6
        b.record(); torch.cuda.synchronize()
        res.append(f"{mode} {a.elapsed_time(b)/100*1e3:.1f} us (host enqueue {host:.1f})")
    a.record()
    for _ in range(100): big.add_(1.0)
    b.record(); torch.cuda.synchronize()
    if dist.get_rank() == 0:
        print(f"{str(dt)[6:]} {x.numel()*x.element_size()/1e6:.2f} MB: " + "; ".join(res) + f"; big kernel alone {a.elapsed_time(b)/100*1e3:.1f} us", flush=True)
dist.destroy_process_group()
