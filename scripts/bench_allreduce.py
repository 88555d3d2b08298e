"""NCCL all-reduce time for the gradient vectors of the two configs (fp32), one rank per GPU:
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_allreduce.py"""
import os, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
for name, n in (("octo-small 21.7 M fp32", 21_700_000), ("octo-base 85 M fp32", 85_050_000), ("one octo-small layer 1.77 M fp32", 1_774_000)):
    g = torch.randn(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(g)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(g)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if rank == 0:
        w = dist.get_world_size()
        print(f"{name}: {ms*1e3:8.1f} us per all-reduce, algbw {n*4/ms/1e6:7.1f} GB/s, busbw {n*4/ms/1e6*2*(w-1)/w:7.1f} GB/s (N = {w})", flush=True)
dist.destroy_process_group()
