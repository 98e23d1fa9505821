"""Copy-only ceiling of the end-to-end path: every rank moves what one bench step moves between pinned host
memory and its GPU (H2D of 64 image pairs, D2H of 64 u8 maps) with nothing else running.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/copy_ceiling.py
"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
if "--pin" in sys.argv:
    import bench
    print(rank, bench.pin_rank_to_cores(local, world), flush=True)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, n = 64, 1920 * 1200
hin = torch.empty((2, B, n), dtype=torch.uint8).pin_memory(); hout = torch.empty((B, n), dtype=torch.uint8).pin_memory()
din = torch.empty_like(hin, device="cuda"); dout = torch.empty((B, n), dtype=torch.uint8, device="cuda")
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
def step():
    with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
for _ in range(3): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
K = 20
for _ in range(K): step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    dt = float(t.item())
    gb = K * (2 * B * n + B * n) / 1e9
    print("copy ceiling: %d rank(s), %.2f ms per step, %.1f GB/s per rank, %.1f GB/s total = %.0f frames/s total" % (
        world, 1e3 * dt / K, gb / dt, world * gb / dt, world * K * B / dt), flush=True)
if world > 1: dist.destroy_process_group()
