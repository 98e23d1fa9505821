"""Frame sharding across GPUs.

Stereo pairs are independent (the reference builds a fresh Elas per frame,
point_cloud.cpp:416-418), so a batch is split contiguously by frame index: rank g of G
gets frames [g*B/G, (g+1)*B/G).  There is no collective on the data path; the only
exchange is the final gather of the per-frame results (90-bin scan + meta, ~760 B/frame),
done with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def frame_shard(n_frames, rank, world):
    """Contiguous [begin, end) of the frames owned by `rank`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def gather_scans(local_ranges, n_frames, rank, world, device="cpu"):
    """Gathers per-frame scan rows (local_n x K float64) from all ranks into frame order.

    Every rank returns the full (n_frames x K) array.  Ranks may own different frame counts.
    """
    import torch
    import torch.distributed as dist
    local = torch.as_tensor(np.ascontiguousarray(local_ranges), dtype=torch.float64, device=device)
    K = local.shape[1]
    if world == 1:
        return local.cpu().numpy()
    counts = [frame_shard(n_frames, r, world) for r in range(world)]
    cap = max(e - b for b, e in counts)
    pad = torch.zeros((cap, K), dtype=torch.float64, device=device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.cat([bufs[r][: e - b] for r, (b, e) in enumerate(counts)], 0)
    return out.cpu().numpy()
