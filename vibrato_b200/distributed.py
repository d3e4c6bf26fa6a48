"""Multi-GPU plumbing: one process per GPU, sentences sharded, dictionary image broadcast once.

The tokenisation path shards over sentences with a read-only dictionary; there is no
exchange step inside the data path.  `torch.distributed` (NCCL on GPUs, gloo in CPU tests) is used for
(1) broadcasting the packed dictionary image from rank 0 and (2) all-gathering per-rank token counts
so that every rank knows the global layout of the result.
"""
import numpy as np


def shard_by_bytes(byte_offsets, world):
    """Contiguous sentence ranges [(lo, hi)] * world with balanced byte counts (work ~ characters)."""
    off = np.asarray(byte_offsets, dtype=np.uint64)
    n = len(off) - 1
    total = int(off[-1] - off[0])
    bounds = [0]
    for r in range(1, world):
        target = int(off[0]) + total // world * r  # the rule of the library's own split (multi_engine.cu split_by_bytes)
        k = int(np.searchsorted(off, target, side="left"))
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def broadcast_dictionary_image(image, src=0, device=None):
    """Rank `src` passes the packed image (uint8 ndarray / tensor); everyone returns a tensor holding
    the same bytes on `device` (None = CPU, for gloo)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    size = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        t = torch.from_numpy(image) if isinstance(image, np.ndarray) else image
        size[0] = t.numel()
    dist.broadcast(size, src)
    buf = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
    if rank == src:
        buf.copy_(t)
    dist.broadcast(buf, src)
    return buf


def all_gather_counts(n_tokens_local, n_sent_local, device=None):
    """-> int64 ndarray [world, 2] of (n_sent, n_tokens) per rank."""
    import torch
    import torch.distributed as dist
    dev = torch.device(device) if device is not None else torch.device("cpu")
    mine = torch.tensor([n_sent_local, n_tokens_local], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return torch.stack(out).cpu().numpy()


def merge_shard_offsets(per_rank_tok_offsets):
    """Concatenates per-shard token offset arrays (each starting at 0) into one global array."""
    out = [np.zeros(1, dtype=np.uint64)]
    base = np.uint64(0)
    for off in per_rank_tok_offsets:
        off = np.asarray(off, dtype=np.uint64)
        out.append(off[1:] + base)
        base = base + off[-1]
    return np.concatenate(out)


def gather_token_records(records, counts, dst=0):
    """The "gather of token spans over NVLink" of a multi-process run: every rank's token records (int64 tensor,
    3 words per 24-byte vbt_token) are sent to rank `dst`, which receives them at their final positions.
    `counts` = tokens per rank (from all_gather_counts).  Returns the concatenated tensor on `dst`, None elsewhere.
    Point-to-point sends / receives, so it runs on NCCL (GPU tensors) and on gloo (CPU tests) alike."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank != dst:
        if records.numel():
            dist.send(records, dst)
        return None
    starts = np.concatenate([[0], np.cumsum(np.asarray(counts, dtype=np.int64) * 3)])
    out = torch.empty(int(starts[-1]), dtype=records.dtype, device=records.device)
    out[int(starts[dst]):int(starts[dst + 1])] = records
    for r in range(world):
        if r != dst and counts[r]:
            dist.recv(out[int(starts[r]):int(starts[r + 1])], r)
    return out
