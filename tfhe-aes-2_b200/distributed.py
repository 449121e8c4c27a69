"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink) only for the one-off replication of
the evaluation keys and the expanded AES key; AES-CTR blocks are sharded by contiguous counter ranges and there is NO
collective on the per-round path (reference: blocks are independent, src/bin/main.rs:141-159)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """contiguous, balanced partition: returns (start, stop) of rank's share of range(n_items)"""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class _DeviceSpan:
    """exposes a raw device allocation of the C library to torch without copying (CUDA array interface v2)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_bytes(ptr, nbytes, device):
    return torch.as_tensor(_DeviceSpan(ptr, nbytes), device=device)


def broadcast_tensors(tensors, src=0, chunk_bytes=1 << 28):
    """broadcast a list of flat uint8 tensors from `src` (works for NCCL device tensors and gloo CPU tensors alike)"""
    for t in tensors:
        flat = t.reshape(-1)
        for off in range(0, flat.numel(), chunk_bytes):
            dist.broadcast(flat[off:off + chunk_bytes], src=src)


def replicate_keys(ctx, client_key, device, src=0):
    """rank `src` uploads the evaluation keys (BSK converted to the Fourier domain on its GPU); every other rank allocates
    and receives them by broadcast; all ranks then finalise (correction rows of the keyswitch GEMMs)."""
    rank = dist.get_rank()
    if rank == src:
        ctx.upload_keys(client_key)
    else:
        ctx.alloc_keys()
    ctx.sync()
    spans = [device_bytes(*ctx.key_buffer(w), device) for w in (0, 1, 2)]
    broadcast_tensors(spans, src)
    torch.cuda.synchronize(device)
    if rank != src:
        ctx.keys_ready()
    return sum(s.numel() for s in spans)


def replicate_key_schedule(ctx, key_sched_host, device, src=0):
    rank = dist.get_rank()
    if rank == src:
        ctx.aes_set_key_schedule(key_sched_host)
    span = device_bytes(*ctx.aes_key_schedule_buffer(), device)
    broadcast_tensors([span], src)
    torch.cuda.synchronize(device)
    return span.numel()
