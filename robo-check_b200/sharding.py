"""Multi-GPU host logic: configurations are independent, so they shard across ranks with no
data-path exchange; the only collective is the all-gather of the per-configuration validity
bitmask.  One process per GPU, torch.distributed for the plumbing (NCCL on GPUs, gloo
in the CPU tests).  The reference has no multi-device code at all."""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous shard [lo, hi) of rank `rank`: ceil(n / world / 32) * 32 configurations per rank,
    so every shard is a whole number of 32-bit bitmask words (the last shards may be short or
    empty)."""
    per = ((n + world - 1) // world + 31) // 32 * 32
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def words_per_shard(n, world):
    return shard_bounds(n, world, 0)[2] // 32


def pack_bitmask(flags, words):
    """uint8 flags -> `words` little-endian 32-bit words (bit i%32 of word i/32 = flag i)."""
    bits = np.zeros(words * 32, np.uint8)
    bits[: len(flags)] = flags
    return np.packbits(bits, bitorder="little").view(np.uint32).copy()


def unpack_bitmask(words, n):
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n]


def check_sharded(check_local, confs, dist, device=None):
    """Every rank checks its shard with `check_local(confs_shard, d_mask_words)` -- on a GPU that is
    Context.check writing the bitmask straight into this rank's slot of the gather buffer -- then
    one all_gather_into_tensor leaves the full bitmask on every rank.  Returns the flags of all
    `len(confs)` configurations (uint8)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    n = len(confs)
    lo, hi, per = shard_bounds(n, world, rank)
    words = per // 32
    gather = torch.zeros(world * words, dtype=torch.int32, device=device)
    mine = gather[rank * words:(rank + 1) * words]
    check_local(confs[lo:hi], mine)
    dist.all_gather_into_tensor(gather, mine.clone() if device is None else mine)
    return unpack_bitmask(gather.cpu().numpy(), world * per)[:n] if world * per >= n else None


class PeerGather:
    """The all-gather of the bitmask fused behind the check (rc_check_gather in robocheck.h): one
    symmetric buffer per rank, mapped into every rank of the group with
    torch.distributed._symmetric_memory, so that a small kernel behind the two check kernels stores
    the rank's bitmask words straight into the peers' buffers over NVLink and waits for theirs -- no
    communicator call in the step.  Raises if the symmetric allocation or the rendezvous is not
    available; callers then fall back to check_sharded (NCCL all-gather)."""

    def __init__(self, rc, dist, device, words_per_rank):
        import torch
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank, self.words = dist.get_world_size(), dist.get_rank(), int(words_per_rank)
        total = int(rc.lib().rc_gather_buffer_words(self.world, self.words))
        if total <= 0:
            raise ValueError("world size / shard size not supported by rc_check_gather")
        self.buf = symm.empty(total, dtype=torch.int32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.ptrs = [int(self.hdl.get_buffer(r, (total,), torch.int32).data_ptr()) for r in range(self.world)]
        assert self.ptrs[self.rank] == self.buf.data_ptr()
        torch.cuda.synchronize(device)
        dist.barrier()  # every buffer is zeroed before anybody delivers into it

    def check_host(self, ctx, h_confs, d_confs, n, h_mask, stream):
        """Host buffers in, gathered bitmask of ALL ranks in host memory out: the rank's `n`
        configurations travel from the page-locked tensor `h_confs` to `d_confs`, are checked and
        gathered (rc_check_gather), and the gathered words come back into the page-locked tensor
        `h_mask` (world * words int32).  Returns when the words are there."""
        import torch
        with torch.cuda.stream(stream):
            d_confs[: n * 6].copy_(h_confs[: n * 6], non_blocking=True)
            words = self.check(ctx, d_confs, n, None, stream)
            h_mask.copy_(words, non_blocking=True)
        stream.synchronize()
        return h_mask

    def check(self, ctx, d_confs, n, d_valid=None, stream=None):
        """Returns the gathered bitmask of this call: a view of world * words int32 words inside the
        own buffer, valid (in stream order) until the second next call."""
        off = ctx.check_gather(d_confs, n, d_valid, self.ptrs, self.rank, self.words, stream)
        return self.buf[off:off + self.world * self.words]
