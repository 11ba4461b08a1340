"""Multi-GPU plumbing: one process per GPU (torchrun), batch sharded by image.

Every image is independent on this path, so the data
path has no collective.  The only exchanges are
  * one SUM all-reduce of the two loss scalars per training step (the
    reference's losses are batch sums, loss.py:100-101, hence SUM not AVG);
  * one all-gather of the padded detections (+ counts) per detect step;
  * optionally an all-gather of the variable-length stacked GT rows, for callers
    that want the global ``stacked_gt_bboxes`` of reference loss.py:53.
Contiguous image blocks per rank keep batch order, so concatenating the rank
outputs reproduces the single-GPU mask / stacked-GT / detection order.
Works over NCCL (GPU) and gloo (CPU tensors; used by the CPU tests).
"""
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world():
    return (dist.get_rank(), dist.get_world_size()) if is_dist() else (0, 1)


def shard_range(batch_size, rank=None, world_size=None):
    """Contiguous block [lo, hi) of images owned by `rank`; sizes differ by at most 1."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(batch_size), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, batch_size, rank=None, world_size=None):
    """Slices every tensor whose leading dimension is the batch."""
    lo, hi = shard_range(batch_size, rank, world_size)
    return {k: (v[lo:hi] if hasattr(v, "shape") and len(v.shape) > 0 and v.shape[0] == batch_size else v)
            for k, v in tensors.items()}


def allreduce_losses(losses, group=None, async_op=False):
    """In-place SUM all-reduce of a small loss vector (e.g. the two float64
    losses in results[4:8].view(float64)).  Returns the work handle if async."""
    if not is_dist() or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(losses, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def gather_detections(post, group=None):
    """All-gathers the padded per-rank detection tensors (same B_local on every
    rank) along the batch dimension; returns a dict of global tensors in batch order.

    CUDA tensors travel as ONE packed byte buffer per rank (one NCCL all-gather for boxes, scores, prior
    indices and counts together -- at 8 ranks the four separate all-gathers of the list API cost several
    hundred microseconds of launch and staging overhead, more than the detect kernel itself); CPU tensors
    (gloo, the CPU tests) take the per-tensor list path."""
    if not is_dist() or dist.get_world_size(group) == 1:
        return dict(post)
    ws = dist.get_world_size(group)
    keys = [k for k, t in post.items() if t is not None]
    if keys and all(post[k].is_cuda for k in keys):
        # segments ordered by element size (8-byte boxes first) so that every segment start is aligned
        keys.sort(key=lambda k: -post[k].element_size())
        ts = [post[k].contiguous() for k in keys]
        segs = [t.view(-1).view(torch.uint8) for t in ts]
        total = sum(x.numel() for x in segs)
        pad = (-total) % 16
        if pad:
            segs.append(torch.zeros(pad, dtype=torch.uint8, device=ts[0].device))
        flat = torch.cat(segs)
        gathered = torch.empty((ws, flat.numel()), dtype=torch.uint8, device=flat.device)
        dist.all_gather_into_tensor(gathered.view(-1), flat, group=group)
        out, off = {}, 0
        for k, t in zip(keys, ts):
            n = t.numel() * t.element_size()
            part = gathered[:, off:off + n].view(t.dtype)                      # [ws, numel] strided over ranks
            out[k] = part.reshape((ws * t.shape[0],) + tuple(t.shape[1:]))     # contiguous, batch order
            off += n
        return out
    out = {}
    for k in keys:
        t = post[k]
        parts = [torch.empty_like(t) for _ in range(ws)]
        dist.all_gather(parts, t.contiguous(), group=group)
        out[k] = torch.cat(parts, dim=0)
    return out


def gather_stacked_gt(stacked_gt, group=None):
    """All-gather of the variable-length [N_r,4] matched GT rows -> global [N,4]
    in (rank, image, prior) = (image, prior) order."""
    if not is_dist() or dist.get_world_size(group) == 1:
        return stacked_gt
    ws = dist.get_world_size(group)
    n = torch.tensor([stacked_gt.shape[0]], dtype=torch.int64, device=stacked_gt.device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap, 4), dtype=stacked_gt.dtype, device=stacked_gt.device)
    padded[:stacked_gt.shape[0]] = stacked_gt
    parts = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def gather_variable_batch(t, batch_size, group=None):
    """All-gather of per-rank tensors whose leading dim is that rank's shard of
    `batch_size` images (shards may differ by one image)."""
    if not is_dist() or dist.get_world_size(group) == 1:
        return t
    ws = dist.get_world_size(group)
    sizes = [shard_range(batch_size, r, ws) for r in range(ws)]
    cap = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    padded[:t.shape[0]] = t
    parts = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


class PeerAllreduce:
    """Peer-memory plumbing for the all-reduce that is FUSED into the matching/loss kernel
    (mbx_match_loss_allreduce): one small symmetric buffer per rank, mapped into every
    process of the NVLink box through torch's symmetric memory (CUDA VMM handles exchanged
    over the process group).  PyTorch only allocates and maps; the exchange itself happens inside
    the kernel: tagged 8-byte words, pushed into the peers' tables (blocking mode) or left in the
    rank's own outbox, from where a relay kernel on a side stream forwards them into the peers' tables
    (deferred mode: the matching kernel itself does no NVLink access)."""

    def __init__(self, group=None, device=None):
        import ctypes
        from . import _lib
        if is_dist():        # rank / size INSIDE the group the buffers are exchanged over
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.group = group
        self.ptr_array = None
        self._keep = None
        if self.world == 1:
            return
        if self.world > _lib.MAX_PEERS:
            raise ValueError("fused all-reduce supports up to %d ranks" % _lib.MAX_PEERS)
        import torch.distributed._symmetric_memory as symm_mem
        lib = _lib.load()
        nbytes = int(lib.mbx_allreduce_buffer_bytes())
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, group=dist.group.WORLD if group is None else group)
        torch.cuda.synchronize(device)
        dist.barrier(group)                      # every buffer is zeroed before anyone writes
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        assert len(ptrs) == self.world
        self.ptr_array = (ctypes.c_ulonglong * self.world)(*ptrs)
        self._keep = (buf, hdl)

    def reset(self):
        """Collective: re-zeroes every rank's symmetric buffer (step counters, launch tickets, the sticky
        timeout flag, outbox and table rings).  Call on ALL ranks after a step reported
        MBX_STATUS_AR_TIMEOUT -- the ranks' counters are out of step after a timeout; until the reset every
        later reduction fails at once instead of waiting again -- with no step in flight."""
        if self.world == 1:
            return
        buf = self._keep[0]
        torch.cuda.synchronize(buf.device)
        dist.barrier(self.group)
        buf.zero_()
        torch.cuda.synchronize(buf.device)
        dist.barrier(self.group)


class LoopbackPeers:
    """`world` ranks of the fused all-reduce emulated on ONE device: the symmetric buffers are plain device
    allocations of this process and `rank(r)` is the `peer=` object of rank r.  The kernels cannot tell the
    difference (they only see the pointer table), so the whole exchange protocol -- outboxes, relay kernel,
    tables, the pull route, rings, the timeout flag -- runs on a single-GPU box (tests/test_gpu_loopback.py).
    Deferred-mode ranks may share a stream (no rank waits for a same-step peer); blocking-mode ranks
    need one stream each, because a rank's kernel spins until every other rank's kernel has posted."""

    class _Rank:
        def __init__(self, ptr_array, world, rank):
            self.ptr_array, self.world, self.rank = ptr_array, world, rank

    def __init__(self, world, device=None):
        import ctypes
        from . import _lib
        if not 2 <= world <= _lib.MAX_PEERS:
            raise ValueError("world must be 2..%d" % _lib.MAX_PEERS)
        lib = _lib.load()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = int(lib.mbx_allreduce_buffer_bytes())
        self.bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=device) for _ in range(world)]
        self.world = world
        self.ptr_array = (ctypes.c_ulonglong * world)(*[b.data_ptr() for b in self.bufs])
        torch.cuda.synchronize(device)

    def rank(self, r):
        return LoopbackPeers._Rank(self.ptr_array, self.world, r)

    def reset(self):
        torch.cuda.synchronize(self.bufs[0].device)
        for b in self.bufs:
            b.zero_()
        torch.cuda.synchronize(self.bufs[0].device)
