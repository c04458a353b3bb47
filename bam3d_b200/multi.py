"""Multi-GPU plumbing (one process per GPU, torch.distributed): pair lists shard contiguously across ranks with no
data-path collective; the one exchange step is the gather of the compacted contacts (bam3d_contact40 records, 10 x u32
each, tagged with the global pair index) so that every rank ends with the full contact list.
Device-agnostic on purpose: the same code runs over NCCL on GPUs and over gloo on CPU tensors in the tests."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _ffi

REC_WORDS = 10  # bam3d_contact40 = 40 bytes


class Comm:
    """bam3d_comm: the library's own NCCL communicator (include/bam3d_gpu.h "multi-GPU"). The 128-byte rendezvous id is made by
    rank 0 and handed round by `exchange`, any callable that broadcasts a uint8 numpy array from rank 0 (default: the
    torch.distributed process group, used for this bootstrap only — the data path below does not touch PyTorch)."""

    def __init__(self, device, rank, n_ranks, exchange=None):
        self._lib = _ffi.lib()
        uid = np.zeros(_ffi.UNIQUE_ID_BYTES, dtype=np.uint8)
        if rank == 0:
            rc = self._lib.bam3d_comm_unique_id(C.c_void_p(uid.ctypes.data))
            if rc != _ffi.OK:
                raise _ffi.Bam3dError(rc, "bam3d_comm_unique_id: " + self._lib.bam3d_comm_unavailable_reason().decode())
        uid = (exchange or _torch_broadcast)(uid)
        h = C.c_void_p()
        rc = self._lib.bam3d_comm_create(int(device), C.c_void_p(uid.ctypes.data), int(rank), int(n_ranks), C.byref(h))
        if rc != _ffi.OK:
            raise _ffi.Bam3dError(rc, "bam3d_comm_create: " + self._lib.bam3d_comm_unavailable_reason().decode())
        self.handle, self.rank, self.n_ranks = h, int(rank), int(n_ranks)
        self._counts = np.zeros(n_ranks, dtype=np.uint64)

    def begin(self, ctx, d_count):
        """Start the exchange of the ranks' record counts behind everything enqueued on ctx.stream -> ticket."""
        t = C.c_uint64(0)
        ctx.check(self._lib.bam3d_gather_contacts_begin(ctx.handle, self.handle, d_count, C.byref(t)))
        return t.value

    def finish(self, ctx, ticket, d_compact, d_gathered, capacity):
        """Enqueue the payload exchange of `ticket` on the comm stream -> (per-rank counts, total)."""
        total = C.c_uint64(0)
        ctx.check(self._lib.bam3d_gather_contacts_finish(ctx.handle, self.handle, ticket, d_compact, d_gathered, int(capacity),
                                                         C.c_void_p(self._counts.ctypes.data), C.byref(total)))
        return self._counts.copy(), int(total.value)

    def wait(self, ctx, ticket):
        """ctx.stream waits for that ticket's payload (before its buffers are reused / read)."""
        ctx.check(self._lib.bam3d_gather_contacts_wait(ctx.handle, self.handle, int(ticket)))

    def join(self, ctx):
        ctx.check(self._lib.bam3d_comm_join(ctx.handle, self.handle))

    def gather_dev(self, ctx, d_compact, d_count, d_gathered, capacity):
        total = C.c_uint64(0)
        ctx.check(self._lib.bam3d_gather_contacts_dev(ctx.handle, self.handle, d_compact, d_count, d_gathered, int(capacity),
                                                      C.c_void_p(self._counts.ctypes.data), C.byref(total)))
        return self._counts.copy(), int(total.value)

    def gather(self, ctx, contacts, status, base, capacity):
        """Host buffers: this rank's dense (contacts[n, 8], status[n]) -> every rank's hits, bam3d_contact40 rows as uint32[total, 10]."""
        contacts = np.ascontiguousarray(contacts, dtype=np.float32).reshape(-1, 8)
        status = np.ascontiguousarray(status, dtype=np.uint8)
        out = np.zeros((max(1, int(capacity)), REC_WORDS), dtype=np.uint32)
        total = C.c_uint64(0)
        ctx.check(self._lib.bam3d_gather_contacts(ctx.handle, self.handle, C.c_void_p(contacts.ctypes.data), C.c_void_p(status.ctypes.data),
                                                  len(status), int(base), C.c_void_p(out.ctypes.data), int(capacity),
                                                  C.c_void_p(self._counts.ctypes.data), C.byref(total)))
        return out[: int(total.value)], self._counts.copy()

    def close(self):
        if self.handle:
            self._lib.bam3d_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _torch_broadcast(uid):
    t = torch.from_numpy(uid.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


def shard_range(n_total, rank, world):
    """Contiguous slice [begin, end) of n_total pairs owned by `rank`."""
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def sap_shard_range(n_boxes, shard, n_shards):
    """Sorted-slot range [begin, end) whose pairs (by their higher slot j) shard `shard` searches — the same split as
    bam3d_sap_find_pairs_shard (include/bam3d_gpu.h)."""
    return n_boxes * shard // n_shards, n_boxes * (shard + 1) // n_shards


def gather_pairs(pairs, count, counts_buf=None, out=None):
    """All-gather the per-rank slices of a sharded sweep-and-prune pair list (int32 tensor, 2 words per pair). Only
    needed when every rank wants the full list: the narrow phase consumes each rank's slice in place.
    -> (gathered, counts) laid out like gather_contacts; unpack_gathered(..., words=2) restores the unsharded order."""
    return gather_contacts(pairs, count, counts_buf=counts_buf, out=out, words=2)


def gather_contacts(compact, count, counts_buf=None, out=None, words=REC_WORDS):
    """All-gather variable-length compacted contact records.

    compact: int32 tensor, this rank's records, with room for the LARGEST count over all ranks (the payload is padded to
    it: checked); count: 1-element int64 tensor with the number of valid records. Returns (gathered, counts): `gathered` holds world * max_count records (rank r's records
    start at r * max_count), `counts` the per-rank record counts. Two collectives: counts (8 bytes per rank), then the
    payload padded to the largest count."""
    world = dist.get_world_size()
    counts = counts_buf if counts_buf is not None else torch.zeros(world, dtype=torch.int64, device=compact.device)
    dist.all_gather_into_tensor(counts, count)
    m = int(counts.max().item())
    if compact.numel() < m * words:
        raise ValueError(f"gather_contacts: this rank's buffer holds {compact.numel() // words} records but the largest rank has {m}; "
                         "every rank's buffer must hold the largest count (the payload is padded to it)")
    words = m * words
    if out is None or out.numel() < world * words:
        out = torch.empty(world * max(words, 1), dtype=torch.int32, device=compact.device)
    if words:
        dist.all_gather_into_tensor(out[: world * words], compact[:words])
    return out[: world * words], counts


def unpack_gathered(gathered, counts, words=REC_WORDS):
    """-> concatenated valid records (total x words) in rank order."""
    world = counts.numel()
    m = int(counts.max().item())
    g = gathered.view(world, m, words) if m else gathered.view(world, 0, words)
    return torch.cat([g[r, : int(counts[r].item())] for r in range(world)], dim=0)
