"""Multi-GPU plumbing (one process per GPU, torch.distributed): pair lists shard contiguously across ranks with no
data-path collective; the one exchange step is the gather of the compacted contacts (bam3d_contact40 records, 10 x u32
each, tagged with the global pair index) so that every rank ends with the full contact list.
Device-agnostic on purpose: the same code runs over NCCL on GPUs and over gloo on CPU tensors in the tests."""
import torch
import torch.distributed as dist

REC_WORDS = 10  # bam3d_contact40 = 40 bytes


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

    compact: int32 tensor, this rank's records (>= count * REC_WORDS words); count: 1-element int64 tensor with the
    number of valid records. Returns (gathered, counts): `gathered` holds world * max_count records (rank r's records
    start at r * max_count), `counts` the per-rank record counts. Two collectives: counts (8 bytes per rank), then the
    payload padded to the largest count."""
    world = dist.get_world_size()
    counts = counts_buf if counts_buf is not None else torch.zeros(world, dtype=torch.int64, device=compact.device)
    dist.all_gather_into_tensor(counts, count)
    m = int(counts.max().item())
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
