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


def gather_contacts(compact, count, counts_buf=None, out=None):
    """All-gather variable-length compacted contact records.

    compact: int32 tensor, this rank's records (>= count * REC_WORDS words); count: 1-element int64 tensor with the
    number of valid records. Returns (gathered, counts): `gathered` holds world * max_count records (rank r's records
    start at r * max_count), `counts` the per-rank record counts. Two collectives: counts (8 bytes per rank), then the
    payload padded to the largest count."""
    world = dist.get_world_size()
    counts = counts_buf if counts_buf is not None else torch.zeros(world, dtype=torch.int64, device=compact.device)
    dist.all_gather_into_tensor(counts, count)
    m = int(counts.max().item())
    words = m * REC_WORDS
    if out is None or out.numel() < world * words:
        out = torch.empty(world * max(words, 1), dtype=torch.int32, device=compact.device)
    if words:
        dist.all_gather_into_tensor(out[: world * words], compact[:words])
    return out[: world * words], counts


def unpack_gathered(gathered, counts):
    """-> concatenated valid records (total x REC_WORDS) in rank order."""
    world = counts.numel()
    m = int(counts.max().item())
    g = gathered.view(world, m, REC_WORDS) if m else gathered.view(world, 0, REC_WORDS)
    return torch.cat([g[r, : int(counts[r].item())] for r in range(world)], dim=0)
