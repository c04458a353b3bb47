"""World-size-N check of the library's NCCL exchange, run under torchrun on a multi-GPU box:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/gather_2gpu.py
Every rank runs GJK+EPA on its own slice of one scene, gathers through (a) the device-pointer calls with begin / finish
split across two steps, (b) the host-buffer call, (c) the frame queue's gather mode, and compares the gathered list with the
single-process result of the whole scene (rank order = pair order, so the concatenation must be the full hit list)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bam3d_b200 as b3  # noqa: E402
from bam3d_b200 import multi, scenes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b3.Context(local)
    gjk = b3.GJK.new(ctx)
    comm = multi.Comm(local, rank, world)
    n = 20_000 + 37 * rank          # uneven slices
    starts = [sum(20_000 + 37 * r for r in range(k)) for k in range(world + 1)]
    FR = b3.CollisionStrategy.FullResolution
    full = scenes.mixed_pairs(starts[world], s=1.5, seed=0xB3D00A12)
    mine = scenes.mixed_pairs(n, s=1.5, seed=0xB3D00A12, start=starts[rank])
    shapes = b3.ShapeSet(ctx, mine["kind"], mine["params"], mine["xform"])
    contacts, st = gjk.intersection_batch(FR, shapes, mine["pairs"])
    fshapes = b3.ShapeSet(ctx, full["kind"], full["params"], full["xform"])
    fc, fst = gjk.intersection_batch(FR, fshapes, full["pairs"])
    fhits = np.nonzero(fst == 0)[0]
    want = np.concatenate([fhits[:, None].astype(np.uint32), np.zeros((len(fhits), 1), np.uint32), fc[fhits].view(np.uint32)], axis=1)
    # (b) host buffers
    rec, counts = comm.gather(ctx, contacts, st, starts[rank], len(fst))
    per_rank = [int(((fst[starts[r]:starts[r + 1]]) == 0).sum()) for r in range(world)]
    assert counts.tolist() == per_rank, (counts, per_rank)
    off = np.concatenate([[0], np.cumsum(per_rank)])
    got = np.concatenate([rec[off[r]:off[r + 1]][np.argsort(rec[off[r]:off[r + 1], 0], kind="stable")] for r in range(world)])
    assert got.tobytes() == want.tobytes(), "host-buffer gather differs"
    # (a) device pointers, two steps in flight
    d_pairs = torch.from_numpy(mine["pairs"].astype(np.int32).reshape(-1)).cuda()
    bufs = []
    for k in range(2):
        bufs.append(dict(c=torch.empty(n * 8, dtype=torch.float32, device="cuda"), s=torch.empty(n, dtype=torch.uint8, device="cuda"),
                         compact=torch.empty(n * 10, dtype=torch.int32, device="cuda"), count=torch.zeros(1, dtype=torch.int64, device="cuda"),
                         g=torch.empty(len(fst) * 10, dtype=torch.int32, device="cuda")))
    import ctypes as C
    p = lambda t: C.c_void_p(t.data_ptr())
    tickets = []
    for k in range(2):
        b = bufs[k]
        gjk.intersection_dev(FR, shapes, p(d_pairs), n, p(b["c"]), p(b["s"]))
        gjk.compact_contacts_dev(p(b["c"]), p(b["s"]), n, starts[rank], p(b["compact"]), n, p(b["count"]))
        tickets.append(comm.begin(ctx, p(b["count"])))
    for k in range(2):
        cnt, total = comm.finish(ctx, tickets[k], p(bufs[k]["compact"]), p(bufs[k]["g"]), len(fst))
        assert cnt.tolist() == per_rank and total == len(fhits)
    comm.join(ctx)
    ctx.synchronize()
    for k in range(2):
        g = bufs[k]["g"].cpu().numpy().view(np.uint32).reshape(-1, 10)[: len(fhits)]
        got = np.concatenate([g[off[r]:off[r + 1]][np.argsort(g[off[r]:off[r + 1], 0], kind="stable")] for r in range(world)])
        assert got.tobytes() == want.tobytes(), "device gather differs"
    # (c) frame queue, gather mode
    fq = b3.FrameQueue(gjk, mine["kind"], mine["params"], mine["xform"], max_pairs=n, depth=2)
    hx = np.ascontiguousarray(mine["xform"], dtype=np.float32).reshape(-1, 16)
    hp = np.ascontiguousarray(mine["pairs"], dtype=np.uint32)
    def check(f, g, hs):
        fq.wait(f)
        cnt, total = fq.gathered(f)
        assert cnt.tolist() == per_rank and total == len(fhits) and np.array_equal(hs, st)
        got = np.concatenate([g[off[r]:off[r + 1]][np.argsort(g[off[r]:off[r + 1], 0], kind="stable")] for r in range(world)])
        assert got.tobytes() == want.tobytes(), "frame-queue gather differs"

    frames = []
    for k in range(4):
        if k >= 2:
            check(*frames[k - 2])   # depth 2: a frame's slot is reused two submissions later
        g, hs = b3.pinned_empty(ctx, (len(fst), 10), np.uint32), b3.pinned_empty(ctx, (n,), np.uint8)
        frames.append((fq.submit_contact_gather(comm, FR, hx, hp, starts[rank], g, hs), g, hs))
    check(*frames[2])
    check(*frames[3])
    fq.close()
    dist.barrier()
    comm.close()
    if rank == 0:
        print(f"gather ok on {world} ranks: {len(fhits)} contacts, per rank {per_rank}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
