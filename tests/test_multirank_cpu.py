"""world_size-2 gloo test (CPU) of the multi-GPU host logic: round-robin frame sharding and the single
collective of the path (all_gather of 72-byte pose records)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pram_b200.runner import shard_frames, pack_pose_records, gather_pose_records
    ids = shard_frames(n_frames, rank, world)
    per_rank = (n_frames + world - 1) // world
    pad = per_rank - len(ids)
    fid = ids + [-1] * pad
    n = len(fid)
    # a "pose" that encodes the frame id, so the gathered table can be checked exactly
    qv = torch.tensor([[1.0, 0, 0, f * 1e-3] for f in fid], dtype=torch.float64)
    tv = torch.tensor([[f, 2.0 * f, -f] for f in fid], dtype=torch.float64)
    ni = torch.tensor([100 + f for f in fid], dtype=torch.int32)
    table = gather_pose_records(pack_pose_records(fid, qv, tv, ni))
    if rank == 0:
        q.put(table.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_pose_gather_world2():
    n_frames, world = 7, 2  # ragged on purpose
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    table = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert table.shape == (n_frames, 9)
    assert list(table[:, 0]) == list(range(n_frames))          # every frame exactly once, sorted
    assert all(table[:, 6] == 2.0 * table[:, 0]) and all(table[:, 8] == 100 + table[:, 0])


def test_shard_partition_properties():
    from pram_b200.runner import shard_frames
    for n, w in [(0, 1), (1, 4), (8, 8), (257, 8), (32, 3)]:
        parts = [shard_frames(n, r, w) for r in range(w)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
