"""Multi-rank path on CPU: world_size-2 gloo all-gather of fixed-shape detections must
reproduce the single-rank result after undoing the round-robin sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clouds, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from det3d_b200.apis import dist as d3dist

    r, w, _ = d3dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    mine = d3dist.shard_indices(n_clouds, rank, world)
    # "detections" of cloud i: a deterministic function of i, fixed shape [D=5, F=10]
    packed = torch.stack([torch.full((5, 10), float(i)) + torch.arange(10.0) for i in mine])
    gathered = d3dist.all_gather_detections(packed)
    ordered = d3dist.interleave_rank_major(gathered, world)
    torch.save(ordered, os.path.join(out_dir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_rank(tmp_path):
    world, n_clouds = 2, 8
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_clouds, str(tmp_path)), nprocs=world, join=True)
    expect = torch.stack([torch.full((5, 10), float(i)) + torch.arange(10.0) for i in range(n_clouds)])
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), "r%d.pt" % r))
        assert torch.equal(got, expect)


def test_shard_indices_cover_everything():
    from det3d_b200.apis.dist import shard_indices

    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_indices(32, r, world))
        assert seen == list(range(32))
