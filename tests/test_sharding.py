"""Multi-GPU path on CPU: world_size-2 gloo processes shard pictures (i mod G), reconstruct their
shard (with the oracle standing in for the GPU -- this is test infrastructure), gather, and the
merged result must equal the single-process result."""
import os
import socket

import numpy as np
import pytest


def test_shard_indices_partition():
    from minivideo_b200 import shard
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 4, 8):
            parts = [shard.shard_indices(n, world, r) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if n else np.array([], np.int64)
            assert np.array_equal(allidx, np.arange(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)


def test_merge_inverts_sharding():
    from minivideo_b200 import shard
    data = np.arange(11 * 3).reshape(11, 3)
    per = [data[shard.shard_indices(11, 4, r)] for r in range(4)]
    assert np.array_equal(shard.merge_results(11, 4, per), data)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from minivideo_b200 import shard, synth
    from oracle import cpu
    _, soa = synth.generate(5, want_stream=False, width_mbs=6, height_mbs=4, profile_idc=100, transform8x8=1, seed=21)
    idx = shard.shard_indices(soa.n_pics, world, rank)
    mine, _ = cpu.reconstruct(shard.take_pictures(soa, idx))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    if rank == 0:
        q.put(shard.merge_results(soa.n_pics, world, gathered))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_reconstruction_matches_single_process():
    import torch.multiprocessing as mp
    from minivideo_b200 import synth
    from oracle import cpu
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, soa = synth.generate(5, want_stream=False, width_mbs=6, height_mbs=4, profile_idc=100, transform8x8=1, seed=21)
    want, _ = cpu.reconstruct(soa)
    assert np.array_equal(merged, want)
