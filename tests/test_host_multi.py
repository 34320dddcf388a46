"""CPU tests of the host-side multi-GPU logic: slab partition, ownership, and the unique-id
hand-off over a world_size-2 gloo process group."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import oracle_lib as ol


def test_balanced_slabs_cover_domain_and_balance_particles():
    from mpm_b200 import slabs as sl

    for N, g in ((256, 2), (322, 2), (408, 4), (512, 8)):
        slabs = sl.balanced_slabs(N, g, 0.1, 0.9)
        assert slabs[0][0] == 0 and slabs[-1][1] == N
        assert all(slabs[i][1] == slabs[i + 1][0] for i in range(g - 1))
        x = ol.dense_block_positions(200_000, seed=7)[:, 0]
        counts = np.bincount(sl.owner(x, N, slabs), minlength=g)
        assert counts.sum() == len(x)
        assert counts.max() / counts.mean() < 1.03  # balanced to a few planes' worth


def test_owner_matches_oracle_base_node():
    from mpm_b200 import slabs as sl

    N = 60  # dx_inv is not exactly N here
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.02, 1.02, (5000, 3)).astype(np.float32)
    p = ol.new_particles(x)
    keys = ol.cell_keys(p, 1e-4, N)
    assert np.array_equal(sl.base_node_x(x[:, 0], N), (keys // (N * N)).astype(np.int32))
    slabs = sl.balanced_slabs(N, 3)
    own = sl.owner(x[:, 0], N, slabs)
    for r, (b, e) in enumerate(slabs):
        bx = keys[own == r] // (N * N)
        assert ((bx >= b) & (bx < e)).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from mpm_b200 import slabs as sl

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = sl.share_unique_id(dist, rank, lambda: bytes(range(128)))
    N = 64
    slabs = sl.balanced_slabs(N, world, 0.1, 0.9)
    x = ol.dense_block_positions(50_000, seed=11)
    mine = np.where(sl.owner(x[:, 0], N, slabs) == rank)[0]
    q.put((rank, uid, slabs[rank], len(mine), int(mine.sum() % 1000003)))
    dist.barrier()
    dist.destroy_process_group()


def test_unique_id_handoff_and_partition_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res[0][1] == res[1][1] == bytes(range(128))          # same id on both ranks
    assert res[0][2][1] == res[1][2][0]                         # contiguous slabs
    assert res[0][3] + res[1][3] == 50_000                      # every particle has exactly one owner
