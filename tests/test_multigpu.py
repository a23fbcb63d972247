"""2-GPU checks (skipped on a single-GPU box): NCCL view-sharded fusion equals the single-GPU
gather formulation, and sharded pair processing returns every pair exactly once."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import bsgpu
    from bsgpu import parallel
    from oracle import fusion_oracle as fo
    from tests import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = bsgpu.Context(rank)
    G = synth.field((30, 40, 110), seed=3, sigma=1.5)
    gv = []
    for i, t in enumerate([(0.0, 0.0, 0.0), (30.4, 1.2, -0.7), (61.1, -1.5, 1.1), (15.3, 2.2, 0.4)]):
        vol = synth.tile_from(G, (1, 2, int(t[0]) + 2), (24, 32, 40), 50 + i, noise=5.0)
        M = synth.translation(t)
        border, rng = fo.adjust_blending(M)
        gv.append(dict(src_to_world=M, vol_handle=ctx.volume_upload(vol), blend_border=border, blend_range=rng))
    bmin, bsz = (-2, -1, -1), (100, 34, 26)
    p = ctx.fuse_params("AVG_BLEND")
    mine = [gv[i] for i in parallel.partition_views(range(4), rank, world)]
    sharded = parallel.fuse_block_view_sharded(ctx, mine, bmin, bsz, p)
    whole = ctx.fuse_block(gv, bmin, bsz, p)
    # the same exchange behind the C ABI (bs_comm_init + bs_fuse_allreduce on the context's stream)
    parallel.comm_init_from_torch(ctx)
    native_sharded = parallel.fuse_block_view_sharded_native(ctx, mine, bmin, bsz, p)
    assert np.array_equal(native_sharded, sharded)
    ctx.comm_destroy()
    pairs = [synth.shifted_pair((32, 40, 48), s, seed=40 + i) for i, s in enumerate([(1, 2, 3), (-2, 0, 1), (3, -3, 2)])]
    idx = parallel.shard_items(list(range(3)), rank, world)
    local = [(i, ctx.pcm_pair(*pairs[i]).shift_int) for i in idx]
    allres = parallel.gather_results(local)
    if rank == 0:
        q.put((sharded, whole, allres))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_two_gpu_view_sharded_fusion_and_pair_sharding():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    c = mp.get_context("spawn")
    q = c.Queue()
    procs = [c.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded, whole, allres = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    err = np.abs(sharded - whole) / np.maximum(np.abs(whole), 1.0)
    assert err.max() < 1e-5
    assert sorted(allres) == [(0, (1, 2, 3)), (1, (-2, 0, 1)), (2, (3, -3, 2))]
