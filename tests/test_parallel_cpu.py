"""N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding helpers and the
view-sharded weight-sum all-reduce (the one exchange step of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bsgpu
from bsgpu import parallel
from oracle import fusion_oracle as fo
from tests import synth


def test_shard_helpers():
    assert [parallel.shard_range(112, r, 8) for r in range(8)] == [(14 * r, 14 * r + 14) for r in range(8)]
    assert [parallel.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sum((parallel.shard_items(list(range(113)), r, 8) for r in range(8)), []) == list(range(113))
    assert [parallel.slab_for_rank(2048, 128, r, 8) for r in range(8)] == [(256 * r, 256 * r + 256) for r in range(8)]
    assert [parallel.slab_for_rank(300, 128, r, 2) for r in range(2)] == [(0, 256), (256, 300)]
    assert parallel.partition_views([5, 1, 3, 2], 0, 2) == [1, 3] and parallel.partition_views([5, 1, 3, 2], 1, 2) == [2, 5]


def _scene():
    G = synth.field((30, 40, 110), seed=3, sigma=1.5)
    views = []
    for i, t in enumerate([(0.0, 0.0, 0.0), (30.4, 1.2, -0.7), (61.1, -1.5, 1.1), (15.3, 2.2, 0.4)]):
        vol = synth.tile_from(G, (1, 2, int(t[0]) + 2), (24, 32, 40), 50 + i, noise=5.0)
        M = synth.translation(t)
        border, rng = fo.adjust_blending(M)
        views.append(fo.View(vol, M, border, rng))
    return views


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    views = _scene()
    mine = [views[i] for i in parallel.partition_views(range(len(views)), rank, world)]
    bmin, bsz = (-2, -1, -1), (100, 34, 26)
    swi, sw = fo.accumulate_block(mine, bmin, bsz, fo.AVG_BLEND)
    tswi, tsw = torch.from_numpy(swi), torch.from_numpy(sw)
    parallel.allreduce_partials(tswi, tsw)
    out = np.zeros_like(swi)
    np.divide(tswi.numpy(), tsw.numpy(), out=out, where=tsw.numpy() > 0)
    # pairs: every rank processes its shard, results are gathered on all ranks
    pairs = [(i, i + 1) for i in range(7)]
    local = [(p, rank) for p in parallel.shard_items(pairs, rank, world)]
    allres = parallel.gather_results(local)
    if rank == 0:
        q.put((out, allres))
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharded_allreduce_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, allres = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = fo.fuse_block(_scene(), (-2, -1, -1), (100, 34, 26), fo.AVG_BLEND)
    err = np.abs(out - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < 1e-5           # float re-association only
    assert [p for p, _ in allres] == [(i, i + 1) for i in range(7)]
    assert [r for _, r in allres] == [0, 0, 0, 0, 1, 1, 1]


def _cmd_worker(rank, world, port, tmp, q):
    """The command bodies sharded over two gloo ranks (oracle-backed fake context): pairs[r::w] + all_gather_object for
    `stitching`, contiguous z-slabs + barrier for `affine-fusion` with a re-read pyramid."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bsgpu import commands
    from tests.fake_ctx import FakeContext

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    ctx = FakeContext()
    xml = os.path.join(tmp, "dataset.xml")
    raw = commands.stitching(xml, ctx, downsampling=(1, 1, 1), shard=(rank, world), allgather=allgather)
    dist.barrier()
    out = os.path.join(tmp, "fused.n5")
    commands.affine_fusion(out, ctx, "AVG_BLEND", block_scale=(1, 1, 1), shard=(rank, world), barrier=dist.barrier)
    dist.barrier()
    if rank == 0:
        q.put(([None if r is None else (r.pair, np.rint(r.transform[:, 3]).tolist()) for r in raw], ctx.calls["pcm"]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_commands_gloo_world2(tmp_path):
    from bsgpu import commands, n5 as bn5, spimdata
    from tests.test_commands_cpu import _dataset
    xml, vols, tiles, planted, nominal = _dataset(tmp_path)
    out = str(tmp_path / "fused.n5")
    commands.create_fusion_container(xml, out, block_size=(16, 16, 16), downsamplings=[(2, 2, 1)], compression="raw")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cmd_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    raw, npcm0 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(raw) == 3 and npcm0 == 2                       # rank 0 correlated pairs 0 and 2, rank 1 pair 1
    assert raw[0][1] == list(planted[1]) and raw[1][1] == list(planted[2])
    assert len(spimdata.SpimData2.load(xml).stitching_results()) == 3
    views = []
    for t in tiles:
        M = synth.translation(t["translation_xyz"])
        border, rng = fo.adjust_blending(M)
        views.append(fo.View(vols[t["setup"]], M, border, rng))
    ext = (nominal + 48, nominal + 48, 48)
    want = fo.fuse_block(views, (0, 0, 0), ext, fo.AVG_BLEND)
    st, _ = bn5.read_fusion_container(out)
    assert np.array_equal(st.read_volume("ch0tp0/s0"), want)              # two ranks, three z-slabs, no seam
    assert np.array_equal(st.read_volume("ch0tp0/s1"), fo.downsample2x(want, (2, 2, 1)))
