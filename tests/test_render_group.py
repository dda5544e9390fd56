"""Multi-GPU host logic on CPU: 2 ranks over gloo (127.0.0.1). The per-rank accumulators a GPU context would hold are
produced here by the CPU oracle with the partition rules of voxeltoy_b200.group; the exchange (SUM-reduce + normalisation)
is the product's own code. Tile mode must reproduce the single-rank running average bit for bit, sample mode within 1e-5."""
import os
import socket

import numpy as np
import pytest

from tests import util

WORLD = 2
PASSES = 3


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, port, mode, q):
    import torch
    import torch.distributed as dist
    from oracle import vto
    from voxeltoy_b200 import group as G
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        d = util.make_frame(util.scene_fall_volume(), 160, 130, bounces=2, theta=120, phi=30)
        s = vto.make_scene(d)
        g = G.RenderGroup(mode, rank, WORLD)
        acc = np.zeros((130, 160, 4), np.float32)
        if mode == G.PART_TILES:
            mask = g.tile_mask(160, 130)
            for p in range(PASSES):                          # global sample index, own pixels only, running average
                smp = vto.render_pass(s, g.sample_index(p), n_threads=2, want_hits=False)[0]
                tmp = acc.copy(); vto.accumulate(tmp, smp, p); acc[mask] = tmp[mask]
        else:
            for p in range(PASSES):                          # own sample indices, running SUM
                acc += vto.render_pass(s, g.sample_index(p), n_threads=2, want_hits=False)[0]
        out = g.combine(torch.from_numpy(acc), PASSES)
        rec = g.broadcast_action([0.25, 0.75, 0.0, 0.0, 2.0, 1.0, 0, 0] if rank == 0 else np.zeros(8))
        assert rec[1] == 0.75 and rec[4] == 2.0
        if rank == 0:
            n_total = PASSES if mode == G.PART_TILES else PASSES * WORLD
            ref = vto.render_average(s, n_total, n_threads=2)
            got = out.numpy()
            if mode == G.PART_TILES:
                ok = bool(util.same_bits(got, ref).all())
            else:
                ok = bool(np.allclose(got, ref, rtol=1e-5, atol=1e-6))
            q.put(("ok" if ok else "mismatch: max |delta| %g" % float(np.nanmax(np.abs(got - ref)))))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", [1, 2], ids=["tiles", "samples"])
def test_two_ranks_gloo(mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, mode, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == "ok"


def test_partition_rules():
    from voxeltoy_b200 import group as G
    g0, g1 = G.RenderGroup(G.PART_TILES, 0, 2), G.RenderGroup(G.PART_TILES, 1, 2)
    m0, m1 = g0.tile_mask(200, 130), g1.tile_mask(200, 130)
    assert (m0 ^ m1).all() and m0[0, 0] and m1[0, 64] and m0[64, 0] and m1[64, 64]      # 4 tiles per row: tile (0,1) is tile 4
    assert [G.RenderGroup(G.PART_SAMPLES, r, 4).sample_index(p) for p in range(2) for r in range(4)] == list(range(8))
    with pytest.raises(ValueError):
        G.RenderGroup(G.PART_TILES, 2, 2)
    assert G.RenderGroup(G.PART_SAMPLES, 0, 1).mode == G.PART_NONE
