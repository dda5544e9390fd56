"""Render groups behind the C ABI (vt_group_*, csrc/vt_group.inl): partition + combination of the accumulators + replicated
edits, through ctypes. On one GPU the group holds two contexts on the same device (peer-memory exchange: NCCL refuses duplicate
devices); with >= 2 GPUs the same checks run over NCCL, in one process (ncclCommInitAll) and with one process per GPU
(ncclCommInitRank, the NCCL id handed round through a file). TILES must equal a single-context render bit for bit, SAMPLES
within 1e-5 of the oracle's average (fp summation order)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import voxeltoy_b200 as vt
from oracle import vto
from tests import util
from voxeltoy_b200 import group as G

pytestmark = pytest.mark.gpu

W, H = 200, 136            # 4 x 3 tiles, both edges clipped


def _scene():
    return util.make_frame(util.mixed_scene(), W, H, bounces=3, theta=130, phi=25)


def _contexts(devices, d):
    cs = [vt.Context(dev) for dev in devices]
    for c in cs:
        util.upload(c, d)
    return cs


def _single(d, n_passes):
    c = vt.Context(0)
    util.upload(c, d)
    c.render(0, n_passes)
    out = c.read_average()
    c.close()
    return out


def _check_group(devices, exchange=None):
    d = _scene()
    want = _single(d, 3)
    cs = _contexts(devices, d)
    g = G.DeviceGroup.adopt(cs, G.PART_TILES)
    try:
        if exchange is not None:
            g.set_exchange(exchange)
        assert g.size() == len(devices)
        g.render(0, 3)
        got = g.read_average()
        assert util.same_bits(got, want).all(), "tile partition over %r differs from one context" % (devices,)
        n_tiles = ((W + 63) // 64) * ((H + 63) // 64)
        if g.exchange() == G.DeviceGroup.EXCHANGE_NCCL:          # the root receives only the other ranks' tiles
            per_rank = (n_tiles + len(devices) - 1) // len(devices)
            assert g.exchange_bytes() == per_rank * 64 * 64 * 16 * (len(devices) - 1)
        # asynchronous combination: the exchange works on a snapshot, rendering continues meanwhile
        g.begin_combine()
        g.render(3, 2)
        snap = g.end_combine()
        assert util.same_bits(snap, want).all()
        assert util.same_bits(g.read_average(), _single(d, 5)).all()
    finally:
        g.close()
        for c in cs:
            c.close()
    # sample partition: rank r renders sampleCount = p * N + r; SUM over ranks / total
    n = len(devices)
    cs = _contexts(devices, d)
    g = G.DeviceGroup.adopt(cs, G.PART_SAMPLES)
    try:
        if exchange is not None:
            g.set_exchange(exchange)
        g.render(0, 2)
        got = g.read_average()
        s = vto.make_scene(d)
        ref = sum(vto.render_pass(s, k, want_hits=False)[0].astype(np.float64) for k in range(2 * n)) / (2 * n)
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-6, equal_nan=True)
        if g.exchange() == G.DeviceGroup.EXCHANGE_PEER:          # fixed summation order: rank 0 + rank 1 + ..., one division
            acc = None
            for c in cs:
                a = c.read_average()
                acc = a if acc is None else acc + a
            assert util.same_bits(got, acc / np.float32(2 * n)).all()
        # edits reach every replica and reset the accumulation (renderer/actions.cpp:20-52)
        grid = d["grid"].copy()
        g.pick(W * 0.5, H * 0.5)
        ri, rn = vto.pick(vto.make_scene(d), W * 0.5, H * 0.5, near_z=d["near_z"])
        g.add_voxel(0.0, 0.0)
        vto.add_voxel(vto.make_scene(d), grid, ri, rn, 0.0, 0.0)
        for c in cs:
            assert np.array_equal(c.get_selection()[0], ri)
            assert np.array_equal(c.read_volume(), grid)
            assert c.num_samples() == 0
    finally:
        g.close()
        for c in cs:
            c.close()


def test_group_two_contexts_on_one_gpu_peer_exchange():
    """ctypes, 2 contexts on device 0: the group falls back to its own peer-memory kernel (NCCL refuses duplicate devices)."""
    d = _scene()
    cs = _contexts([0, 0], d)
    g = G.DeviceGroup.adopt(cs, G.PART_TILES)
    assert g.exchange() == G.DeviceGroup.EXCHANGE_PEER
    with pytest.raises(vt.VtError):
        g.set_exchange(G.DeviceGroup.EXCHANGE_NCCL)
    g.close()
    for c in cs:
        c.close()
    _check_group([0, 0])
    _check_group([0, 0, 0])                   # 12 tiles over 3 ranks


def test_group_nccl_in_process():
    if vt.load().vt_device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    assert G.DeviceGroup._lib().vt_nccl_version() > 0
    _check_group([0, 1])                      # NCCL by default on distinct devices
    _check_group([0, 1], exchange=G.DeviceGroup.EXCHANGE_PEER)


def _rank_main(rank, world, id_path, mode, q):
    import time
    import voxeltoy_b200 as vt
    from voxeltoy_b200 import group as G
    from tests import util
    d = util.make_frame(util.mixed_scene(), W, H, bounces=3, theta=130, phi=25)
    c = vt.Context(rank)
    util.upload(c, d)
    if rank == 0:
        with open(id_path + ".tmp", "wb") as f:
            f.write(G.DeviceGroup.unique_id())
        os.rename(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.05)
    g = G.DeviceGroup.join(c, open(id_path, "rb").read(), rank, world, mode)
    g.render(0, 2)
    rec = np.array([0.25, 0.75, 0, 0, 2, 1, 0, 0], np.float32) if rank == 0 else np.zeros(8, np.float32)
    rec = g.broadcast(rec, root=0)
    g.begin_combine()                         # a collective: every rank takes part, only rank 0 receives the frame
    out = g.end_combine(want=(rank == 0))
    if rank == 0:
        q.put((out, rec))
    else:
        q.put((None, rec))
    g.close(); c.close()


@pytest.mark.parametrize("mode", [G.PART_TILES, G.PART_SAMPLES], ids=["tiles", "samples"])
def test_group_nccl_one_process_per_gpu(mode):
    if vt.load().vt_device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    id_path = os.path.join(tempfile.mkdtemp(), "nccl_id")
    ps = [ctx.Process(target=_rank_main, args=(r, 2, id_path, mode, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(60)
    frame = [r[0] for r in res if r[0] is not None][0]
    for _, rec in res:
        assert rec[1] == 0.75 and rec[4] == 2.0          # the edit record reached every rank
    d = _scene()
    if mode == G.PART_TILES:
        assert util.same_bits(frame, _single(d, 2)).all()
    else:
        s = vto.make_scene(d)
        ref = sum(vto.render_pass(s, k, want_hits=False)[0].astype(np.float64) for k in range(4)) / 4
        assert np.allclose(frame, ref, rtol=1e-5, atol=1e-6, equal_nan=True)


def test_cpp_renderer_group_demo_matches_single_renderer(tmp_path):
    """tools/vt_group_demo.cpp -- a C++ program without Python over RendererGroup (host/vt_host.h): BASELINE config 2 (reduced to
    320x180) on two replicas that share GPU 0, tile partition; its frame must equal the single Renderer's bit for bit."""
    from voxeltoy_b200 import host, scenes
    demo = os.path.join(os.path.dirname(vt.LIB_PATH), "vt_group_demo")
    assert os.path.exists(demo), "build voxeltoy_b200 first (python -m voxeltoy_b200.build)"
    vox = host.plain_path(util.SCENE_FALL)
    env = str(tmp_path / "env.pfm")
    host.write_pfm(env, scenes.synthetic_env(256, 128))
    out = str(tmp_path / "frame.pfm")
    n = vt.load().vt_device_count()
    devices = "0,1" if n >= 2 else "0,0"
    log = subprocess.check_output([demo, "--vox", vox, "--env", env, "--devices", devices, "--mode", "tiles", "--width", "320", "--height", "180",
                                   "--bounces", "4", "--passes", "6", "--steps", "1", "--out", out], timeout=300).decode()
    assert '"msamples_per_s"' in log
    r = host.Renderer(); r.initialize("", 0)
    r.resizeFrame(320, 180)
    r.loadVoxFile(util.SCENE_FALL)
    r.setRenderSettings(maxBounces=4, backgroundImage=env)
    cam = r.camera()
    cam.setLensModel(host.CLM_THIN_LENS)
    cam.controller().orbitAroundTarget(np.radians(120.0), np.radians(30.0))
    cam.setFStop(2.8)
    r.resetRender()
    r.context().set_selection([-1, -1, -1, 0], [1, 0, 0, 0])
    r.requestAction(0.5, 0.5, 0.0, 0.0, host.PA_SELECT_FOCAL_POINT)
    r.renderPasses(6)
    r.resetRender()
    r.renderPasses(6)
    want = r.readAverage()[..., :3]
    r.close()
    got = host.load_image(out)
    got = np.asarray(got).reshape(180, 320, 3)
    # the PFM holds rows bottom-up like the accumulator; saveImage-style flips are not applied by writePFM
    assert util.same_bits(got, want).all() or util.same_bits(got[::-1], want).all()
