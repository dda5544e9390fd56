"""CPU tests of the product's C++ host classes (voxeltoy_b200/host, flat C wrappers) against the oracle's independent
numpy restatement of the reference's host code (oracle/scene.py): loaders, mesh transform, emissive pruning, CDFs.
These run without a GPU (they never create a Renderer context)."""
import struct

import numpy as np
import pytest

from oracle import scene as oscene
from oracle import vto
from tests import util


@pytest.fixture(scope="module")
def host():
    from voxeltoy_b200 import build
    build.build()
    import voxeltoy_b200 as vt
    return vt.host


def test_magicavoxel_loader_matches_oracle(host):
    got = host.load_vox(util.SCENE_FALL)
    ref = util.scene_fall_volume()
    assert tuple(got["res"]) == tuple(ref["res"])
    assert np.array_equal(got["grid"], ref["grid"])
    assert np.array_equal(got["materials"], ref["materials"])
    assert got["emissive"].size == 0


def _write_vox(path, size, voxels, palette=None):
    xyzi = struct.pack("<i", len(voxels)) + b"".join(struct.pack("<4B", *v) for v in voxels)
    chunks = b"SIZE" + struct.pack("<ii", 12, 0) + struct.pack("<iii", *size)
    chunks += b"XYZI" + struct.pack("<ii", len(xyzi), 0) + xyzi
    if palette is not None:
        chunks += b"RGBA" + struct.pack("<ii", 1024, 0) + palette.tobytes()
    with open(path, "wb") as f:
        f.write(b"VOX " + struct.pack("<i", 150) + b"MAIN" + struct.pack("<ii", 0, len(chunks)) + chunks)


def test_magicavoxel_default_palette_and_errors(host, tmp_path):
    rng = np.random.RandomState(0)
    vox = [(int(x), int(y), int(z), int(c)) for x, y, z, c in zip(rng.randint(0, 7, 60), rng.randint(0, 5, 60),
                                                                     rng.randint(0, 9, 60), rng.randint(1, 256, 60))]
    p = str(tmp_path / "a.vox")
    _write_vox(p, (7, 5, 9), vox)
    got, ref = host.load_vox(p), oscene.load_vox(p)
    assert tuple(got["res"]) == (7, 9, 5) == tuple(ref["res"])
    assert np.array_equal(got["grid"], ref["grid"]) and np.array_equal(got["materials"], ref["materials"])
    pal = rng.randint(0, 256, size=(256, 4)).astype(np.uint8)
    _write_vox(p, (7, 5, 9), vox, pal)
    got, ref = host.load_vox(p), oscene.load_vox(p)
    assert np.array_equal(got["grid"], ref["grid"]) and np.array_equal(got["materials"], ref["materials"])
    with open(p, "wb") as f:
        f.write(b"NOPE" + b"\0" * 32)
    with pytest.raises(IOError):
        host.load_vox(p)
    with pytest.raises(IOError):
        host.load_vox(str(tmp_path / "missing.vox"))


def test_obj_loader_and_mesh_transform_match_oracle(host, tmp_path):
    v, i = host.load_obj(util.BUNNY)
    rv, ri = oscene.load_obj(util.BUNNY)
    assert np.array_equal(v, rv) and np.array_equal(i, ri)
    for res in [(64, 64, 64), (512, 512, 512), (48, 32, 40)]:
        bmin, bmax = oscene.mesh_bounds(rv)
        assert np.array_equal(host.compute_mesh_transform(bmin, bmax, res), oscene.mesh_transform(bmin, bmax, res))
    # polygons (fan), negative indices, v/vt/vn corners, groups
    p = str(tmp_path / "q.obj")
    open(p, "w").write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1\nvn 0 0 1\nvt 0 0\nf 1 2 3 4\ng top\nf -1/1/1 1//1 2/1\nf 5 3 4\n")
    v, i = host.load_obj(p)
    rv, ri = oscene.load_obj(p)
    assert i.size == 12 and np.array_equal(i[:6], [0, 1, 2, 0, 2, 3])
    assert np.array_equal(v[i], rv[ri])          # same triangles


def test_prune_interior_emissive_matches_oracle(host):
    vol = util.mixed_scene(32, seed=9)
    from voxeltoy_b200 import scenes
    em = scenes.emissive_list(vol["grid"], vol["materials"])
    assert em.size > 10
    got = host.prune_interior_emissive(vol["grid"], vol["res"], em)
    ref = oscene.prune_interior_emissive(vol["grid"], vol["res"], em)
    assert np.array_equal(got, ref) and got.size < em.size


def test_calculate_cdf_matches_oracle(host, tmp_path):
    from voxeltoy_b200 import scenes
    # 1000x500, 1500x750 and 3200x1600 reduce by non-integer factors (area-weighted box filter)
    for size in [(256, 128), (1024, 512), (64, 48), (1000, 500), (1500, 750), (3200, 1600), (700, 1100)]:
        rgb = scenes.synthetic_env(*size)
        got = host.calculate_cdf(rgb)
        ref = oscene.build_env(rgb)
        assert got["cdf_u"].shape == ref["cdf_u"].shape
        assert np.array_equal(got["cdf_u"], ref["cdf_u"]) and np.array_equal(got["cdf_v"], ref["cdf_v"])
        assert got["integral"] == ref["integral"]
        assert np.all(np.diff(got["cdf_v"]) >= 0) and got["cdf_v"][-1] == 1.0
    # integer factors: the general filter IS the plain box mean
    rgb = scenes.synthetic_env(1024, 512)
    small = vto.resize_box(rgb, 512, 256)
    assert np.array_equal(small, rgb.reshape(256, 2, 512, 2, 3).astype(np.float64).mean(axis=(1, 3)).astype(np.float32))


def test_image_file_readers(host, tmp_path):
    from voxeltoy_b200 import scenes
    p = str(tmp_path / "e.pfm")
    rgb = scenes.synthetic_env(64, 32)
    host.write_pfm(p, rgb)
    assert np.array_equal(host.load_image(p), rgb)
    h = str(tmp_path / "e.hdr")
    px = np.array([[[1.0, 0.5, 0.25], [0, 0, 0]], [[100.0, 50.0, 3.0], [0.001, 0.002, 0.003]]], np.float32)
    with open(h, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 2\n")
        for row in px:
            for c in row:
                m = float(c.max())
                if m < 1e-32:
                    f.write(bytes([0, 0, 0, 0]))
                else:
                    e = int(np.floor(np.log2(m))) + 1
                    f.write(bytes([int(c[0] / 2.0 ** e * 256), int(c[1] / 2.0 ** e * 256), int(c[2] / 2.0 ** e * 256), e + 128]))
    got = host.load_image(h)
    assert np.allclose(got, px, rtol=0.02, atol=1e-4)
    with pytest.raises(IOError):
        host.load_image(str(tmp_path / "nope.hdr"))


def _decode_png(path):
    """minimal PNG reader for the test: checks signature and chunk CRCs, inflates IDAT with zlib, expects filter 0 rows"""
    import struct
    import zlib
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(b):
        n, typ = struct.unpack(">I4s", b[pos:pos + 8])
        data = b[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0] == (zlib.crc32(typ + data) & 0xFFFFFFFF)
        chunks.append((typ, data)); pos += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    w, h, depth, ctype, comp, flt, lace = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, flt, lace) == (8, 6, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(h, 1 + 4 * w)
    assert not raw[:, 0].any()
    return raw[:, 1:].reshape(h, w, 4)


def test_png_writer_round_trip(host, tmp_path):
    rng = np.random.default_rng(3)
    for (h, w) in ((1, 1), (7, 5), (300, 257)):            # the last one spans several 65535-byte stored blocks
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        p = str(tmp_path / ("t%dx%d.png" % (w, h)))
        host.write_png(p, img)
        assert np.array_equal(_decode_png(p), img)


def test_magicavoxel_corrupt_chunk_sizes_are_rejected(host, tmp_path):
    """chunk sizes come from the file. Negative sizes (the cursor would stand still or move backwards: an endless
    loop), and chunks that end beyond MAIN or beyond the file, must fail with 'corrupt chunk' instead of hanging."""
    import signal

    def vox(chunks, main_children=None):
        return b"VOX " + struct.pack("<i", 150) + b"MAIN" + struct.pack("<ii", 0, len(chunks) if main_children is None else main_children) + chunks

    size = b"SIZE" + struct.pack("<ii", 12, 0) + struct.pack("<iii", 2, 2, 2)
    cases = {
        "negative content size": vox(size + b"PACK" + struct.pack("<ii", -12, 0) + b"\0" * 8),
        "negative children size": vox(size + b"PACK" + struct.pack("<ii", 0, -24)),
        "chunk beyond MAIN": vox(size + b"XYZI" + struct.pack("<ii", 4000, 0) + struct.pack("<i", 0)),
        "MAIN with negative size": b"VOX " + struct.pack("<i", 150) + b"MAIN" + struct.pack("<ii", -4, 100) + size,
    }
    old = signal.signal(signal.SIGALRM, lambda *_: (_ for _ in ()).throw(TimeoutError("load_vox hung")))
    try:
        for name, data in cases.items():
            p = str(tmp_path / "bad.vox")
            with open(p, "wb") as f:
                f.write(data)
            signal.alarm(10)
            with pytest.raises(IOError):
                host.load_vox(p)
            signal.alarm(0)
    finally:
        signal.alarm(0); signal.signal(signal.SIGALRM, old)
    # MAIN may announce more children bytes than the file holds (truncated download): what is there is still read
    p = str(tmp_path / "short.vox")
    with open(p, "wb") as f:
        f.write(vox(size + b"XYZI" + struct.pack("<ii", 8, 0) + struct.pack("<i", 1) + struct.pack("<4B", 1, 1, 1, 7), main_children=4096))
    got = host.load_vox(p)
    assert tuple(got["res"]) == (2, 2, 2) and (got["grid"] >= 0).sum() == 1


def test_magicavoxel_palette_rules_match_oracle(host, tmp_path):
    """New-build extension: palette index -> Lambert / Metal / Plastic record with emission and roughness; the
    record layouts are the reference's (material/material.h:15-33), the emissive list follows voxLoader.h:23-24."""
    rng = np.random.RandomState(4)
    vox = [(int(x), int(y), int(z), int(c)) for x, y, z, c in zip(rng.randint(0, 6, 80), rng.randint(0, 6, 80),
                                                                     rng.randint(0, 6, 80), rng.choice([3, 9, 17, 40, 200], 80))]
    p = str(tmp_path / "rules.vox")
    _write_vox(p, (6, 6, 6), vox)
    rules = [(9, host.MT_METAL, (0.0, 0.0, 0.0), 120.0), (17, host.MT_LAMBERT, (4.0, 3.0, 2.0), 0.0),
             (40, host.MT_PLASTIC, (0.5, 0.0, 0.0), 35.0), (99, host.MT_METAL, (1.0, 1.0, 1.0), 5.0)]      # 99: unused colour
    got, ref = host.load_vox(p, palette_rules=rules), oscene.load_vox(p, palette_rules=rules)
    assert np.array_equal(got["grid"], ref["grid"]) and np.array_equal(got["materials"], ref["materials"])
    assert np.array_equal(got["emissive"], ref["emissive"]) and got["emissive"].size > 0
    types = sorted(set(int(got["materials"][o]) for o in np.unique(got["grid"][got["grid"] >= 0])))
    assert types == [0, 1, 2]
    plain = host.load_vox(p)
    assert np.array_equal(plain["materials"], oscene.load_vox(p)["materials"]) and plain["emissive"].size == 0
