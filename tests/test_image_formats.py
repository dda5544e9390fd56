"""OpenEXR / PNG readers and the OpenEXR writer of the host classes (voxeltoy_b200/host/image_formats.cpp; the reference reads and
writes image files through OpenImageIO: renderer/image.cpp:28-59, renderer.cpp:1108-1140). Checked against files hand-built here
from the format specifications, and -- when OpenCV with OpenEXR support is importable -- against an independent codec both ways."""
import os
import struct
import zlib

import numpy as np
import pytest

from voxeltoy_b200 import host

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")


def _cv2():
    try:
        import cv2
        return cv2
    except Exception:
        return None


def _image(h=37, w=53, seed=1):
    rng = np.random.RandomState(seed)
    img = (rng.rand(h, w, 3).astype(np.float32) * 10) ** 2
    img[0, 0] = [0.0, 1e-8, 65504.0]; img[1, 1] = [np.inf, 1.0, 0.5]
    return img


def test_exr_round_trip_is_exact(tmp_path):
    img = _image()
    p = str(tmp_path / "a.exr")
    host.write_exr(p, img)
    back = host.load_image(p)
    assert back.shape == img.shape and np.array_equal(back.view(np.uint32), img.view(np.uint32))
    rgba = np.concatenate([img, np.full(img.shape[:2] + (1,), 0.25, np.float32)], axis=2)      # alpha is carried, RGB still found
    host.write_exr(p, rgba)
    assert np.array_equal(host.load_image(p), img)
    tall = _image(h=70, w=5, seed=2)                                                            # several 16-line ZIP blocks + a short last one
    host.write_exr(p, tall)
    assert np.array_equal(host.load_image(p), tall)


def _hand_made_exr(path, img, pixel_type, compression):
    """An OpenEXR scanline file straight from the file-layout document: channels B, G, R of HALF or FLOAT, NONE or ZIPS."""
    h, w, _ = img.shape
    def attr(name, typ, payload):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload
    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", pixel_type, 0, 0, 0, 0, 1, 1) for n in ("B", "G", "R")) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    hdr = struct.pack("<II", 20000630, 2) + attr("channels", "chlist", chlist) + attr("compression", "compression", bytes([compression])) \
        + attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0") \
        + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0)) \
        + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    chunks = []
    for y in range(h):
        line = b"".join((img[y, :, c].astype(np.float16 if pixel_type == 1 else np.float32)).tobytes() for c in (2, 1, 0))
        if compression == 2:                                   # ZIPS: even / odd byte split, delta predictor, zlib
            a = np.frombuffer(line, np.uint8)
            t = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)
            d = t.copy(); d[1:] = (t[1:] - t[:-1] + 128) & 0xff
            z = zlib.compress(d.astype(np.uint8).tobytes())
            line = z if len(z) < len(line) else line
        chunks.append(struct.pack("<ii", y, len(line)) + line)
    off = len(hdr) + 8 * h
    table = b""
    for c in chunks:
        table += struct.pack("<Q", off); off += len(c)
    open(path, "wb").write(hdr + table + b"".join(chunks))


@pytest.mark.parametrize("pixel_type,compression", [(1, 0), (2, 0), (1, 2), (2, 2)])
def test_exr_reader_on_hand_made_files(tmp_path, pixel_type, compression):
    img = _image(h=9, w=31, seed=3)
    img[1, 1, 0] = 2.0                                                         # no inf in the half case below
    p = str(tmp_path / "h.exr")
    _hand_made_exr(p, img, pixel_type, compression)
    want = img.astype(np.float16).astype(np.float32) if pixel_type == 1 else img
    with np.errstate(over="ignore"):
        got = host.load_image(p)
    assert np.array_equal(got, want)


def test_exr_against_an_independent_codec(tmp_path):
    cv2 = _cv2()
    if cv2 is None or not hasattr(cv2, "IMWRITE_EXR_COMPRESSION"):
        pytest.skip("OpenCV with OpenEXR support is not importable")
    img = _image()
    img[1, 1, 0] = 3.0
    p = str(tmp_path / "c.exr")
    host.write_exr(p, img)
    theirs = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    if theirs is None:
        pytest.skip("this OpenCV build cannot read OpenEXR")
    assert np.array_equal(theirs[..., ::-1], img)                              # their reader, our writer
    for typ, want in ((cv2.IMWRITE_EXR_TYPE_HALF, img.astype(np.float16).astype(np.float32)), (cv2.IMWRITE_EXR_TYPE_FLOAT, img)):
        for comp in (0, 1, 2, 3):                                              # NONE, RLE, ZIPS, ZIP
            assert cv2.imwrite(p, img[..., ::-1].copy(), [cv2.IMWRITE_EXR_TYPE, typ, cv2.IMWRITE_EXR_COMPRESSION, comp])
            assert np.array_equal(host.load_image(p), want), (typ, comp)      # our reader, their writer
        assert cv2.imwrite(p, img[..., ::-1].copy(), [cv2.IMWRITE_EXR_TYPE, typ, cv2.IMWRITE_EXR_COMPRESSION, 4])
        with pytest.raises(IOError):                                           # PIZ: refused, not misread
            host.load_image(p)


def _hand_made_png(path, arr, ctype, depth, palette=None, filters=(0, 1, 2, 3, 4)):
    """A PNG from RFC 2083: one IDAT, the five filter types cycled over the rows."""
    h, w = arr.shape[:2]
    comps = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    if depth == 16:
        rows = [arr[y].astype(">u2").tobytes() for y in range(h)]
    elif depth == 8:
        rows = [arr[y].astype(np.uint8).tobytes() for y in range(h)]
    else:
        rows = [np.packbits(np.unpackbits(arr[y].astype(np.uint8).reshape(-1, 1), axis=1)[:, 8 - depth:].reshape(-1)).tobytes() for y in range(h)]
    bpp = max(1, comps * depth // 8)
    out = b""; prev = bytes(len(rows[0]))
    for y, row in enumerate(rows):
        f = filters[y % len(filters)]
        cur = bytearray(len(row))
        for i in range(len(row)):
            a = row[i - bpp] if i >= bpp else 0; b = prev[i]; c = prev[i - bpp] if i >= bpp else 0
            if f == 0: pred = 0
            elif f == 1: pred = a
            elif f == 2: pred = b
            elif f == 3: pred = (a + b) >> 1
            else:
                pp = a + b - c; pa, pb, pc = abs(pp - a), abs(pp - b), abs(pp - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            cur[i] = (row[i] - pred) & 0xff
        out += bytes([f]) + bytes(cur); prev = row
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if palette is not None:
        png += chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    z = zlib.compress(out)
    png += chunk(b"IDAT", z[:len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b"")
    open(path, "wb").write(png)


def test_png_reader_on_hand_made_files(tmp_path):
    rng = np.random.RandomState(4)
    p = str(tmp_path / "a.png")
    h, w = 11, 19
    for ctype, comps in ((2, 3), (6, 4), (0, 1), (4, 2)):
        for depth in (8, 16):
            arr = rng.randint(0, 1 << depth, size=(h, w, comps))
            _hand_made_png(p, arr, ctype, depth)
            got = host.load_image(p)
            scale = np.float32(1.0) / np.float32((1 << depth) - 1)
            want = (arr[..., :3] if comps >= 3 else np.repeat(arr[..., :1], 3, axis=2)).astype(np.float32) * scale
            assert got.shape == (h, w, 3) and np.array_equal(got, want), (ctype, depth)
    for depth in (1, 2, 4):                                                   # packed grey and palette rows
        arr = rng.randint(0, 1 << depth, size=(h, w, 1))
        _hand_made_png(p, arr, 0, depth)
        assert np.array_equal(host.load_image(p), np.repeat(arr, 3, axis=2).astype(np.float32) * (np.float32(1.0) / np.float32((1 << depth) - 1)))
        pal = rng.randint(0, 256, size=(1 << depth, 3))
        _hand_made_png(p, arr, 3, depth, palette=pal)
        assert np.array_equal(host.load_image(p), pal[arr[..., 0]].astype(np.float32) * (np.float32(1.0) / np.float32(255.0)))


def test_png_writer_is_read_back(tmp_path):
    rng = np.random.RandomState(5)
    rgba = rng.randint(0, 256, size=(23, 40, 4)).astype(np.uint8)
    p = str(tmp_path / "w.png")
    host.write_png(p, rgba)
    assert np.array_equal(host.load_image(p), rgba[..., :3].astype(np.float32) * (np.float32(1.0) / np.float32(255.0)))
    cv2 = _cv2()
    if cv2 is not None:
        theirs = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        assert theirs is not None and np.array_equal(theirs[..., [2, 1, 0, 3]], rgba)


def test_hdr_writer_is_read_back(tmp_path):
    """Radiance RGBE keeps 8 bits of the largest component and a shared exponent: the round trip is exact for what the format holds,
    within 1 / 128 of the largest component otherwise; OpenCV's reader agrees with ours."""
    img = _image(h=21, w=33, seed=6)
    img[1, 1, 0] = 4.0
    p = str(tmp_path / "w.hdr")
    host.write_hdr(p, img)
    back = host.load_image(p)
    m = img.max(axis=2, keepdims=True)
    assert back.shape == img.shape and (np.abs(back - img) <= m / 128.0 + 1e-30).all()
    host.write_hdr(p, back)                                   # what RGBE can hold survives unchanged
    assert np.array_equal(host.load_image(p), back)
    cv2 = _cv2()
    if cv2 is not None:
        theirs = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        if theirs is not None:
            assert np.allclose(theirs[..., ::-1], back, rtol=1e-6, atol=0)


def test_reference_screenshot_if_present():
    """The reference's own PNG resources (this container only; the file does not travel)."""
    path = "/root/reference/resources/screenshot01.png"
    cv2 = _cv2()
    if not os.path.exists(path) or cv2 is None:
        pytest.skip("reference resources or OpenCV not available")
    theirs = cv2.imread(path, cv2.IMREAD_COLOR)
    got = host.load_image(path)
    assert got.shape == theirs.shape and np.array_equal(got, theirs[..., ::-1].astype(np.float32) * (np.float32(1.0) / np.float32(255.0)))


def test_bad_files_are_refused(tmp_path):
    p = str(tmp_path / "bad.exr")
    open(p, "wb").write(struct.pack("<II", 20000630, 2) + b"channels\0chlist\0" + struct.pack("<i", 1 << 30))
    with pytest.raises(IOError):
        host.load_image(p)
    open(p, "wb").write(b"\x89PNG\r\n\x1a\n" + b"\0" * 40)
    with pytest.raises(IOError):
        host.load_image(p)
    box = struct.pack("<iiii", 0, 0, 65535, 65535)                             # a 4 Gi-pixel data window in a 200-byte file
    open(p, "wb").write(struct.pack("<II", 20000630, 2) + b"channels\0chlist\0" + struct.pack("<i", 19) + b"R\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1) + b"\0"
                        + b"compression\0compression\0" + struct.pack("<i", 1) + b"\0" + b"dataWindow\0box2i\0" + struct.pack("<i", 16) + box + b"\0")
    with pytest.raises(IOError):
        host.load_image(p)
    open(p, "wb").write(struct.pack("<II", 20000630, 2 | 0x200))               # tiled
    with pytest.raises(IOError):
        host.load_image(p)
