"""Differential tests of the texture decoders (load_png / load_hdr in ygl_sceneio.cpp) against the reference's
load_texture (yocto_sceneio.cpp:1796, i.e. stb_image): seeded random PNG files written by the encoder below - every
colour type and bit depth (1 / 2 / 4 / 8 / 16), palette with short PLTE and tRNS, colour keys, Adam7 interlacing, all five
scanline filters, split IDAT, ancillary chunks - and Radiance files (run-length and flat scanlines, odd widths) must give
bit-identical texels; broken files must be refused by both. Host-only."""
import json
import os
import random
import struct
import zlib

import numpy as np
import pytest

from ygl_b200 import lib


def _chunk(tag, body):
    return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)


def _filtered_rows(samples, depth, rng):
    """samples: (h, w, channels) ints < 2**depth -> the filtered scanlines of one (sub)image"""
    h, w, channels = samples.shape
    if depth == 16:
        rows = samples.astype(">u2").tobytes()
        stride = w * channels * 2
    elif depth == 8:
        rows = samples.astype(np.uint8).tobytes()
        stride = w * channels
    else:
        per_row = []
        for j in range(h):
            bits = "".join(format(int(v), "0%db" % depth) for v in samples[j].reshape(-1))
            bits += "0" * (-len(bits) % 8)
            per_row.append(int(bits, 2).to_bytes(len(bits) // 8, "big"))
        rows = b"".join(per_row)
        stride = len(per_row[0])
    bpp = max(1, channels * depth // 8)
    out = bytearray()
    prev = bytes(stride)
    for j in range(h):
        row = rows[j * stride:(j + 1) * stride]
        kind = rng.randrange(5)
        line = bytearray(stride)
        for i in range(stride):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            if kind == 0:
                pred = 0
            elif kind == 1:
                pred = a
            elif kind == 2:
                pred = b
            elif kind == 3:
                pred = (a + b) >> 1
            else:
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            line[i] = (row[i] - pred) & 255
        out += bytes([kind]) + line
        prev = row
    return bytes(out)


def make_png(rng, w, h, ctype, depth, interlace):
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    top = 1 << depth
    extra = b""
    if ctype == 3:
        entries = rng.randint(1, min(256, top))
        extra += _chunk(b"PLTE", bytes(rng.randrange(256) for _ in range(entries * 3)))
        if rng.random() < 0.6:
            extra += _chunk(b"tRNS", bytes(rng.randrange(256) for _ in range(rng.randint(1, entries))))
        samples = np.array([[[rng.randrange(entries)] for _ in range(w)] for _ in range(h)])   # (stb reads uninitialised
        # memory for an index past the PLTE chunk; we return transparent black)
    else:
        few = [rng.randrange(top) for _ in range(3)]          # few distinct values so that a colour key hits
        samples = np.array([[[rng.choice(few) if rng.random() < 0.5 else rng.randrange(top) for _ in range(channels)]
                             for _ in range(w)] for _ in range(h)])
        if ctype in (0, 2) and rng.random() < 0.6:
            key = [few[0]] * channels if rng.random() < 0.7 else [rng.randrange(top) for _ in range(channels)]
            extra += _chunk(b"tRNS", b"".join(struct.pack(">H", v) for v in key))
    if rng.random() < 0.3:
        extra = _chunk(b"gAMA", struct.pack(">I", 45455)) + extra
    if interlace:
        data = b""
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            part = samples[y0::dy, x0::dx]
            if part.shape[0] and part.shape[1]:
                data += _filtered_rows(part, depth, rng)
    else:
        data = _filtered_rows(samples, depth, rng)
    packed = zlib.compress(data, rng.choice([0, 1, 6, 9]))
    cuts = sorted(rng.sample(range(1, len(packed)), min(len(packed) - 1, rng.choice([0, 0, 1, 2]))))
    idat = b"".join(_chunk(b"IDAT", packed[a:b]) for a, b in zip([0] + cuts, cuts + [len(packed)]))
    if rng.random() < 0.2:
        idat += _chunk(b"tEXt", b"Comment\0fuzz")
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace)) + extra + idat
            + _chunk(b"IEND", b""))


def _both(ref, tmp_path, name, data):
    """texels from our loader and the reference's, or None where the file was refused"""
    os.makedirs(tmp_path / "textures", exist_ok=True)
    (tmp_path / "textures" / name).write_bytes(data)
    scene = tmp_path / (name + ".json")
    scene.write_text(json.dumps({"asset": {"version": "4.2"}, "textures": [{"name": "t", "uri": "textures/" + name}]}))
    try:
        ours = lib.load_scene(scene).textures[0]
    except lib.YglError:
        ours = None
    try:
        theirs = ref.load_scene(scene).textures[0]
    except RuntimeError:
        theirs = None
    return ours, theirs


def _same(ours, theirs):
    return (ours["pixels"].shape == theirs["pixels"].shape and ours["pixels"].dtype == theirs["pixels"].dtype
            and ours["pixels"].tobytes() == theirs["pixels"].tobytes() and ours["linear"] == theirs["linear"])


VALID = [(0, d) for d in (1, 2, 4, 8, 16)] + [(3, d) for d in (1, 2, 4, 8)] + [(c, d) for c in (2, 4, 6) for d in (8, 16)]


@pytest.mark.parametrize("seed", [7, 8])
def test_random_png_files_decode_like_stb_image(ref, seed, tmp_path):
    rng = random.Random(seed)
    for k in range(120):
        ctype, depth = VALID[k % len(VALID)]
        w, h = rng.choice([1, 2, 3, 5, 8, 9, 17]), rng.choice([1, 2, 4, 7, 8, 13])
        data = make_png(rng, w, h, ctype, depth, interlace=k % 2)
        ours, theirs = _both(ref, tmp_path, f"p{k}.png", data)
        assert ours is not None and theirs is not None, (k, ctype, depth, w, h, ours is None, theirs is None)
        assert _same(ours, theirs), (k, ctype, depth, w, h, k % 2)


def test_broken_png_files_are_refused_by_both(ref, tmp_path):
    rng = random.Random(3)
    good = make_png(rng, 5, 4, 2, 8, 0)
    ihdr = lambda w, h, depth, ctype, il=0, comp=0, flt=0: b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, comp, flt, il))
    body = good[33:]
    cases = {
        "signature": b"\x89PNX" + good[4:],
        "truncated": good[:len(good) // 2],
        "zero width": ihdr(0, 4, 8, 2) + body,
        "depth 3": ihdr(5, 4, 3, 2) + body,
        "palette 16 bit": ihdr(5, 4, 16, 3) + body,
        "colour type 1": ihdr(5, 4, 8, 1) + body,
        "interlace 2": ihdr(5, 4, 8, 2, il=2) + body,
        "filter method": ihdr(5, 4, 8, 2, flt=1) + body,
        "palette without PLTE": ihdr(5, 4, 8, 3) + body,
        "too few pixels": ihdr(5, 9, 8, 2) + body,
        "critical chunk": good[:33] + _chunk(b"ABCD", b"x") + body,
        "tRNS with alpha": ihdr(5, 4, 8, 6) + _chunk(b"tRNS", b"\0\0") + body,
        "scanline filter 5": ihdr(2, 1, 8, 0) + _chunk(b"IDAT", zlib.compress(b"\x05\x01\x02")) + _chunk(b"IEND", b""),
        "no IDAT": ihdr(2, 1, 8, 0) + _chunk(b"IEND", b""),
    }
    for what, data in cases.items():
        ours, theirs = _both(ref, tmp_path, "bad.png", data)
        assert ours is None and theirs is None, (what, ours is None, theirs is None)
    ours, theirs = _both(ref, tmp_path, "good.png", good)
    assert _same(ours, theirs)


def make_hdr(rng, w, h, rle, header=b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n"):
    rgbe = np.array([[[rng.randrange(256), rng.randrange(256), rng.randrange(256), rng.choice([0, 100, 128, 129, 140, 160, 255])]
                      for _ in range(w)] for _ in range(h)], np.uint8)
    out = header + b"-Y %d +X %d\n" % (h, w)
    for j in range(h):
        if not rle:
            out += rgbe[j].tobytes()
            continue
        out += bytes([2, 2, w >> 8, w & 255])
        for c in range(4):
            row, i = rgbe[j, :, c], 0
            while i < w:
                run = 1
                while i + run < w and run < 127 and row[i + run] == row[i]:
                    run += 1
                if run >= 3 or rng.random() < 0.2:
                    out += bytes([128 + run, row[i]])
                    i += run
                else:
                    n = min(w - i, rng.randint(1, 9))
                    out += bytes([n]) + row[i:i + n].tobytes()
                    i += n
    return out


def test_random_radiance_files_decode_like_stb_image(ref, tmp_path):
    rng = random.Random(11)
    for k in range(40):
        w = rng.choice([1, 3, 7, 8, 9, 40, 129])       # stb run-length decodes only 8 <= width < 32768, else reads flat
        h = rng.choice([1, 2, 5])
        data = make_hdr(rng, w, h, rle=(w >= 8 and k % 2 == 0), header=rng.choice(
            [b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n", b"#?RGBE\nEXPOSURE=1.0\nFORMAT=32-bit_rle_rgbe\n\n"]))
        ours, theirs = _both(ref, tmp_path, f"h{k}.hdr", data)
        assert ours is not None and theirs is not None, (k, w, h)
        assert _same(ours, theirs), (k, w, h)
    for what, data in {"signature": b"#?NOPE\nFORMAT=32-bit_rle_rgbe\n\n-Y 1 +X 1\n\0\0\0\0",
                       "format": b"#?RADIANCE\n\n-Y 1 +X 1\n\0\0\0\0",
                       "orientation": b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n+Y 1 +X 1\n\0\0\0\0"}.items():
        ours, theirs = _both(ref, tmp_path, "bad.hdr", data)
        assert ours is None and theirs is None, (what, ours is None, theirs is None)


def _jpeg_cases(rng, count):
    """(bytes, description) of JPEG files written by Pillow and OpenCV: grey / RGB / CMYK, 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 /
    4:1:1 chroma layouts, baseline and progressive, default and optimised Huffman tables, restart intervals, sizes that
    are not multiples of the MCU"""
    import io
    PIL = pytest.importorskip("PIL.Image")
    cv2 = pytest.importorskip("cv2")
    for k in range(count):
        w, h = rng.choice([1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 64, 70]), rng.choice([1, 3, 8, 9, 16, 17, 25, 40])
        kind = rng.random()
        smooth = np.add.outer(np.linspace(0, 255, h), np.linspace(0, 120, w))[..., None] % 256
        noise = np.array([[[rng.randrange(256) for _ in range(4)] for _ in range(w)] for _ in range(h)], np.float64)
        pixels = (smooth * 0.6 + noise * rng.choice([0.05, 0.4, 1.0])).clip(0, 255).astype(np.uint8)
        quality = rng.choice([5, 30, 60, 75, 90, 100])
        if kind < 0.6:
            mode = rng.choice(["L", "RGB", "RGB", "RGB", "CMYK"])
            channels = {"L": 1, "RGB": 3, "CMYK": 4}[mode]
            image = PIL.fromarray(pixels[..., 0] if channels == 1 else pixels[..., :channels], mode)
            options = dict(quality=quality, progressive=rng.random() < 0.4, optimize=rng.random() < 0.4)
            if mode == "RGB":
                options["subsampling"] = rng.choice([0, 1, 2])
            if rng.random() < 0.3:
                options["restart_marker_blocks"] = rng.choice([1, 2, 5])
            out = io.BytesIO()
            try:
                image.save(out, "JPEG", **options)
            except TypeError:
                options.pop("restart_marker_blocks", None)
                image.save(out, "JPEG", **options)
            yield out.getvalue(), f"pillow {mode} {w}x{h} {options}"
        else:
            params = [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_PROGRESSIVE, int(rng.random() < 0.4),
                      cv2.IMWRITE_JPEG_OPTIMIZE, int(rng.random() < 0.4), cv2.IMWRITE_JPEG_RST_INTERVAL, rng.choice([0, 0, 1, 3, 7])]
            if hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
                params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, rng.choice([0x411111, 0x221111, 0x211111, 0x121111, 0x111111])]
            ok, data = cv2.imencode(".jpg", pixels[..., 0] if rng.random() < 0.2 else pixels[..., :3], params)
            assert ok
            yield data.tobytes(), f"opencv {w}x{h} {params}"


@pytest.mark.parametrize("seed", [21, 22])
def test_jpeg_files_decode_like_stb_image(ref, seed, tmp_path):
    """Every texel of 120 JPEG files per seed against stb_image: its integer IDCT, its triangle-filter chroma upsampling
    and its fixed-point YCbCr conversion are what 'the same image' means for this format."""
    rng = random.Random(seed)
    for k, (data, what) in enumerate(_jpeg_cases(rng, 120)):
        ours, theirs = _both(ref, tmp_path, f"j{k}.jpg", data)
        assert ours is not None and theirs is not None, (k, what, ours is None, theirs is None)
        assert _same(ours, theirs), (k, what)
    for what, data in {"no SOI": b"\xff\xd9", "truncated": data[:len(data) // 3], "text": b"not a jpeg at all"}.items():
        ours, theirs = _both(ref, tmp_path, "bad.jpg", data)
        assert ours is None and theirs is None, (what, ours is None, theirs is None)


def test_exr_files_decode_like_tinyexr(ref, tmp_path):
    """OpenEXR scanline files written by OpenCV - half and float channels, 1 / 3 / 4 channels, compression none / RLE / ZIPS
    / ZIP / PIZ, sizes around the 16-line ZIP and 32-line PIZ blocks, values that exercise both PIZ lifting steps - against
    tinyexr's LoadEXR: every float bit-identical (half -> float is
    exact, including subnormals, infinities and NaN payloads). And the reference's own sky1.exr (2048 x 1024, ZIP, half)."""
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    import scene_data
    rng = random.Random(31)
    done = 0
    for k in range(90):
        w, h = rng.choice([1, 2, 5, 16, 17, 40, 130]), rng.choice([1, 3, 15, 16, 17, 33, 50, 70])
        channels = rng.choice([1, 3, 4])
        pixels = np.array([[[rng.choice([0.0, 1.0, -2.5, 6.1e-5, 1e-7, 65504.0, 1e6, float("inf"), rng.uniform(-10, 10), rng.random()])
                             for _ in range(channels)] for _ in range(w)] for _ in range(h)], np.float32)
        if channels == 1:
            pixels = pixels[..., 0]
        params = [cv2.IMWRITE_EXR_TYPE, rng.choice([cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_TYPE_FLOAT]),
                  cv2.IMWRITE_EXR_COMPRESSION, rng.choice([cv2.IMWRITE_EXR_COMPRESSION_NO, cv2.IMWRITE_EXR_COMPRESSION_RLE,
                                                           cv2.IMWRITE_EXR_COMPRESSION_ZIPS, cv2.IMWRITE_EXR_COMPRESSION_ZIP,
                                                           cv2.IMWRITE_EXR_COMPRESSION_PIZ, cv2.IMWRITE_EXR_COMPRESSION_PIZ])]
        with np.errstate(over="ignore"):
            ok, data = cv2.imencode(".exr", pixels, params)
        if not ok:
            pytest.skip("this OpenCV build cannot write OpenEXR")
        ours, theirs = _both(ref, tmp_path, f"e{k}.exr", data.tobytes())
        assert ours is not None and theirs is not None, (k, w, h, channels, params, ours is None, theirs is None)
        assert _same(ours, theirs), (k, w, h, channels, params)
        done += 1
    assert done == 90
    for what, params in {"pxr24": [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PXR24],
                         "b44": [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_B44]}.items():
        ok, data = cv2.imencode(".exr", np.ones((20, 20, 3), np.float32), params)
        ours, theirs = _both(ref, tmp_path, what + ".exr", data.tobytes())
        assert ours is None and theirs is None, what       # codecs tinyexr does not carry either
    sky = os.path.join("/root/reference/tests/_data/textures/sky1.exr")
    if os.path.exists(sky):
        ours, theirs = _both(ref, tmp_path, "sky1.exr", open(sky, "rb").read())
        assert ours is not None and theirs is not None and _same(ours, theirs)
        assert ours["pixels"].shape == (1024, 2048, 4)
