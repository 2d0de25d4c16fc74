"""Display pass (display.wgsl:29-86) -- CPU side: the oracle against hand-derived
known answers and an independent numpy transliteration, the LutManager mirror of lut_manager.rs, the PNG writer."""
import os
import struct
import zlib

import numpy as np
import pytest

import slime_mold_b200 as sm


def numpy_display(trail, lut, tw, th):
    """Independent transliteration of display.wgsl:44-86 with numpy float32 arithmetic."""
    f = np.float32
    H, W = trail.shape
    sim_w, sim_h, tex_w, tex_h = f(W), f(H), f(tw), f(th)
    sim_aspect, tex_aspect = sim_w / sim_h, tex_w / tex_h
    off_x = off_y = f(0)
    if tex_aspect > sim_aspect:
        scale = tex_h / sim_h
        off_x = (tex_w - sim_w * scale) * f(0.5)
    else:
        scale = tex_w / sim_w
        off_y = (tex_h - sim_h * scale) * f(0.5)
    fx = (np.arange(tw, dtype=np.float32) - off_x) / scale
    fy = (np.arange(th, dtype=np.float32) - off_y) / scale
    out = np.zeros((th, tw, 4), np.uint8)
    out[..., 3] = 255
    inx = (fx >= 0) & (fx < sim_w)
    iny = (fy >= 0) & (fy < sim_h)
    xs = fx[inx].astype(np.int32)
    ys = fy[iny].astype(np.int32)
    sub = trail[np.ix_(ys, xs)]
    inten = np.minimum(np.maximum(np.where(np.isnan(sub), f(0), sub), f(0)), f(1))
    li = (inten * f(255.0)).astype(np.uint32)
    block = np.stack([lut[li], lut[li + 256], lut[li + 512], np.full(li.shape, 255, np.uint8)], axis=-1)
    out[np.ix_(np.nonzero(iny)[0], np.nonzero(inx)[0])] = block
    return out


def test_unorm_store_returns_the_lut_byte(oracle):
    """f32(b)/255 stored as rgba8unorm gives b back for every byte value: the kernel may copy LUT bytes."""
    lut = np.concatenate([np.arange(256), np.arange(256)[::-1], (np.arange(256) * 7) % 256]).astype(np.uint8)
    trail = (np.arange(256, dtype=np.float32) / np.float32(255.0)).reshape(1, 256)
    trail = np.nextafter(trail, np.float32(2.0)).astype(np.float32)      # just above i/255: index i after truncation
    out = oracle.display(trail, lut, 256, 1)
    idx = (np.minimum(trail[0], np.float32(1.0)) * np.float32(255.0)).astype(np.uint32)
    assert np.array_equal(out[0, :, 0], lut[idx])
    assert np.array_equal(out[0, :, 1], lut[idx + 256])
    assert np.array_equal(out[0, :, 2], lut[idx + 512])
    assert (out[..., 3] == 255).all()
    assert set(idx.tolist()) == set(range(256))


def test_known_answers(oracle):
    lut = sm.LutManager().load_lut("gray").combined()
    # same aspect, same size: texel (x, y) shows cell (x, y); intensity clamps; NaN shows as 0
    t = np.array([[0.0, 0.5, 1.0, 2.0], [-1.0, np.nan, 0.25, 0.999]], np.float32)
    o = oracle.display(t, lut, 4, 2)
    assert o[..., 0].tolist() == [[0, 127, 255, 255], [0, 0, 63, 254]]
    assert np.array_equal(o[..., 0], o[..., 1]) and np.array_equal(o[..., 0], o[..., 2]) and (o[..., 3] == 255).all()
    # frame wider than the map: fit height, black bars left and right (display.wgsl:61-64, 83-85)
    t = np.ones((2, 2), np.float32)
    o = oracle.display(t, lut, 8, 2)
    assert o[0, :, 0].tolist() == [0, 0, 0, 255, 255, 0, 0, 0]
    # frame taller than the map: fit width, bars above and below; 2x magnification repeats cells
    t = np.array([[0.0, 1.0]], np.float32)
    o = oracle.display(t, lut, 4, 6)
    assert o[:, :, 0].tolist() == [[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 255, 255], [0, 0, 255, 255], [0, 0, 0, 0], [0, 0, 0, 0]]


@pytest.mark.parametrize("W,H,tw,th", [(192, 108, 192, 108), (192, 108, 320, 200), (100, 37, 63, 200), (64, 64, 1, 1),
                                      (33, 77, 500, 40), (1920, 1080, 1600, 900)])
def test_oracle_equals_numpy_transliteration(oracle, W, H, tw, th):
    rng = np.random.default_rng(W * 131 + th)
    t = (rng.random((H, W), dtype=np.float32) * np.float32(1.4) - np.float32(0.2)).astype(np.float32)
    t[rng.integers(0, H, 5), rng.integers(0, W, 5)] = np.nan
    lut = rng.integers(0, 256, 768).astype(np.uint8)
    assert np.array_equal(oracle.display(t, lut, tw, th), numpy_display(t, lut, tw, th))


def test_lut_manager_mirror(tmp_path):
    """lut_manager.rs:149-186: sorted names, 768-byte files split into red / green / blue, reverse()."""
    buf = np.arange(768, dtype=np.uint32).astype(np.uint8)
    (tmp_path / "B_second.lut").write_bytes(buf.tobytes())
    (tmp_path / "A_first.lut").write_bytes(buf[::-1].tobytes())
    (tmp_path / "broken.lut").write_bytes(b"123")
    (tmp_path / "notes.txt").write_text("not a lut")
    lm = sm.LutManager(str(tmp_path))
    assert lm.get_available_luts() == ["A_first", "B_second", "broken", "gray", "gray_r"]
    d = lm.load_lut("B_second")
    assert d.name == "B_second" and np.array_equal(d.red, buf[:256]) and np.array_equal(d.green, buf[256:512]) and np.array_equal(d.blue, buf[512:])
    assert np.array_equal(d.combined(), buf)
    d.reverse()
    assert np.array_equal(d.red, buf[:256][::-1]) and np.array_equal(d.blue, buf[512:][::-1])
    with pytest.raises(ValueError):
        lm.load_lut("broken")              # io::ErrorKind::InvalidData
    with pytest.raises(FileNotFoundError):
        lm.load_lut("missing")             # io::ErrorKind::NotFound
    assert np.array_equal(lm.load_lut("gray_r").red, np.arange(256)[::-1])


def test_png_writer_round_trip(tmp_path):
    rgba = np.random.default_rng(3).integers(0, 256, (17, 23, 4)).astype(np.uint8)
    p = tmp_path / "f.png"
    sm.write_png(str(p), rgba)
    raw = p.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, {}
    while pos < len(raw):
        n, tag = struct.unpack(">I4s", raw[pos:pos + 8])
        data = raw[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
        chunks[tag] = data
        pos += 12 + n
    w, h, depth, ctype = struct.unpack(">IIBB", chunks[b"IHDR"][:10])
    assert (w, h, depth, ctype) == (23, 17, 8, 6)
    lines = np.frombuffer(zlib.decompress(chunks[b"IDAT"]), np.uint8).reshape(17, 1 + 23 * 4)
    assert (lines[:, 0] == 0).all() and np.array_equal(lines[:, 1:].reshape(17, 23, 4), rgba)
