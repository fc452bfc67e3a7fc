"""rt_image_decode: the library's host decoder behind ImageTexture - the browser-only decode of
src/texture/texture_image.ts:76-136 for hosts without a DOM.  Checked against Pillow's decode of the same bytes for every
PNG colour type and filter the encoder produces, BMP both ways up, binary PPM; anything else is refused (the texture
then answers its fallback colour)."""
import ctypes as C
import io

import numpy as np
import pytest

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N

Image = pytest.importorskip("PIL.Image")


def decode(data: bytes):
    lib = N.load()
    w, h, rgb = C.c_uint32(), C.c_uint32(), C.POINTER(C.c_uint8)()
    st = lib.rt_image_decode(data, len(data), C.byref(w), C.byref(h), C.byref(rgb))
    if st != N.RT_OK:
        return st, (lib.rt_last_error(None) or b"").decode()
    try:
        return N.RT_OK, np.ctypeslib.as_array(rgb, shape=(h.value, w.value, 3)).copy()
    finally:
        lib.rt_image_free(rgb)


def encoded(arr, mode, fmt, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr, mode).save(buf, fmt, **kw)
    return buf.getvalue()


def pillow_rgb(data: bytes):
    with Image.open(io.BytesIO(data)) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


@pytest.mark.parametrize("mode,channels", [("RGB", 3), ("RGBA", 4), ("L", 1), ("LA", 2)])
@pytest.mark.parametrize("smooth", [False, True])
def test_png_colour_types_and_filters(mode, channels, smooth):
    rng = np.random.default_rng(7)
    h, w = 37, 53  # odd sizes
    if smooth:  # gradients make the encoder choose the Sub / Up / Average / Paeth filters, noise mostly None
        y, x = np.mgrid[0:h, 0:w]
        a = np.stack([(x * 3 + y * (k + 1)) % 256 for k in range(channels)], -1).astype(np.uint8)
    else:
        a = rng.integers(0, 256, (h, w, channels), dtype=np.uint8)
    a = a[..., 0] if channels == 1 else a
    for kw in ({}, {"optimize": True}, {"compress_level": 0}):
        data = encoded(a, mode, "PNG", **kw)
        st, got = decode(data)
        assert st == N.RT_OK, got
        np.testing.assert_array_equal(got, pillow_rgb(data))


def test_png_palette():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (40, 64, 3), dtype=np.uint8)
    buf = io.BytesIO()
    Image.fromarray(a, "RGB").quantize(colors=200).save(buf, "PNG")
    st, got = decode(buf.getvalue())
    assert st == N.RT_OK, got
    np.testing.assert_array_equal(got, pillow_rgb(buf.getvalue()))


def test_bmp_and_ppm():
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, (19, 31, 3), dtype=np.uint8)  # 31 * 3 bytes: rows are padded to 4 in a BMP
    for fmt in ("BMP", "PPM"):
        data = encoded(a, "RGB", fmt)
        st, got = decode(data)
        assert st == N.RT_OK, got
        np.testing.assert_array_equal(got, a)
    # a top-down BMP (negative height) by hand
    data = bytearray(encoded(a, "RGB", "BMP"))
    hdr_h = int.from_bytes(data[22:26], "little", signed=True)
    data[22:26] = (-hdr_h).to_bytes(4, "little", signed=True)
    st, got = decode(bytes(data))
    assert st == N.RT_OK
    np.testing.assert_array_equal(got, a[::-1])


def test_undecodable_input_is_refused():
    for data in (b"not an image", b"\x89PNG\r\n\x1a\nbroken", encoded(np.zeros((4, 4, 3), np.uint8), "RGB", "JPEG"), b"P6 4 4 65535 "):
        st, msg = decode(data)
        assert st == N.RT_ERR_UNSUPPORTED and msg
    png16 = encoded((np.arange(64, dtype=np.uint16) * 900).reshape(8, 8), "I;16", "PNG")
    assert decode(png16)[0] == N.RT_ERR_UNSUPPORTED  # 16 bits per channel: not decoded here
    truncated = encoded(np.zeros((16, 16, 3), np.uint8), "RGB", "PNG")[:-40]
    assert decode(truncated)[0] == N.RT_ERR_UNSUPPORTED


def test_image_texture_uses_the_library_decoder(tmp_path):
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (12, 20, 4), dtype=np.uint8)
    Image.fromarray(a, "RGBA").save(tmp_path / "t.png")
    assert np.array_equal(rt.ImageTexture._decode_native(str(tmp_path / "t.png")), a[..., :3])
    t = rt.ImageTexture.from_file(str(tmp_path / "t.png"), rt.Color(1, 0, 1, 1), vertical_flip=True)
    assert t.get_size() == (20, 12) and np.array_equal(t.image_data, a[::-1, :, :3])
    Image.fromarray(a[..., :3], "RGB").save(tmp_path / "t.jpg")  # not the library's: Pillow takes over
    assert rt.ImageTexture._decode_native(str(tmp_path / "t.jpg")) is None
    assert rt.ImageTexture.from_file(str(tmp_path / "t.jpg"), rt.Color(1, 0, 1, 1)).get_size() == (20, 12)
