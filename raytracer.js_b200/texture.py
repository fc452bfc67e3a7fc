"""Textures (src/texture/texture.ts, texture_solid.ts, texture_image.ts).

The nearest-texel lookup itself (texture_image.ts:40-63) runs on the GPU; the host objects only carry
the data.  ImageTexture here is array-backed; the reference's decoder (`new Image()` + canvas,
texture_image.ts:76-136) is browser-only, ImageTexture.from_file stands in for it on the host with Pillow
(SURVEY.md 8f N3): same RGB8 texels, same flip semantics."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .color import Color


class TextureError(Exception):
    pass


class Texture:
    def get_size(self) -> Optional[Tuple[int, int]]:
        return None


class SolidTexture(Texture):
    def __init__(self, color: Color):
        self.color = color

    def get_color(self, _u: float = 0.0, _v: float = 0.0) -> Color:
        return self.color


class ImageTexture(Texture):
    """`pixels`: uint8 array [height, width, 3] as load_image would have decoded it (row 0 first), or
    None for an image that is not (yet) loaded, which answers with the fallback colour."""

    def __init__(self, pixels: Optional[np.ndarray], fallback_color: Color, horizontal_flip=False,
                 vertical_flip=False):
        self.fallback_color = fallback_color
        self.image_data: Optional[np.ndarray] = None
        self.width = self.height = 0
        if pixels is not None:
            px = np.ascontiguousarray(pixels, dtype=np.uint8)
            if px.ndim != 3 or px.shape[2] != 3:
                raise TextureError("pixels must be uint8 [height, width, 3]")
            if horizontal_flip:
                px = px[:, ::-1]
            if vertical_flip:
                px = px[::-1]
            self.image_data = np.ascontiguousarray(px)
            self.height, self.width = px.shape[:2]

    def get_size(self):
        return (self.width, self.height) if self.image_data is not None else None

    @classmethod
    def from_file(cls, image_url: str, fallback_color: Color, horizontal_flip=False, vertical_flip=False) -> "ImageTexture":
        """`new ImageTexture(image_url, fallback_color, hflip, vflip)` (texture_image.ts:28-38) with the decode
        done on the host: the image is drawn as the canvas would hold it (8-bit RGB, row 0 first; the alpha
        channel is dropped like `image_data[index+3]` is never read), then flipped as load_image does
        (:103-113).  A file that cannot be decoded leaves the texture unloaded: lookups answer the fallback
        colour, as while the reference's loading promise is pending or rejected."""
        px = cls._decode_native(image_url)
        if px is None:  # a format the library's decoder does not read (JPEG ...): Pillow, where it is installed
            try:
                from PIL import Image
                with Image.open(image_url) as im:
                    px = np.asarray(im.convert("RGB"), dtype=np.uint8)
            except Exception:
                return cls(None, fallback_color)
        return cls(px, fallback_color, horizontal_flip, vertical_flip)

    @staticmethod
    def _decode_native(image_url: str):
        """rt_image_decode (include/rt_b200.h): the library's own host decoder - PNG, BMP, binary PPM - the one the
        N-API shim hands to the TypeScript adapter; None when the file is not one of those (or cannot be read)."""
        import ctypes as C
        try:
            from . import _native as N
            lib = N.load()
            with open(image_url, "rb") as fh:
                data = fh.read()
        except Exception:
            return None
        w, h, rgb = C.c_uint32(), C.c_uint32(), C.POINTER(C.c_uint8)()
        if lib.rt_image_decode(data, len(data), C.byref(w), C.byref(h), C.byref(rgb)) != N.RT_OK:
            return None
        try:
            return np.ctypeslib.as_array(rgb, shape=(h.value, w.value, 3)).copy()
        finally:
            lib.rt_image_free(rgb)
