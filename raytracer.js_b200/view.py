"""View / ToneMapper / Screen (src/view/view.ts:24-41, src/view/tone_mapping.ts:20-79, src/view/screen.ts,
src/view/screen_canvas.ts): the step right after trace_frame().  `GpuView.draw_ebuffer()` is the drop-in for
`View.draw_ebuffer()`: exposure statistics, the tone mapper's dynamic range and the range compression to
8-bit pixels run on the GPU (rt_present / rt_render_present in include/rt_b200.h); the Screen receives the
RGBA8 image a CanvasScreen's ImageData would hold.  There is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _native as N
from .exposure_buffer import ExposureBuffer


class ToneMapper:
    """ToneMapper (src/view/tone_mapping.ts:20-22): the dynamic range is computed on the device; a subclass
    only describes itself through tone_desc()."""

    def tone_desc(self) -> N.Tone:
        raise TypeError(f"unsupported ToneMapper subclass {type(self).__name__}")


class ToneMapper_Identity(ToneMapper):  # :24-32
    instance: "ToneMapper_Identity"

    def tone_desc(self) -> N.Tone:
        return N.Tone(N.RT_TONE_IDENTITY, 0, 0.0, 0.0)


ToneMapper_Identity.instance = ToneMapper_Identity()


class ToneMapper_DRLimited(ToneMapper):  # :34-46
    KIND = -1

    def __init__(self, dynamic_range: int, min_dynamic: float, max_dynamic: float):
        self.dynamic_range = int(dynamic_range)
        self.dynamic_coef = 1 << self.dynamic_range
        self.min_dynamic = float(min_dynamic)
        self.max_dynamic = float(max_dynamic)

    def tone_desc(self) -> N.Tone:
        if self.KIND < 0:
            return super().tone_desc()
        return N.Tone(self.KIND, self.dynamic_range, self.min_dynamic, self.max_dynamic)


class ToneMapper_StdDevAroundMean(ToneMapper_DRLimited):  # :48-63
    KIND = N.RT_TONE_STDDEV


class ToneMapper_AbsDevAroundMean(ToneMapper_DRLimited):  # :65-79
    KIND = N.RT_TONE_ABSDEV


class Screen:
    """A buffered screen (src/view/screen.ts:24-45 / CanvasScreen with flags.buffer_pixels): `image` is the
    ImageData a canvas would hold, uint8 [height, width, 4]."""

    def __init__(self, width: int, height: int):
        self.width, self.height = int(width), int(height)
        self.image = np.zeros((self.height, self.width, 4), np.uint8)
        self.flushes = 0

    @property
    def dynamic_range(self) -> int:  # screen_canvas.ts:105-107
        return 8

    def set_pixel_i(self, i: int, pixel) -> None:  # screen_canvas.ts:45-56 with convert_color :101-103
        flat = self.image.reshape(-1, 4)
        for k in range(3):
            v = pixel[k] if k < len(pixel) else float("nan")
            v = min(max(v, 0.0), 1.0) * 255 if v == v else 0.0
            flat[i, k] = int(v)
        flat[i, 3] = 0xff

    def flush(self) -> None:
        self.flushes += 1


class GpuView:
    """View (src/view/view.ts:24-41) with the work of draw_ebuffer() on the GPU.

    * `GpuView(ebuffer, screen, tone_mapper, raytracer)` + `draw_ebuffer()`: the host ExposureBuffer is sent to
      the device and presented (rt_present);
    * `trace_and_draw(n_frames)`: trace_frame() + draw_ebuffer() with the ExposureBuffer resident on the
      device - only the 8-bit image leaves the GPU (rt_render_present); `download_exposure()` fetches the
      float frame when it is wanted on the host.
    """

    def __init__(self, ebuffer: ExposureBuffer, screen: Screen, tone_mapper: ToneMapper, raytracer):
        self.ebuffer, self.screen, self.tone_mapper, self.raytracer = ebuffer, screen, tone_mapper, raytracer
        self.last_stats: Optional[dict] = None

    def _check(self) -> None:
        if (self.screen.width, self.screen.height) != (self.ebuffer.width, self.ebuffer.height):
            raise IndexError("x or y out of bounds")

    def draw_ebuffer(self) -> Tuple[float, float]:
        self._check()
        rt, tone, st = self.raytracer, self.tone_mapper.tone_desc(), N.ExposureStats()
        N.check(rt.ctx, rt.lib.rt_present(rt.ctx, self.ebuffer.pixels.ctypes.data, self.ebuffer.width, self.ebuffer.height,
                                          C.byref(tone), self.screen.image.ctypes.data, C.byref(st)))
        self.last_stats = st.as_dict()
        return st.drange_low, st.drange_high

    def trace_and_draw(self, n_frames: int = 1) -> Tuple[float, float]:
        self._check()
        rt, eb = self.raytracer, self.ebuffer
        from .raytracer import camera_desc
        p = rt.params(n_frames=n_frames, frame_first=eb.current_frame)
        cd = camera_desc(rt._camera, rt.reference_extents)
        tone, st = self.tone_mapper.tone_desc(), N.ExposureStats()
        N.check(rt.ctx, rt.lib.rt_render_present(rt.ctx, C.byref(cd), C.byref(p), 0, C.byref(tone), self.screen.image.ctypes.data,
                                                 C.byref(st), None))
        for _ in range(n_frames - 1):
            eb.next_frame()
        self.last_stats = st.as_dict()
        return st.drange_low, st.drange_high

    def download_exposure(self) -> None:
        rt, eb = self.raytracer, self.ebuffer
        N.check(rt.ctx, rt.lib.rt_exposure_download(rt.ctx, eb.pixels.ctypes.data, eb.width, eb.height))


View = GpuView  # the drop-in name
