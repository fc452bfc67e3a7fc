"""ExposureBuffer (src/view/exposure_buffer.ts:26-91): the float32 running-mean frame store that
trace_frame() writes.  The GPU renderer updates `pixels` in bulk with the same blend
(col_weight = 1/(1+frame_count)); statistics and tone mapping run on the device (view.py)."""
from __future__ import annotations

import numpy as np


class ExposureBuffer:
    def __init__(self, width: int, height: int, max_exposure_frames: int = -1):
        self.width, self.height = int(width), int(height)
        self.pixels = np.zeros(self.width * self.height * 3, dtype=np.float32)
        self.max_exposure_frames = int(max_exposure_frames)
        self.reset_exposure()

    @property
    def current_frame(self) -> int:
        return self.frame_count

    def next_frame(self) -> bool:  # :53-60
        if self.max_exposure_frames > -1 and self.frame_count >= self.max_exposure_frames:
            return False
        self.frame_count += 1
        self.col_weight = 1 / (1 + self.frame_count)
        return True

    def reset_exposure(self) -> None:  # :62-66
        self.frame_count = 0
        self.col_weight = 1

    def check_bounds(self, x: int, y: int) -> None:  # :181-186
        if x < 0 or x >= self.width or y < 0 or y >= self.height:
            raise IndexError("x or y out of bounds")

    def set_color(self, x: int, y: int, pixel) -> None:  # :68-91
        self.check_bounds(x, y)
        i = (y * self.width + x) * 3
        w = self.col_weight
        for k in range(3):
            self.pixels[i + k] = np.float32(pixel[k] * w + float(self.pixels[i + k]) * (1 - w))

    def image(self) -> np.ndarray:
        """[height, width, 3] float32 view of the pixel store."""
        return self.pixels.reshape(self.height, self.width, 3)
