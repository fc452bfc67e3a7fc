"""Sky (src/sky/sky.ts:26-34, src/sky/sky_sphere.ts:22-27).  get_color(dir) = texture lookup at
uv_map_sphere(dir), evaluated on the GPU for rays that leave the tree."""
from .texture import Texture


class Sky:
    def __init__(self, texture: Texture):
        self.texture = texture


class SkySphere(Sky):
    pass
