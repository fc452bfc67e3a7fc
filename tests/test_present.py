"""View.draw_ebuffer(): exposure statistics, tone mappers and the 8-bit screen.
CPU part: the oracle's restatement against values worked out by hand from the reference's source lines (no
reference test covers src/view/exposure_buffer.ts:93-158, tone_mapping.ts or screen_canvas.ts: "parity unpinned
by tests").  GPU part: rt_present / rt_render_present through the C ABI against the oracle."""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as orc
import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import scenes


@pytest.fixture(scope="module")
def oracle():
    orc.build()
    return orc


def test_oracle_stats_by_hand(oracle):
    # two pixels: y0 = 0.299*1 + 0.587*0.5 + 0.114*0.25, y1 = 0 -> mean y0/2, variance (y0/2)^2, absdev y0/2
    px = np.array([[1.0, 0.5, 0.25], [0.0, 0.0, 0.0]], np.float32)
    y0 = 0.299 * 1.0 + 0.587 * 0.5 + 0.114 * 0.25
    mean, var, dev = oracle.exposure_stats(px)
    assert mean == y0 / 2
    assert var == ((y0 - y0 / 2) ** 2 + (0 - y0 / 2) ** 2) / 2
    assert dev == (abs(y0 - y0 / 2) + abs(0 - y0 / 2)) / 2


def test_oracle_tone_mappers_by_hand(oracle):
    stats = (0.5, 0.04, 0.15)
    assert oracle.dynamic_range(orc.TONE_IDENTITY, 8, 1 / 256, 8.0, stats) == (0.0, 1.0)
    # StdDevAroundMean (src/view/tone_mapping.ts:48-63): max = min(mean + sqrt(var), max_dynamic), min = max / 2^8
    hi = min(0.5 + math.sqrt(0.04), 8.0)
    assert hi / 256 < 1 / 256  # ... which is below min_dynamic: the bottom is min_dynamic and the top follows it
    assert oracle.dynamic_range(orc.TONE_STDDEV, 8, 1 / 256, 8.0, stats) == (1 / 256, (1 / 256) * 256)
    assert oracle.dynamic_range(orc.TONE_STDDEV, 2, 1 / 256, 8.0, stats) == (hi / 4, hi)
    # AbsDevAroundMean (:65-79), lower clamp: min_dynamic wins and the top follows it
    lo, hi2 = oracle.dynamic_range(orc.TONE_ABSDEV, 4, 0.1, 8.0, stats)
    assert (lo, hi2) == (0.1, 0.1 * 16)
    # upper clamp
    assert oracle.dynamic_range(orc.TONE_STDDEV, 2, 0.0, 0.6, stats) == (0.6 / 4, 0.6)


def test_oracle_discretize_by_hand(oracle):
    # range [0, 1]: scale = (y - 0)/1 / (y + eps) ~ 1 - eps/y, so a channel c maps to trunc(fl32(c * scale) * 255);
    # the blue channel is dropped (`slice(i, i+2)`) and alpha is 0xff; out-of-range values clamp, NaN -> 0
    px = np.array([[0.5, 0.25, 0.75], [2.0, -1.0, 0.3], [0.0, 0.0, 0.0], [np.nan, 0.5, 0.5]], np.float32)
    img = oracle.discretize(px, 0.0, 1.0)
    assert img[0].tolist() == [127, 63, 0, 255]
    assert img[1].tolist() == [255, 0, 0, 255]
    assert img[2].tolist() == [0, 0, 0, 255]
    assert img[3].tolist() == [0, 0, 0, 255]  # NaN luma -> NaN scale -> NaN -> 0
    # a compressing range: low = 0.25, high = 0.75, grey 0.5 -> cmpr 0.5, scale = 0.5 / (0.5 + eps) -> 0.5 * scale * 255
    img = oracle.discretize(np.array([[0.5, 0.5, 0.5]], np.float32), 0.25, 0.75)
    assert img[0].tolist() == [127, 127, 0, 255]


def test_host_screen_convert_color():
    s = rt.Screen(2, 1)
    s.set_pixel_i(0, [0.5, 2.0])  # two channels, as discretize_to_screen hands them over
    s.set_pixel_i(1, [1.0, -3.0, 0.25])
    assert s.image.reshape(-1, 4).tolist() == [[127, 255, 0, 255], [255, 0, 63, 255]]
    assert s.dynamic_range == 8
    with pytest.raises(TypeError):
        rt.ToneMapper().tone_desc()  # unknown subclass: unsupported, no CPU fallback
    assert rt.ToneMapper_StdDevAroundMean(8, 1 / 256, 8.0).dynamic_coef == 256


# ----------------------------------------------------------------------------------------------- GPU
def _tracer(W, H, n=2500, mix="mirrors"):
    b = scenes.random_spheres(n, 0.01, 0.05, seed=6.0, mix=mix, box_fraction=0.1)
    cam = scenes.bench_camera(W, H)
    eb = rt.ExposureBuffer(W, H)
    return rt.GpuRaytracer(rt.RaytracerConfig(b.refmax, b.sky, b.default_substance, 1.0), b.tree, cam, eb, rt.FpLcg(1.0)), eb


def _check_image(oracle, pixels, tone, stats, image):
    ost = oracle.exposure_stats(pixels)
    for name, o in zip(("mean", "variance", "absolute_dev"), ost):
        assert stats[name] == pytest.approx(o, rel=1e-11, abs=1e-300), name  # parallel vs sequential float64 sum
    lo, hi = oracle.dynamic_range(tone.kind, tone.dynamic_range, tone.min_dynamic, tone.max_dynamic, ost)
    assert stats["drange_low"] == pytest.approx(lo, rel=1e-11) and stats["drange_high"] == pytest.approx(hi, rel=1e-11)
    # the 8-bit image for the GPU's own range is the oracle's, bit for bit; for the oracle's range (which
    # differs in the last bits of the sums) at most one level on a vanishing fraction of pixels
    exact = oracle.discretize(pixels, stats["drange_low"], stats["drange_high"])
    got = image.reshape(-1, 4)
    assert np.array_equal(got, exact)
    ref = oracle.discretize(pixels, lo, hi)
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-4
    assert (got[:, 2] == 0).all() and (got[:, 3] == 255).all()


@pytest.mark.gpu
@pytest.mark.parametrize("mapper", ["identity", "stddev", "absdev", "stddev-clamped"])
def test_present_matches_oracle(oracle, mapper):
    W = H = 240
    tracer, eb = _tracer(W, H)
    tracer.trace_frame(n_frames=3)
    tm = {"identity": rt.ToneMapper_Identity.instance, "stddev": rt.ToneMapper_StdDevAroundMean(8, 1 / 256, 8.0),
          "absdev": rt.ToneMapper_AbsDevAroundMean(8, 1 / 256, 8.0), "stddev-clamped": rt.ToneMapper_StdDevAroundMean(3, 0.2, 0.5)}[mapper]
    screen = rt.Screen(W, H)
    view = rt.View(eb, screen, tm, tracer)
    lo, hi = view.draw_ebuffer()
    assert (lo, hi) == (view.last_stats["drange_low"], view.last_stats["drange_high"])
    _check_image(oracle, eb.pixels, tm.tone_desc(), view.last_stats, screen.image)


@pytest.mark.gpu
def test_resident_exposure_only_sends_the_screen(oracle):
    """trace_and_draw keeps the ExposureBuffer on the device across frames; the float frame it accumulates is
    the one trace_frame() produces on the host path, and the screen is the oracle's image of it."""
    W, H = 256, 192
    tracer, eb = _tracer(W, H)
    tm = rt.ToneMapper_StdDevAroundMean(8, 1 / 256, 8.0)
    screen = rt.Screen(W, H)
    view = rt.View(eb, screen, tm, tracer)
    view.trace_and_draw(n_frames=2)
    eb.next_frame()
    view.trace_and_draw(n_frames=1)  # continues the resident exposure at frame_count = 2
    stats, image = dict(view.last_stats), screen.image.copy()
    view.download_exposure()
    resident = eb.pixels.copy()
    # the same three frames through the host path
    tracer2, eb2 = _tracer(W, H)
    tracer2.trace_frame(n_frames=3)
    assert np.array_equal(resident, eb2.pixels)
    _check_image(oracle, resident, tm.tone_desc(), stats, image)
    # a continued exposure needs a resident buffer of that size
    tracer3, eb3 = _tracer(64, 64, n=200)
    eb3.next_frame()
    with pytest.raises(N.RtError):
        rt.View(eb3, rt.Screen(64, 64), tm, tracer3).trace_and_draw()


@pytest.mark.gpu
def test_present_full_size_properties(oracle):
    """1920x1080: present of a constant frame is the constant's image; determinism; unknown mapper kinds fail."""
    W, H = 1920, 1080
    tracer, eb = _tracer(64, 64, n=100)
    lib, ctx = tracer.lib, tracer.ctx
    rng = np.random.default_rng(3)
    px = (rng.random(W * H * 3, dtype=np.float32) * np.float32(4.0)).astype(np.float32)
    out1, out2 = np.zeros(W * H * 4, np.uint8), np.zeros(W * H * 4, np.uint8)
    tone = N.Tone(N.RT_TONE_STDDEV, 8, 1 / 256, 8.0)
    st1, st2 = N.ExposureStats(), N.ExposureStats()
    N.check(ctx, lib.rt_present(ctx, px.ctypes.data, W, H, C.byref(tone), out1.ctypes.data, C.byref(st1)))
    N.check(ctx, lib.rt_present(ctx, px.ctypes.data, W, H, C.byref(tone), out2.ctypes.data, C.byref(st2)))
    assert np.array_equal(out1, out2) and st1.as_dict() == st2.as_dict()  # deterministic reductions
    _check_image(oracle, px, tone, st1.as_dict(), out1)
    bad = N.Tone(7, 8, 0.0, 1.0)
    assert lib.rt_present(ctx, px.ctypes.data, W, H, C.byref(bad), out1.ctypes.data, None) == N.RT_ERR_UNSUPPORTED


def test_oracle_present_against_a_python_transliteration(oracle):
    """get_mean / get_variance / get_absolute_dev, the three tone mappers, discretize_to_screen and convert_color
    transliterated in plain Python (sequential float64 sums, Float32Array rounding, Math.min/max NaN rules,
    `<< 0`): the C++ oracle must give the same numbers and the same bytes."""
    rng = np.random.default_rng(9)
    px = (rng.random((37 * 23, 3), dtype=np.float32) * np.float32(3.0)).astype(np.float32)
    px[5] = [0.0, 0.0, 0.0]
    px[6] = [np.nan, 0.5, 0.5]  # a NaN pixel poisons the sums, as in the reference
    for pixels in (px[:5].copy(), np.delete(px, 6, axis=0), px):
        ys = [0.299 * float(p[0]) + 0.587 * float(p[1]) + 0.114 * float(p[2]) for p in pixels]
        mean = 0.0
        for y in ys:
            mean += y
        mean /= len(ys)
        var = dev = 0.0
        for y in ys:
            delta = y - mean
            var += delta * delta
            dev += abs(delta)
        var /= len(ys)
        dev /= len(ys)
        got = oracle.exposure_stats(pixels)
        for g, w in zip(got, (mean, var, dev)):
            assert g == w or (math.isnan(g) and math.isnan(w))

        def js_min(a, b):
            return math.nan if (math.isnan(a) or math.isnan(b)) else min(a, b)

        def clamp(x, lo, hi):
            m = js_min(x, hi)
            return math.nan if (math.isnan(m) or math.isnan(lo)) else max(m, lo)

        for kind, stops, lo_lim, hi_lim in ((orc.TONE_IDENTITY, 8, 0.0, 1.0), (orc.TONE_STDDEV, 8, 1 / 256, 8.0), (orc.TONE_ABSDEV, 5, 0.01, 0.9)):
            if kind == orc.TONE_IDENTITY:
                lo, hi = 0.0, 1.0
            else:
                coef = float(1 << stops)
                d = math.sqrt(var) if kind == orc.TONE_STDDEV else dev
                hi = js_min(mean + d, hi_lim)
                lo = hi / coef
                if lo < lo_lim:
                    lo, hi = lo_lim, lo_lim * coef
            glo, ghi = oracle.dynamic_range(kind, stops, lo_lim, hi_lim, (mean, var, dev))
            assert (glo == lo or (math.isnan(glo) and math.isnan(lo))) and (ghi == hi or (math.isnan(ghi) and math.isnan(hi)))
            drange = hi - lo
            want = np.zeros((len(ys), 4), np.uint8)
            for i, (p, y) in enumerate(zip(pixels, ys)):
                with np.errstate(divide="ignore", invalid="ignore"):
                    scale = float(np.float64((y - lo) / drange if drange != 0 else np.float64(y - lo) / np.float64(drange)) / np.float64(y + 2.220446049250313e-16))
                for k in range(2):
                    c = float(np.float32(clamp(float(p[k]) * scale, 0.0, 1.0)))  # Float32Array.prototype.map
                    v = clamp(c, 0.0, 1.0) * 255
                    want[i, k] = 0 if math.isnan(v) else int(v)
                want[i, 3] = 255
            assert np.array_equal(oracle.discretize(pixels, lo, hi), want)
