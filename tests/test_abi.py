"""The C-ABI shared library: it loads, exports every symbol include/rt_b200.h declares, and refuses
to work without a CUDA device (no CPU fallback).  No compute calls here (CPU suite)."""
import ctypes as C
import os
import re

import pytest

import raytracer_js_b200 as rt
from raytracer_js_b200 import _native as N
from raytracer_js_b200 import build as rtbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    rtbuild.build()
    return N.load()


def test_exports_match_header(lib):
    names = declared_functions()
    assert "rt_render" in names and "rt_scene_upload" in names and len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"librt_b200.so does not export {n}"
    assert sorted(N.EXPORTS) == names  # the python binding covers the whole ABI, nothing more
    assert lib.rt_abi_version() == N.RT_B200_ABI_VERSION


def test_struct_sizes_match_header(lib, tmp_path):
    """sizeof of the ctypes mirrors == sizeof in C (compiled from the header with gcc)."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "rt_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(rt_scene_desc),sizeof(rt_camera),sizeof(rt_params),sizeof(rt_counters),sizeof(rt_tone),'
                   'sizeof(rt_exposure_stats));return 0;}\n')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(N.SceneDesc), C.sizeof(N.CameraDesc), C.sizeof(N.Params), C.sizeof(N.Counters),
                     C.sizeof(N.Tone), C.sizeof(N.ExposureStats)]


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="CUDA device present: the no-device error path cannot be exercised")
def test_fails_loudly_without_gpu(lib):
    ctx = C.c_void_p()
    st = lib.rt_create(-1, C.byref(ctx))
    assert st == N.RT_ERR_CUDA and not ctx.value
    assert b"no CUDA device" in lib.rt_last_error(None) and b"no CPU path" in lib.rt_last_error(None)
    # and through the host API: constructing the renderer raises instead of falling back
    from raytracer_js_b200 import scenes
    b = scenes.random_spheres(10)
    cam = scenes.bench_camera(16, 16)
    with pytest.raises(N.RtError):
        rt.GpuRaytracer(rt.RaytracerConfig(1, b.sky, b.default_substance, 1.0), b.tree, cam,
                        rt.ExposureBuffer(16, 16), rt.FpLcg(1))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under raytracer.js_b200/ may reference it."""
    pkg = os.path.join(ROOT, "raytracer.js_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".ts", ".c")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "hostsim" not in text.replace("tests/hostsim", ""), f


def test_napi_shim_compiles_against_the_abi():
    """bindings/node/rt_napi.c (the reference-side N-API addon) is valid C against include/rt_b200.h and
    Node's N-API prototypes (hand-declared in node_api_min.h: the image has no Node headers), and every
    rt_* function it calls is declared by the header."""
    import subprocess
    shim = os.path.join(ROOT, "bindings", "node", "rt_napi.c")
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), shim])
    used = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", open(shim).read(), flags=re.S)))
    assert used <= set(declared_functions()), used - set(declared_functions())
    assert {"rt_create", "rt_scene_upload", "rt_render", "rt_destroy"} <= used
