"""eidola-b200: B200-native drop-in for the per-frame render loop of
IwakuraRein/CIS-565-Final-VR-Raytracer (Renderer::run + the acceleration structure it traverses).

This Python layer is plumbing only: it loads the C-ABI shared library built from csrc/ (hand-written
sm_100a CUDA + C++ host) and mirrors the reference's Scene / AccelStructure / Renderer classes
(reference src/scene.hpp:60-83, src/accelstruct.hpp:40-46, src/renderer.hpp:52-61) on top of it.

There is no CPU fallback: if libeidola.so is missing or no GPU is usable, calls raise EidolaError.
"""
import ctypes as C
import os

import numpy as np

from . import abi, scenes, sharding
from .abi import (SceneCamera, RtxState, SceneInfo, AccelInfo, FrameStats, GroupInfo, PipelineLayout, SceneArrays, default_rtx_state)

_HERE = os.path.dirname(os.path.abspath(__file__))
# EIDOLA_LIB selects another build of the same library (kernel-variant sweeps, tools/sweep.sh); never a different backend
LIB_PATH = os.environ.get("EIDOLA_LIB") or os.path.join(_HERE, "libeidola.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")
_lib = None


class EidolaError(RuntimeError):
    pass


# every symbol include/eidola.h declares (tests check the .so exports all of them)
EXPORTS = [
    "eid_last_error", "eid_version", "eid_device_count",
    "eid_scene_create", "eid_scene_load_gltf", "eid_scene_load_desc", "eid_scene_provide_image", "eid_scene_destroy", "eid_scene_set_lookat",
    "eid_scene_update_camera", "eid_scene_set_camera", "eid_scene_get_camera", "eid_scene_get_info",
    "eid_scene_table_bytes", "eid_scene_read_table",
    "eid_accel_build", "eid_accel_build_ex", "eid_accel_destroy", "eid_accel_get_info", "eid_accel_trace", "eid_accel_sah_tap",
    "eid_env_create", "eid_env_load_hdr", "eid_env_destroy", "eid_env_integral", "eid_env_average", "eid_env_get_size", "eid_env_read",
    "eid_renderer_set_env",
    "eid_renderer_create", "eid_renderer_resize", "eid_renderer_destroy", "eid_renderer_set_env_constant",
    "eid_renderer_set_strict_math", "eid_renderer_set_denoise_rows", "eid_renderer_set_denoise_tiles", "eid_renderer_set_overlap", "eid_renderer_set_pipeline", "eid_renderer_set_variant", "eid_renderer_set_wavefront", "eid_renderer_set_sun_and_sky", "eid_sun_and_sky_eval", "eid_fn_tap", "eid_renderer_fn_tap", "eid_renderer_run_output", "eid_renderer_run", "eid_renderer_sync", "eid_renderer_get_outputs", "eid_renderer_buffer_bytes",
    "eid_renderer_read", "eid_renderer_write", "eid_renderer_render_host", "eid_renderer_render_host_async", "eid_renderer_wait_host", "eid_renderer_set_profiling",
    "eid_renderer_get_stats", "eid_renderer_set_band", "eid_renderer_run_trace", "eid_renderer_run_post", "eid_renderer_run_post_band", "eid_renderer_run_direct", "eid_renderer_run_indirect",
    "eid_renderer_band_range", "eid_renderer_set_stripes", "eid_renderer_exchange_groups", "eid_renderer_exchange_range",
    "eid_group_layout", "eid_group_unique_id", "eid_group_create", "eid_group_destroy", "eid_group_set_mode", "eid_group_run",
    "eid_group_render_host_async", "eid_group_wait_host", "eid_group_sync", "eid_group_get_info",
    "eid_group_pipeline_layout", "eid_group_random_id", "eid_group_create_pipeline",
]


def lib():
    """Load libeidola.so (once).  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EidolaError("libeidola.so not found at %s - build it with `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (nvcc, sm_100a); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_size_t
    sig = {
        "eid_last_error": (C.c_char_p, []),
        "eid_version": (i32, []),
        "eid_device_count": (i32, []),
        "eid_scene_create": (i32, [C.POINTER(vp), i32]),
        "eid_scene_load_gltf": (i32, [vp, C.c_char_p]),
        "eid_scene_load_desc": (i32, [vp, C.POINTER(abi.SceneDesc)]),
        "eid_scene_provide_image": (i32, [vp, u32, vp, u32, u32]),
        "eid_scene_destroy": (None, [vp]),
        "eid_scene_set_lookat": (i32, [vp, abi.c_float_p, abi.c_float_p, abi.c_float_p, C.c_float]),
        "eid_scene_update_camera": (i32, [vp, u32, u32]),
        "eid_scene_set_camera": (i32, [vp, C.POINTER(SceneCamera)]),
        "eid_scene_get_camera": (i32, [vp, C.POINTER(SceneCamera)]),
        "eid_scene_get_info": (i32, [vp, C.POINTER(SceneInfo)]),
        "eid_scene_table_bytes": (C.c_int64, [vp, i32, u32]),
        "eid_scene_read_table": (i32, [vp, i32, u32, vp, sz]),
        "eid_accel_build": (i32, [vp, C.POINTER(vp)]),
        "eid_accel_build_ex": (i32, [vp, i32, C.POINTER(vp)]),
        "eid_accel_destroy": (None, [vp]),
        "eid_accel_get_info": (i32, [vp, C.POINTER(AccelInfo)]),
        "eid_accel_trace": (i32, [vp, vp, u32, i32, vp]),
        "eid_accel_sah_tap": (i32, [vp, vp, u32, i32, vp, vp, vp, vp, vp, vp, vp]),
        "eid_env_create": (i32, [C.POINTER(vp), i32, vp, u32, u32]),
        "eid_env_load_hdr": (i32, [C.POINTER(vp), i32, C.c_char_p]),
        "eid_env_destroy": (None, [vp]),
        "eid_env_integral": (C.c_float, [vp]),
        "eid_env_average": (C.c_float, [vp]),
        "eid_env_get_size": (i32, [vp, C.POINTER(u32), C.POINTER(u32)]),
        "eid_env_read": (i32, [vp, i32, vp, sz]),
        "eid_renderer_set_env": (i32, [vp, vp]),
        "eid_renderer_create": (i32, [C.POINTER(vp), vp, vp, u32, u32, vp]),
        "eid_renderer_resize": (i32, [vp, u32, u32]),
        "eid_renderer_destroy": (None, [vp]),
        "eid_renderer_set_env_constant": (i32, [vp, abi.c_float_p]),
        "eid_renderer_set_strict_math": (i32, [vp, i32]),
        "eid_renderer_set_denoise_rows": (i32, [vp, i32]),
        "eid_renderer_set_denoise_tiles": (i32, [vp, i32, i32]),
        "eid_renderer_set_wavefront": (i32, [vp, i32, i32]),
        "eid_renderer_set_sun_and_sky": (i32, [vp, vp]),
        "eid_sun_and_sky_eval": (i32, [i32, vp, vp, u32, vp]),
        "eid_fn_tap": (i32, [i32, i32, vp, u32, vp]),
        "eid_renderer_fn_tap": (i32, [vp, vp, i32, vp, u32, vp]),
        "eid_renderer_run_output": (i32, [vp, vp]),
        "eid_renderer_set_overlap": (i32, [vp, i32]),
        "eid_renderer_set_pipeline": (i32, [vp, i32]),
        "eid_renderer_set_variant": (i32, [vp, i32]),
        "eid_renderer_run": (i32, [vp, C.POINTER(RtxState), i32]),
        "eid_renderer_sync": (i32, [vp]),
        "eid_renderer_get_outputs": (i32, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "eid_renderer_buffer_bytes": (C.c_int64, [vp, i32]),
        "eid_renderer_read": (i32, [vp, i32, vp, sz]),
        "eid_renderer_write": (i32, [vp, i32, vp, sz]),
        "eid_renderer_render_host": (i32, [vp, C.POINTER(SceneCamera), C.POINTER(RtxState), i32, vp, vp]),
        "eid_renderer_render_host_async": (i32, [vp, C.POINTER(SceneCamera), C.POINTER(RtxState), i32, vp, vp]),
        "eid_renderer_wait_host": (i32, [vp]),
        "eid_renderer_set_profiling": (i32, [vp, i32]),
        "eid_renderer_get_stats": (i32, [vp, C.POINTER(FrameStats)]),
        "eid_renderer_set_band": (i32, [vp, u32, u32]),
        "eid_renderer_run_trace": (i32, [vp, C.POINTER(RtxState), i32]),
        "eid_renderer_run_post": (i32, [vp, C.POINTER(RtxState), i32]),
        "eid_renderer_run_direct": (i32, [vp, C.POINTER(RtxState), i32]),
        "eid_renderer_run_indirect": (i32, [vp, C.POINTER(RtxState), i32]),
        "eid_renderer_run_post_band": (i32, [vp, C.POINTER(RtxState), i32]),
        "eid_renderer_band_range": (i32, [vp, i32, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64)]),
        "eid_renderer_set_stripes": (i32, [vp, u32, u32, u32]),
        "eid_renderer_exchange_groups": (i32, [vp]),
        "eid_renderer_exchange_range": (i32, [vp, i32, u32, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64)]),
        "eid_group_layout": (i32, [u32, i32, i32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "eid_group_unique_id": (i32, [vp]),
        "eid_group_create": (i32, [C.POINTER(vp), vp, i32, i32, vp]),
        "eid_group_destroy": (None, [vp]),
        "eid_group_set_mode": (i32, [vp, i32, i32, i32]),
        "eid_group_run": (i32, [vp, C.POINTER(RtxState), i32]),
        "eid_group_render_host_async": (i32, [vp, C.POINTER(SceneCamera), C.POINTER(RtxState), i32, vp, vp]),
        "eid_group_wait_host": (i32, [vp]),
        "eid_group_sync": (i32, [vp]),
        "eid_group_get_info": (i32, [vp, C.POINTER(GroupInfo)]),
        "eid_group_pipeline_layout": (i32, [u32, i32, i32, i32, i32, i32, C.POINTER(PipelineLayout)]),
        "eid_group_random_id": (i32, [vp]),
        "eid_group_create_pipeline": (i32, [C.POINTER(vp), vp, i32, i32, vp, u32, i32, i32, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        msg = lib().eid_last_error()
        raise EidolaError("eidola error %d: %s" % (rc, msg.decode() if msg else "?"))


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Scene:
    """Mirror of the reference's Scene (src/scene.hpp:60-83): load / updateCamera / getters / destroy."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().eid_scene_create(C.byref(self._h), device))

    def load(self, filename):            # Scene::load(const std::string&) -> bool (scene.cpp:57)
        _check(lib().eid_scene_load_gltf(self._h, os.fsencode(filename)))
        return True

    def provide_image(self, index, rgba8):   # decoded image `index` for the next load(): (h, w, 4) uint8
        a = np.ascontiguousarray(rgba8, np.uint8)
        _check(lib().eid_scene_provide_image(self._h, index, a.ctypes.data, a.shape[1], a.shape[0]))

    def load_arrays(self, arrays):       # harness path: arrays already in nvh::GltfScene shape
        d = arrays.desc()
        _check(lib().eid_scene_load_desc(self._h, C.byref(d)))
        return True

    def set_lookat(self, eye, center, up, fov_deg):   # CameraManip.setCamera
        _check(lib().eid_scene_set_lookat(self._h, _f3(eye), _f3(center), _f3(up), float(fov_deg)))

    def update_camera(self, width, height):           # Scene::updateCamera(cmdBuf, size) (scene.cpp:777)
        _check(lib().eid_scene_update_camera(self._h, width, height))

    def set_camera(self, cam):
        _check(lib().eid_scene_set_camera(self._h, C.byref(cam)))

    def get_camera(self):                              # Scene::getCamera
        cam = SceneCamera()
        _check(lib().eid_scene_get_camera(self._h, C.byref(cam)))
        return cam

    def info(self):                                    # Scene::getStat / m_trigLightWeight / m_puncLightWeight
        i = SceneInfo()
        _check(lib().eid_scene_get_info(self._h, C.byref(i)))
        return i

    def table(self, which, index=0):
        n = lib().eid_scene_table_bytes(self._h, which, index)
        if n < 0:
            raise EidolaError("no such table %d[%d]" % (which, index))
        buf = np.empty(n, np.uint8)
        if n:
            _check(lib().eid_scene_read_table(self._h, which, index, buf.ctypes.data, n))
        return buf.view(abi.TABLE_DTYPES[which])

    def destroy(self):
        if self._h:
            lib().eid_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class AccelStructure:
    """Mirror of AccelStructure (src/accelstruct.hpp:40-46): create(scene) / destroy; plus a batch ray tap."""

    def __init__(self):
        self._h = C.c_void_p()

    def create(self, scene, mode=abi.ACCEL_AUTO):   # AccelStructure::create(gltfScene, vertexBufs, indexBufs); mode: flat / two-level (BLAS + TLAS)
        self.destroy()
        _check(lib().eid_accel_build_ex(scene._h, int(mode), C.byref(self._h)))
        self._scene = scene

    def info(self):
        i = AccelInfo()
        _check(lib().eid_accel_get_info(self._h, C.byref(i)))
        return i

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros(rays.shape[0], abi.HIT_DT)
        _check(lib().eid_accel_trace(self._h, rays.ctypes.data, rays.shape[0], int(any_hit), hits.ctypes.data))
        return hits

    def destroy(self):
        if self._h:
            lib().eid_accel_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def sah_tree(lo, hi, threads=0):
    """Tap of the EID_ACCEL_FAST_TRACE topology builder (csrc/sah_host.cpp; host only): dict of order / left / right / parentInner /
    parentLeaf / rangeFirst / rangeLast for the boxes lo, hi (n x 3 float32)."""
    lo = np.ascontiguousarray(lo, np.float32).reshape(-1, 3)
    hi = np.ascontiguousarray(hi, np.float32).reshape(-1, 3)
    n = lo.shape[0]
    ni = max(n - 1, 0)
    out = dict(order=np.zeros(n, np.uint32), left=np.zeros(ni, np.int32), right=np.zeros(ni, np.int32), parentInner=np.zeros(ni, np.int32),
               parentLeaf=np.zeros(n, np.int32), rangeFirst=np.zeros(ni, np.int32), rangeLast=np.zeros(ni, np.int32))
    _check(lib().eid_accel_sah_tap(lo.ctypes.data, hi.ctypes.data, n, int(threads), *[out[k].ctypes.data for k in
                                   ("order", "left", "right", "parentInner", "parentLeaf", "rangeFirst", "rangeLast")]))
    return out


class HdrSampling:
    """Mirror of HdrSampling (src/hdr_sampling.hpp:43-48): loadEnvironment / getIntegral / getAverage."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self.device = device

    def load_environment(self, path):               # HdrSampling::loadEnvironment(hdrImage): Radiance .hdr
        self.destroy()
        _check(lib().eid_env_load_hdr(C.byref(self._h), self.device, os.fsencode(path)))

    def set_pixels(self, rgba):                     # RGBA32F texels already in memory, shape (h, w, 4)
        self.destroy()
        a = np.ascontiguousarray(rgba, np.float32)
        h, w = a.shape[0], a.shape[1]
        _check(lib().eid_env_create(C.byref(self._h), self.device, a.ctypes.data, w, h))

    def get_integral(self):
        return float(lib().eid_env_integral(self._h))

    def get_average(self):
        return float(lib().eid_env_average(self._h))

    def size(self):
        w, h = C.c_uint32(), C.c_uint32()
        _check(lib().eid_env_get_size(self._h, C.byref(w), C.byref(h)))
        return w.value, h.value

    def accel(self):
        w, h = self.size()
        out = np.zeros(w * h, abi.IMPT_DT)
        _check(lib().eid_env_read(self._h, 0, out.ctypes.data, out.nbytes))
        return out

    def pixels(self):
        w, h = self.size()
        out = np.zeros((h, w, 4), np.float32)
        _check(lib().eid_env_read(self._h, 1, out.ctypes.data, out.nbytes))
        return out

    def destroy(self):
        if self._h:
            lib().eid_env_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Group:
    """eid_group: this process's rank of an N-GPU frame (include/eidola.h).  `id128` comes from Group.unique_id() on rank 0."""

    def __init__(self):
        self._h = C.c_void_p()

    @staticmethod
    def layout(height, world, rank=0):
        y0, y1, ph = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().eid_group_layout(height, world, rank, C.byref(y0), C.byref(y1), C.byref(ph)))
        return y0.value, y1.value, ph.value

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        _check(lib().eid_group_unique_id(buf))
        return bytes(buf)

    def create(self, renderer, rank, world, id128=None):
        self.destroy()
        buf = (C.c_ubyte * 128).from_buffer_copy(id128) if id128 is not None else None
        _check(lib().eid_group_create(C.byref(self._h), renderer._h, rank, world, buf))
        self._renderer = renderer

    @staticmethod
    def pipeline_layout(height, world, rank=0, stages=(0, 0, 0)):
        """eid_group_pipeline_layout: who runs which stage on which rows (stages = ranks per stage, zeros = default split)."""
        out = PipelineLayout()
        _check(lib().eid_group_pipeline_layout(height, world, rank, stages[0], stages[1], stages[2], C.byref(out)))
        return out

    @staticmethod
    def random_id():
        buf = (C.c_ubyte * 128)()
        _check(lib().eid_group_random_id(buf))
        return bytes(buf)

    def create_pipeline(self, renderer, rank, world, id128, height, stages=(0, 0, 0)):
        """Stage pipeline over CUDA-IPC peer mappings (one process per rank); the renderer must have pipeline_layout().paddedHeight rows."""
        self.destroy()
        buf = (C.c_ubyte * 128).from_buffer_copy(id128)
        _check(lib().eid_group_create_pipeline(C.byref(self._h), renderer._h, rank, world, buf, height, stages[0], stages[1], stages[2]))
        self._renderer = renderer

    def set_mode(self, post_sharded=True, history=2, gather_final=True):
        _check(lib().eid_group_set_mode(self._h, int(post_sharded), int(history), int(gather_final)))

    def run(self, state, frames):
        _check(lib().eid_group_run(self._h, C.byref(state), frames))

    def render_host_async(self, cam, state, frames, direct_out, indirect_out):
        _check(lib().eid_group_render_host_async(
            self._h, C.byref(cam) if cam is not None else None, C.byref(state), frames,
            C.c_void_p(direct_out) if direct_out else None, C.c_void_p(indirect_out) if indirect_out else None))

    def wait_host(self):
        _check(lib().eid_group_wait_host(self._h))

    def sync(self):
        _check(lib().eid_group_sync(self._h))

    def info(self):
        i = GroupInfo()
        _check(lib().eid_group_get_info(self._h, C.byref(i)))
        return i

    def destroy(self):
        if self._h:
            lib().eid_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Renderer:
    """Mirror of Renderer (src/renderer.hpp:52-61): create / run / update / destroy."""

    def __init__(self):
        self._h = C.c_void_p()

    def create(self, size, scene, accel, stream=None):   # Renderer::create(size, layouts, Scene*)
        self.destroy()
        w, h = size
        _check(lib().eid_renderer_create(C.byref(self._h), scene._h, accel._h, w, h, C.c_void_p(stream or 0)))
        self._scene, self._accel, self.size = scene, accel, (w, h)

    def update(self, size):               # Renderer::update(size) — resize, history dropped
        _check(lib().eid_renderer_resize(self._h, size[0], size[1]))
        self.size = tuple(size)

    def set_env(self, env):               # install an HdrSampling environment (None -> constant environment)
        _check(lib().eid_renderer_set_env(self._h, env._h if env is not None else None))
        self._env = env

    def set_env_constant(self, rgb):
        _check(lib().eid_renderer_set_env_constant(self._h, _f3(rgb)))

    def set_variant(self, flags):          # abi.VARIANT_* bits: the reference's compile-time shader switches
        _check(lib().eid_renderer_set_variant(self._h, int(flags)))

    def set_pipeline(self, frames_in_flight):    # 1 strict (default); 2: direct_stage of frame f + 1 overlaps the later stages of frame f
        _check(lib().eid_renderer_set_pipeline(self._h, int(frames_in_flight)))

    def set_denoise_tiles(self, mode, rows=0):   # 1 (default): smem tiles via TMA, 2: via cp.async, 0: legacy L1-served kernel
        _check(lib().eid_renderer_set_denoise_tiles(self._h, int(mode), int(rows)))

    def set_denoise_rows(self, n):        # A-Trous pixels per thread sharing tap rows: 1, 2 (default) or 4
        _check(lib().eid_renderer_set_denoise_rows(self._h, int(n)))

    def run_output(self, tm):             # RenderOutput::run -> post.frag; results in BUF_DISPLAY_F32 / BUF_DISPLAY_RGBA8
        _check(lib().eid_renderer_run_output(self._h, C.byref(tm)))

    def set_sun_and_sky(self, ss):        # abi.SunAndSky (host_device.h:353-376); in_use = 1 selects the procedural sky
        _check(lib().eid_renderer_set_sun_and_sky(self._h, C.byref(ss)))

    def set_wavefront(self, on, trace_blocks=0):   # K2 as ray queues + dynamic-fetch traversal (default on); bit-identical either way
        _check(lib().eid_renderer_set_wavefront(self._h, int(on), int(trace_blocks)))

    def set_strict_math(self, on):        # bit-reproducible denoiser exp (parity runs); default is the fast MUFU path
        _check(lib().eid_renderer_set_strict_math(self._h, int(on)))

    def set_overlap(self, on):            # K3 on a second stream beside K2/K4 (default on)
        _check(lib().eid_renderer_set_overlap(self._h, int(on)))

    def run(self, state, frames):         # Renderer::run(cmdBuf, state, profiler, descSets, frames); async
        _check(lib().eid_renderer_run(self._h, C.byref(state), frames))

    def run_trace(self, state, frames):
        _check(lib().eid_renderer_run_trace(self._h, C.byref(state), frames))

    def run_direct(self, state, frames):
        _check(lib().eid_renderer_run_direct(self._h, C.byref(state), frames))

    def run_indirect(self, state, frames):
        _check(lib().eid_renderer_run_indirect(self._h, C.byref(state), frames))

    def run_post(self, state, frames):
        _check(lib().eid_renderer_run_post(self._h, C.byref(state), frames))

    def run_post_band(self, state, frames):
        _check(lib().eid_renderer_run_post_band(self._h, C.byref(state), frames))

    def set_band(self, y0, y1):
        _check(lib().eid_renderer_set_band(self._h, y0, y1))

    def set_stripes(self, rank, world, stripe_rows):
        _check(lib().eid_renderer_set_stripes(self._h, rank, world, stripe_rows))

    def exchange_groups(self):
        return lib().eid_renderer_exchange_groups(self._h)

    def exchange_range(self, which, group):
        base, off, n = C.c_void_p(), C.c_uint64(), C.c_uint64()
        _check(lib().eid_renderer_exchange_range(self._h, which, group, C.byref(base), C.byref(off), C.byref(n)))
        return base.value, off.value, n.value

    def band_range(self, which):
        base, off, n = C.c_void_p(), C.c_uint64(), C.c_uint64()
        _check(lib().eid_renderer_band_range(self._h, which, C.byref(base), C.byref(off), C.byref(n)))
        return base.value, off.value, n.value

    def sync(self):
        _check(lib().eid_renderer_sync(self._h))

    def outputs(self):
        d, i = C.c_void_p(), C.c_void_p()
        _check(lib().eid_renderer_get_outputs(self._h, C.byref(d), C.byref(i)))
        return d.value, i.value

    def read(self, which):
        n = lib().eid_renderer_buffer_bytes(self._h, which)
        if n < 0:
            raise EidolaError("no such buffer %d" % which)
        buf = np.empty(n, np.uint8)
        _check(lib().eid_renderer_read(self._h, which, buf.ctypes.data, n))
        return buf.view(abi.BUFFER_DTYPES[which])

    def write(self, which, arr):
        a = np.ascontiguousarray(arr)
        _check(lib().eid_renderer_write(self._h, which, a.ctypes.data, a.nbytes))

    def render_host(self, cam, state, frames, direct_out, indirect_out):
        _check(lib().eid_renderer_render_host(
            self._h, C.byref(cam) if cam is not None else None, C.byref(state), frames,
            C.c_void_p(direct_out) if direct_out else None, C.c_void_p(indirect_out) if indirect_out else None))

    def render_host_async(self, cam, state, frames, direct_out, indirect_out):
        _check(lib().eid_renderer_render_host_async(
            self._h, C.byref(cam) if cam is not None else None, C.byref(state), frames,
            C.c_void_p(direct_out) if direct_out else None, C.c_void_p(indirect_out) if indirect_out else None))

    def wait_host(self):
        _check(lib().eid_renderer_wait_host(self._h))

    def set_profiling(self, on):
        _check(lib().eid_renderer_set_profiling(self._h, int(on)))

    def stats(self):
        s = FrameStats()
        _check(lib().eid_renderer_get_stats(self._h, C.byref(s)))
        return s

    def destroy(self):
        if self._h:
            lib().eid_renderer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
