"""TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (oracle/liboracle.so) and, where it could be
built (/root/reference present), of oracle/_ref/libref.so.  Mirrors the product's Python classes so parity
tests drive both with the same code.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
import this."""
import ctypes as C
import os
import subprocess

import numpy as np

import eidola_b200 as eid
from eidola_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref.so")
_lib = None
_ref = None


def build(force=False):
    if force or not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(REF_SO) or not os.path.exists(os.path.join(os.path.dirname(REF_SO), "librefscene.so"))):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        vp, i32, u32, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
        fp = abi.c_float_p
        sig = {
            "orc_fp_contract_selftest": (i32, []), "orc_num_threads": (i32, []), "orc_set_num_threads": (None, [i32]),
            "orc_tea": (u32, [u32, u32]), "orc_rand_chain": (None, [u32, i32, vp, vp]), "orc_hash8bit": (u32, [u32]),
            "orc_compress_unit_vec": (u32, [C.c_float] * 3), "orc_decompress_unit_vec": (None, [u32, vp]),
            "orc_offset_ray": (None, [vp, vp, vp]), "orc_alias_table": (None, [vp, i32, vp, vp]),
            "orc_pack_unorm4x8": (u32, [vp]), "orc_sizeof": (i32, [C.c_char_p]),
            "orc_detmath": (None, [i32, vp, vp, i32, vp]),
            "orc_scene_create": (vp, []), "orc_scene_destroy": (None, [vp]), "orc_scene_set_use_bvh": (None, [vp, i32]),
            "orc_scene_load_desc": (i32, [vp, C.POINTER(abi.SceneDesc)]),
            "orc_scene_set_lookat": (i32, [vp, fp, fp, fp, C.c_float]), "orc_scene_update_camera": (i32, [vp, u32, u32]),
            "orc_scene_set_camera": (i32, [vp, C.POINTER(abi.SceneCamera)]),
            "orc_scene_get_camera": (i32, [vp, C.POINTER(abi.SceneCamera)]),
            "orc_scene_get_info": (i32, [vp, C.POINTER(abi.SceneInfo)]),
            "orc_scene_table_bytes": (C.c_int64, [vp, i32, u32]), "orc_scene_read_table": (i32, [vp, i32, u32, vp, sz]),
            "orc_accel_trace": (i32, [vp, vp, u32, i32, vp]),
            "orc_env_create": (vp, [vp, u32, u32]), "orc_env_destroy": (None, [vp]), "orc_env_integral": (C.c_float, [vp]),
            "orc_env_average": (C.c_float, [vp]), "orc_env_read_accel": (i32, [vp, vp, sz]), "orc_env_texture": (None, [vp, vp, i32, vp]),
            "orc_renderer_set_env": (i32, [vp, vp]),
            "orc_renderer_create": (vp, [vp, u32, u32]), "orc_renderer_destroy": (None, [vp]),
            "orc_renderer_set_env_constant": (i32, [vp, fp]),
            "orc_fn": (i32, [i32, vp, i32, vp]), "orc_ctx_fn": (i32, [vp, vp, i32, vp, i32, vp]),
            "orc_renderer_set_sun_and_sky": (i32, [vp, vp]), "orc_renderer_set_variant": (i32, [vp, i32]), "orc_renderer_run_output": (i32, [vp, vp, vp, vp]), "orc_sun_and_sky": (None, [vp, vp, i32, vp]),
            "orc_renderer_run": (i32, [vp, C.POINTER(abi.RtxState), i32]),
            "orc_renderer_run_trace": (i32, [vp, C.POINTER(abi.RtxState), i32, i32, i32]),
            "orc_renderer_run_post": (i32, [vp, C.POINTER(abi.RtxState), i32]),
            "orc_renderer_get_stats": (i32, [vp, C.POINTER(abi.FrameStats)]),
            "orc_renderer_buffer_bytes": (C.c_int64, [vp, i32]), "orc_renderer_read": (i32, [vp, i32, vp, sz]),
            "orc_renderer_write": (i32, [vp, i32, vp, sz]),
            "orc_scene_instance_flags": (i32, [vp, vp, i32]),
            "orc_mip_chain_average": (None, [vp, i32, i32, vp]), "orc_tone_exposure": (None, [vp, vp, i32, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        assert L.orc_fp_contract_selftest() == 0, "oracle was compiled with FMA contraction"
        _lib = L
    return _lib


def ref():
    """oracle/_ref/libref.so (the reference's own compress.glsl / alias_table.hpp / host_device.h) or None."""
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(REF_SO):
            return None
        L = C.CDLL(REF_SO)
        L.ref_compress_unit_vec.restype, L.ref_compress_unit_vec.argtypes = C.c_uint32, [C.c_float] * 3
        L.ref_decompress_unit_vec.restype, L.ref_decompress_unit_vec.argtypes = None, [C.c_uint32, C.c_void_p]
        L.ref_pack_unorm4x8.restype, L.ref_pack_unorm4x8.argtypes = C.c_uint32, [C.c_void_p]
        L.ref_alias_table.restype, L.ref_alias_table.argtypes = None, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_sizeof.restype, L.ref_sizeof.argtypes = C.c_int, [C.c_char_p]
        L.ref_env_accel.restype, L.ref_env_accel.argtypes = None, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_fn.restype, L.ref_fn.argtypes = C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_sun_and_sky.restype, L.ref_sun_and_sky.argtypes = None, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_trace_run.restype, L.ref_trace_run.argtypes = None, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_trace_run_variant.restype, L.ref_trace_run_variant.argtypes = None, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_post_dispatch_variant.restype, L.ref_post_dispatch_variant.argtypes = C.c_int, [C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 7
        L.ref_post_run.restype, L.ref_post_run.argtypes = None, [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7
        L.ref_post_dispatch.restype, L.ref_post_dispatch.argtypes = C.c_int, [C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 7
        L.ref_display_run.restype, L.ref_display_run.argtypes = None, [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.ref_scene_set.restype, L.ref_scene_set.argtypes = None, [C.c_void_p] * 10 + [C.c_uint32, C.c_uint32]
        L.ref_ctx_fn.restype, L.ref_ctx_fn.argtypes = C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _ref = L
    return _ref


_ref_scene = None


def ref_scene_lib():
    """oracle/_ref/librefscene.so — the reference's own src/scene.cpp compiled where it lies (stand-ins: oracle/ref_shim/scene/) — or None."""
    global _ref_scene
    if _ref_scene is None:
        build()
        path = os.path.join(os.path.dirname(REF_SO), "librefscene.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_scene_load.restype, L.ref_scene_load.argtypes = C.c_void_p, [C.c_void_p]
        L.ref_scene_destroy.restype, L.ref_scene_destroy.argtypes = None, [C.c_void_p]
        L.ref_scene_table.restype, L.ref_scene_table.argtypes = C.c_long, [C.c_void_p, C.c_int, C.c_uint, C.c_void_p, C.c_long]
        L.ref_scene_weights.restype, L.ref_scene_weights.argtypes = None, [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_scene_set_lookat.restype, L.ref_scene_set_lookat.argtypes = None, [C.c_void_p] * 3 + [C.c_float]
        L.ref_scene_update_camera.restype, L.ref_scene_update_camera.argtypes = None, [C.c_void_p, C.c_uint, C.c_uint]
        L.ref_scene_accel.restype, L.ref_scene_accel.argtypes = C.c_int, [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_output_create.restype, L.ref_output_create.argtypes = C.c_void_p, [C.c_uint, C.c_uint]
        L.ref_output_destroy.restype, L.ref_output_destroy.argtypes = None, [C.c_void_p]
        L.ref_output_roles.restype, L.ref_output_roles.argtypes = None, [C.c_void_p, C.c_void_p]
        L.ref_output_wiring.restype, L.ref_output_wiring.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_output_run.restype, L.ref_output_run.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_glue_state.restype, L.ref_glue_state.argtypes = None, [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.ref_default_state.restype, L.ref_default_state.argtypes = C.c_int, [C.c_int, C.c_void_p, C.c_int]
        L.ref_renderer_create.restype, L.ref_renderer_create.argtypes = C.c_void_p, [C.c_uint, C.c_uint]
        L.ref_renderer_update.restype, L.ref_renderer_update.argtypes = None, [C.c_void_p, C.c_uint, C.c_uint]
        L.ref_renderer_destroy.restype, L.ref_renderer_destroy.argtypes = None, [C.c_void_p]
        L.ref_renderer_roles.restype, L.ref_renderer_roles.argtypes = None, [C.c_void_p, C.c_void_p]
        L.ref_renderer_resource.restype, L.ref_renderer_resource.argtypes = C.c_int, [C.c_int, C.c_void_p]
        L.ref_renderer_wiring.restype, L.ref_renderer_wiring.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_renderer_run.restype, L.ref_renderer_run.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _ref_scene = L
    return _ref_scene


class RefScene:
    """The reference's Scene (src/scene.cpp) loaded with a harness scene: Scene::load's table builders and Scene::updateCamera."""

    def __init__(self, arrays):
        self.L = ref_scene_lib()
        self._desc = arrays.desc()
        self._h = C.c_void_p(self.L.ref_scene_load(C.byref(self._desc)))
        assert self._h

    def table(self, which, index=0):
        """Raw bytes of table `which` (eid_scene_table numbering) or None."""
        n = self.L.ref_scene_table(self._h, which, index, None, 0)
        if n < 0:
            return None
        b = np.zeros(n, np.uint8)
        self.L.ref_scene_table(self._h, which, index, b.ctypes.data, n)
        return b

    def weights(self):
        t, p = C.c_float(), C.c_float()
        self.L.ref_scene_weights(self._h, C.byref(t), C.byref(p))
        return t.value, p.value

    def set_lookat(self, eye, center, up, fov_deg):
        self.L.ref_scene_set_lookat(_f3(eye), _f3(center), _f3(up), float(fov_deg))

    def update_camera(self, w, h):
        self.L.ref_scene_update_camera(self._h, w, h)

    def accel(self, max_inst=4096, max_blas=4096):
        """AccelStructure::create (src/accelstruct.cpp) on this scene -> (instances (n, 5) int32: instanceCustomIndex, mask, sbt offset, flags,
        BLAS index; transforms (n, 3, 4) float32 row-major; blas (m, 6) int32: primitiveCount, maxVertex, vertexStride, geometry flags,
        vertexFormat, indexType; (BLAS build flags, TLAS build flags))."""
        inst = np.zeros((max_inst, 5), np.int32); xf = np.zeros((max_inst, 3, 4), np.float32); blas = np.zeros((max_blas, 6), np.int32)
        nb, build = C.c_int(), (C.c_int * 2)()
        n = self.L.ref_scene_accel(self._h, inst.ctypes.data, xf.ctypes.data, max_inst, blas.ctypes.data, max_blas, C.byref(nb), build)
        return inst[:n].copy(), xf[:n].copy(), blas[:nb.value].copy(), (build[0], build[1])

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_scene_destroy(self._h)
            self._h = None


class RefOutput:
    """The reference's RenderOutput (src/render_output.cpp) on the recording stand-in device."""

    def __init__(self, w, h):
        self.L = ref_scene_lib()
        self._h = C.c_void_p(self.L.ref_output_create(w, h))

    def roles(self):
        """(resource ids of m_directResult[0], [1], m_indirectResult[0], [1]), (their mip level counts)."""
        o = np.zeros(8, np.int32)
        self.L.ref_output_roles(self._h, o.ctypes.data)
        return tuple(int(v) for v in o[:4]), tuple(int(v) for v in o[4:])

    def wiring(self):
        w = np.zeros((64, 3), np.int32)
        n = self.L.ref_output_wiring(self._h, w.ctypes.data, 64)
        return {(int(r[0]), int(r[1])): int(r[2]) for r in w[:n]}

    def run(self, state, zoom, ratio, frames, gen_mips):
        """genMipmap (optionally) + run -> (rows (what, a, b, c), pushed bytes: Tonemapper + debugging_mode)."""
        rows = np.zeros((32, 4), np.int32)
        push = np.zeros(64, np.uint8)
        n = self.L.ref_output_run(self._h, C.byref(state), float(zoom), float(ratio[0]), float(ratio[1]), frames, int(gen_mips), rows.ctypes.data, 32, push.ctypes.data, 64)
        return [tuple(int(v) for v in r) for r in rows[:n]], push.tobytes()

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_output_destroy(self._h)
            self._h = None


def ref_default_state(which):
    """Bytes of the reference's own default objects: 0 RtxState m_rtxState, 1 SunAndSky m_sunAndSky (sample_example.hpp:154-203), 2 Tonemapper m_tm,
    3 Tonemapper m_depthTm (render_output.hpp:44-60) — their initialisers lifted from the headers and compiled (oracle/ref_shim/defaults_prep.py)."""
    b = np.zeros(256, np.uint8)
    n = ref_scene_lib().ref_default_state(which, b.ctypes.data, 256)
    assert n > 0
    return b[:n].tobytes()


class RefRenderer:
    """The reference's Renderer (src/renderer.cpp) on the recording stand-in device (oracle/ref_shim/ref_renderer.cpp)."""
    ROLES = ("gbuffer0", "gbuffer1", "directResv0", "directResv1", "indirectResv0", "indirectResv1", "directTemp", "indirectTemp", "motion",
             "denoiseTemp0", "denoiseTemp1", "denoiseTemp2", "denoiseTemp3")

    def __init__(self, w, h):
        self.L = ref_scene_lib()
        self._h = C.c_void_p(self.L.ref_renderer_create(w, h))

    def update(self, w, h):
        self.L.ref_renderer_update(self._h, w, h)

    def resources(self):
        """role -> (resource id, kind 0 buffer / 1 image, bytes, width, height, VkFormat)."""
        ids = np.zeros(13, np.int32)
        self.L.ref_renderer_roles(self._h, ids.ctypes.data)
        out = {}
        for role, i in zip(self.ROLES, ids):
            o = np.zeros(5, np.int64)
            assert self.L.ref_renderer_resource(int(i), o.ctypes.data) == 0
            out[role] = (int(i),) + tuple(int(v) for v in o)
        return out

    def wiring(self):
        """{(set number 1|2, binding): (resource id, range bytes)} of the last updateDescriptorSet."""
        w = np.zeros((26, 4), np.int64)
        n = self.L.ref_renderer_wiring(self._h, w.ctypes.data, 26)
        return {(int(r[0]), int(r[1])): (int(r[2]), int(r[3])) for r in w[:n]}

    def run(self, state, frames):
        """Command sequence of Renderer::run: list of (what, a, b, c) + the pushed RtxState blobs (bytes)."""
        rows = np.zeros((64, 4), np.int32)
        push = np.zeros(64 * C.sizeof(state), np.uint8)
        n = self.L.ref_renderer_run(self._h, C.byref(state), frames, rows.ctypes.data, 64, push.ctypes.data, push.size)
        sz = C.sizeof(state)
        npush = int((rows[:n, 0] == 2).sum())
        return [tuple(int(v) for v in r) for r in rows[:n]], [push[i * sz:(i + 1) * sz].tobytes() for i in range(npush)]

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_renderer_destroy(self._h)
            self._h = None


def call_fn(L, name, which, x, nout):
    """orc_fn / ref_fn / eid_fn_tap: one row of `x` per item -> (n, nout) float32."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros((x.shape[0], nout), np.float32)
    rc = getattr(L, name)(which, x.ctypes.data, x.shape[0], out.ctypes.data)
    assert rc == 0, (name, which, rc)
    return out


def ctx_setup(scenes, abi, common, maker_name, kind):
    """Oracle scene + renderer (+ environment) of a tests/ref_fn_inputs.CTX_CONFIGS entry -> (scene, renderer, env or None, SunAndSky, RtxState)."""
    import ref_fn_inputs as fi
    osc = OracleScene()
    osc.load_arrays(getattr(scenes, maker_name)())
    orr = OracleRenderer(osc, fi.CTX_SIZE)
    orr.set_env_constant(common.ENV)
    osc.update_camera(*fi.CTX_SIZE)
    osc.update_camera(*fi.CTX_SIZE)
    ss = abi.default_sun_and_sky(in_use=0)
    env = None
    if kind == "hdr":
        env = OracleEnv(fi.ctx_env_image(scenes))
        orr.set_env(env)
    elif kind == "sky":
        ss = fi.sun_sky(abi, fi.CTX_SKY)
        orr.set_sun_and_sky(ss)
    st = fi.ctx_state(common, abi, osc.info(), kind, env.get_integral() if env else None)
    return osc, orr, env, ss, st


def ref_scene_set(R, osc, env, ss, st, abi):
    """Binds the oracle scene's tables (what layouts.glsl binds) to the reference-GLSL library; returns the arrays to keep alive."""
    tabs = [np.ascontiguousarray(osc.table(t)) for t in (abi.TABLE_CAMERA, abi.TABLE_LIGHT_INFO, abi.TABLE_MATERIALS, abi.TABLE_TRIG_LIGHTS, abi.TABLE_PUNC_LIGHTS)]
    acc = env.accel() if env else np.zeros(1, abi.IMPT_DT)
    fnp = C.cast(lib().orc_sample, C.c_void_p) if env else None     # the sampler is the contract's, not the reference's arithmetic
    R.ref_scene_set(C.addressof(st), tabs[0].ctypes.data, C.addressof(ss), tabs[1].ctypes.data, tabs[2].ctypes.data, tabs[3].ctypes.data, tabs[4].ctypes.data,
                    acc.ctypes.data, fnp, env._h if env else None, env.w if env else 0, env.h if env else 0)
    return tabs, acc


class RefTraceBind(C.Structure):      # oracle/ref_shim/ref_trace.cpp
    _fields_ = [(n, C.c_void_p) for n in ("state", "camera", "sunSky", "lightInfo", "geoInfo", "materials", "trigLights", "puncLights", "envAccel",
                                          "envSamplerFn", "env")] + [("envW", C.c_uint32), ("envH", C.c_uint32), ("traceFn", C.c_void_p), ("scene", C.c_void_p),
                                                                     ("allocW", C.c_int32), ("allocH", C.c_int32)] + [
        (n, C.c_void_p) for n in ("thisG", "lastG", "motion", "thisDR", "lastDR", "thisIR", "lastIR", "direct", "indirect", "indA", "instanceXforms", "tempDR", "dirA")]


class RefTracer:
    """Runs the reference's direct_stage.comp / indirect_stage.comp (compiled as C++, oracle/ref_shim/ref_trace.cpp) frame after frame on the
    tables of an oracle scene, with the oracle's intersector standing in for the driver's ray queries.  Buffers ping-pong like
    Renderer::run (frame f uses descriptor set (f+1)%2: last* = [set], this* = [!set])."""

    def __init__(self, R, abi, arrays, osc, size, env=None, sun_sky=None, variant=0):
        self.R, self.abi, self.osc, self.size, self.env = R, abi, osc, size, env
        self.variant = variant      # abi.VARIANT_* bits: the stage whose compile-time switch is flipped runs from the variant build (ref_trace_run_variant)
        w, h = size
        self.tabs = {k: np.ascontiguousarray(osc.table(getattr(abi, k))) for k in ("TABLE_MATERIALS", "TABLE_TRIG_LIGHTS", "TABLE_PUNC_LIGHTS", "TABLE_LIGHT_INFO")}
        # one vertex / index buffer per prim mesh (scene.cpp:209-289), addressed through InstanceData like the shaders do
        self.vbufs = [np.ascontiguousarray(osc.table(abi.TABLE_VERTICES, i)) for i in range(len(arrays.prim_meshes))]
        self.ibufs = [np.ascontiguousarray(osc.table(abi.TABLE_INDICES, i)) for i in range(len(arrays.prim_meshes))]
        geo = np.zeros(len(arrays.prim_meshes), abi.INSTANCE_DT)
        for i, p in enumerate(arrays.prim_meshes):
            geo[i]["vertexAddress"] = self.vbufs[i].ctypes.data
            geo[i]["indexAddress"] = self.ibufs[i].ctypes.data
            geo[i]["materialIndex"] = p["materialIndex"]
        self.geo = geo
        self.xforms = np.zeros((max(1, len(arrays.nodes)), 24), np.float32)      # what the ray query reports per instance
        assert lib().orc_scene_instance_xforms(osc._h, self.xforms.ctypes.data_as(C.c_void_p), len(arrays.nodes)) == len(arrays.nodes)
        self.ss = sun_sky if sun_sky is not None else abi.default_sun_and_sky(in_use=0)
        self.acc = env.accel() if env else np.zeros(1, abi.IMPT_DT)
        self.G = [np.zeros((h, w, 4), np.uint32) for _ in range(2)]
        self.DR = [np.zeros(w * h, abi.DIRECT_RESV_DT) for _ in range(2)]
        self.IR = [np.zeros((w // 2) * (h // 2), abi.INDIRECT_RESV_DT) for _ in range(2)]
        self.tempDR = np.zeros(w * h, abi.DIRECT_RESV_DT)      # tempDirectResv: one buffer, persists across frames (renderer.cpp:235)
        self.motion = np.zeros((h, w, 2), np.int16)
        self.direct, self.indirect, self.indA, self.dirA = (np.zeros((h, w, 4), np.float32) for _ in range(4))
        self.rays = np.zeros(2, np.uint64)

    def run(self, st, frame, direct=True, indirect=True, bufs=None):
        """bufs (replay of the reference's own command sequence, test_reference_renderer_replay...): dict of arrays thisG, lastG, motion,
        thisDR, lastDR, tempDR, thisIR, lastIR, direct, indirect, indA bound explicitly instead of this object's ping-pong."""
        abi, s = self.abi, (frame + 1) % 2
        cam = np.ascontiguousarray(self.osc.table(abi.TABLE_CAMERA))
        b = RefTraceBind()
        b.state, b.camera, b.sunSky, b.lightInfo = C.addressof(st), cam.ctypes.data, C.addressof(self.ss), self.tabs["TABLE_LIGHT_INFO"].ctypes.data
        b.geoInfo, b.materials = self.geo.ctypes.data, self.tabs["TABLE_MATERIALS"].ctypes.data
        b.trigLights, b.puncLights, b.envAccel = self.tabs["TABLE_TRIG_LIGHTS"].ctypes.data, self.tabs["TABLE_PUNC_LIGHTS"].ctypes.data, self.acc.ctypes.data
        b.envSamplerFn = C.cast(lib().orc_sample, C.c_void_p)      # the contract's samplers: environment map (index < 0) and texturesMap[i]
        if self.env:
            b.env, b.envW, b.envH = self.env._h, self.env.w, self.env.h
        b.traceFn, b.scene = C.cast(lib().orc_accel_next_candidate, C.c_void_p), self.osc._h      # the driver's candidate enumeration (contract order)
        b.allocW, b.allocH = self.size
        b.thisG, b.lastG, b.motion = self.G[1 - s].ctypes.data, self.G[s].ctypes.data, self.motion.ctypes.data
        b.thisDR, b.lastDR, b.thisIR, b.lastIR = self.DR[1 - s].ctypes.data, self.DR[s].ctypes.data, self.IR[1 - s].ctypes.data, self.IR[s].ctypes.data
        b.direct, b.indirect, b.indA = self.direct.ctypes.data, self.indirect.ctypes.data, self.indA.ctypes.data
        b.instanceXforms = self.xforms.ctypes.data
        b.tempDR = self.tempDR.ctypes.data
        b.dirA = self.dirA.ctypes.data
        if bufs is not None:
            for k in ("thisG", "lastG", "motion", "thisDR", "lastDR", "tempDR", "thisIR", "lastIR", "direct", "indirect", "indA"):
                setattr(b, k, bufs[k].ctypes.data)
        if not self.variant:
            self.R.ref_trace_run(C.byref(b), int(direct), int(indirect), self.rays.ctypes.data)
        else:       # per stage: the regular build, or the one with DENOISER_DIRECT_BILATERAL (direct) / FETCH_GEOM_CHECK_4_SUBPIXELS (indirect) flipped
            total = np.zeros(2, np.uint64)
            if direct:      # bit 8: direct_gen.comp + direct_reuse.comp (variant build, mode 2); bit 1: direct_stage.comp with DENOISER_DIRECT_BILATERAL
                if self.variant & 8:
                    self.R.ref_trace_run_variant(C.byref(b), 2, 0, self.rays.ctypes.data)
                else:
                    (self.R.ref_trace_run_variant if (self.variant & 1) else self.R.ref_trace_run)(C.byref(b), 1, 0, self.rays.ctypes.data)
                total += self.rays
            if indirect:    # bit 4: indirect_stage.comp with FETCH_GEOM_CHECK_4_SUBPIXELS
                (self.R.ref_trace_run_variant if (self.variant & 4) else self.R.ref_trace_run)(C.byref(b), 0, 1, self.rays.ctypes.data)
                total += self.rays
            self.rays[:] = total
        return {"gbuffer": self.G[1 - s], "motion": self.motion, "direct_resv": self.DR[1 - s], "indirect_resv": self.IR[1 - s],
                "direct": self.direct, "ind_tmp_a": self.indA, "dir_tmp_a": self.dirA}


def trace_setup(scenes, abi, common, cfg):
    """Oracle scene + renderer (+ environment / sun & sky) of a tests/ref_fn_inputs.TRACE_CONFIGS entry."""
    import ref_fn_inputs as fi
    tag, maker_name, size, frames, kind, over = cfg
    arrays = getattr(scenes, maker_name)()
    osc = OracleScene()
    osc.load_arrays(arrays)
    orr = OracleRenderer(osc, size)
    orr.set_env_constant((0.0, 0.0, 0.0))
    env, ss = None, None
    if kind == "hdr":
        env = OracleEnv(fi.ctx_env_image(scenes))
        orr.set_env(env)
    elif kind == "sky":
        ss = fi.sun_sky(abi, fi.CTX_SKY)
        orr.set_sun_and_sky(ss)
    over = fi.trace_state_overrides(common, kind, env.get_integral() if env else None, over)
    osc.update_camera(*size)
    return arrays, osc, orr, env, ss, over


def trace_snapshot(abi, renderer):
    ids = (("gbuffer", abi.BUF_THIS_GBUFFER), ("motion", abi.BUF_MOTION), ("direct_resv", abi.BUF_THIS_DIRECT_RESV),
           ("indirect_resv", abi.BUF_THIS_INDIRECT_RESV), ("direct", abi.BUF_DIRECT), ("ind_tmp_a", abi.BUF_DENOISE_IND_A))
    return {k: renderer.read(w).copy() for k, w in ids}


POST_BUFS = ("BUF_THIS_GBUFFER", "BUF_DIRECT", "BUF_INDIRECT", "BUF_DENOISE_DIR_A", "BUF_DENOISE_DIR_B", "BUF_DENOISE_IND_A", "BUF_DENOISE_IND_B")


def ref_post_run(R, abi, camera_table, state, size, pre):
    """The reference's denoise_direct.comp x4, denoise_indirect.comp x5, compose.comp (oracle/ref_shim/ref_post.cpp) on copies of the
    pre-post buffers `pre` (dict keyed by POST_BUFS) -> dict of the buffers afterwards."""
    bufs = {k: np.ascontiguousarray(pre[k].copy()) for k in POST_BUFS}
    cam = np.ascontiguousarray(camera_table)
    R.ref_post_run(C.addressof(state), cam.ctypes.data, size[0], size[1], *[bufs[k].ctypes.data for k in POST_BUFS])
    return bufs


def ref_post_run_variant(R, abi, camera_table, state, size, pre, variant):
    """Renderer::run's post schedule (renderer.cpp:178-205) with the bilateral switches of `variant` (abi.VARIANT_* bits): a stage whose
    switch is set runs ONCE from the variant build of the reference's shader text (oracle/ref_shim/ref_post.cpp, -DREF_VARIANT), the other
    runs its A-Trous levels from the regular build; compose follows.  pre / result: dicts keyed by POST_BUFS."""
    import copy
    bufs = {k: np.ascontiguousarray(pre[k].copy()) for k in POST_BUFS}
    cam = np.ascontiguousarray(camera_table)
    w, h = size
    ptrs = [bufs[k].ctypes.data for k in POST_BUFS]
    sw, sh = state.size.x, state.size.y

    def disp(fn, stage, st, gw, gh):
        assert fn(stage, C.addressof(st), cam.ctypes.data, w, h, (gw + 7) // 8, (gh + 7) // 8, *ptrs) == 0
    if state.denoise > 0:
        if variant & 1:
            disp(R.ref_post_dispatch_variant, 5, state, sw, sh)
        else:
            for i in range(4):
                st = copy.copy(state); st.denoiseLevel = i
                disp(R.ref_post_dispatch, 5, st, sw, sh)
        if variant & 2:
            disp(R.ref_post_dispatch_variant, 6, state, sw // 2, sh // 2)
        else:
            for i in range(5):
                st = copy.copy(state); st.denoiseLevel = i
                disp(R.ref_post_dispatch, 6, st, sw // 2, sh // 2)
    disp(R.ref_post_dispatch, 7, state, sw, sh)
    return bufs


def mip_chain_average(img):
    """1x1 level of the mip chain RenderOutput::genMipmap blits from an (h, w, 4) float32 image (the contract's linear blits)."""
    a = np.ascontiguousarray(img, np.float32)
    out = np.zeros(4, np.float32)
    lib().orc_mip_chain_average(a.ctypes.data, a.shape[1], a.shape[0], out.ctypes.data)
    return out


def ref_display_run(R, tm, mode, direct, indirect):
    """The reference's post.frag (main() included, oracle/ref_shim/ref_display.cpp) over (h, w, 4) float32 images -> (h, w, 4) float32.  The
    1x1 mip level textureLod(img, vec2(0.5), 20) reads is the driver's blit chain: it is handed in (the contract's value)."""
    d, i = np.ascontiguousarray(direct, np.float32), np.ascontiguousarray(indirect, np.float32)
    h, w = d.shape[:2]
    md, mi = mip_chain_average(d), mip_chain_average(i)
    out = np.zeros((h, w, 4), np.float32)
    R.ref_display_run(C.byref(tm), int(mode), w, h, d.ctypes.data, i.ctypes.data, md.ctypes.data, mi.ctypes.data, out.ctypes.data)
    return out


def display_frames(scenes, abi, common, mode):
    """Oracle renderer after DISPLAY_FRAMES frames of the display-pass scene in debug view `mode` -> (renderer, last RtxState, scene)."""
    import ref_fn_inputs as fi
    size = fi.DISPLAY_SIZE
    osc = OracleScene()
    osc.load_arrays(getattr(scenes, fi.DISPLAY_SCENE)())
    orr = OracleRenderer(osc, size)
    orr.set_env_constant(common.ENV)
    osc.update_camera(*size)
    info = osc.info()
    st = None
    for f in range(fi.DISPLAY_FRAMES):
        osc.update_camera(*size)
        st = common.frame_state(size[0], size[1], info, f, maxDepth=3, debugging_mode=mode)
        orr.run(st, f)
    return orr, st, osc


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class OracleScene:
    def __init__(self, use_bvh=True):
        self._h = C.c_void_p(lib().orc_scene_create())
        lib().orc_scene_set_use_bvh(self._h, int(use_bvh))

    def load_arrays(self, arrays):
        d = arrays.desc()
        assert lib().orc_scene_load_desc(self._h, C.byref(d)) == 0

    def set_lookat(self, eye, center, up, fov_deg):
        lib().orc_scene_set_lookat(self._h, _f3(eye), _f3(center), _f3(up), float(fov_deg))

    def update_camera(self, w, h):
        lib().orc_scene_update_camera(self._h, w, h)

    def set_camera(self, cam):
        lib().orc_scene_set_camera(self._h, C.byref(cam))

    def get_camera(self):
        cam = abi.SceneCamera()
        lib().orc_scene_get_camera(self._h, C.byref(cam))
        return cam

    def info(self):
        i = abi.SceneInfo()
        lib().orc_scene_get_info(self._h, C.byref(i))
        return i

    def table(self, which, index=0):
        n = lib().orc_scene_table_bytes(self._h, which, index)
        assert n >= 0
        buf = np.empty(n, np.uint8)
        if n:
            assert lib().orc_scene_read_table(self._h, which, index, buf.ctypes.data, n) == 0
        return buf.view(abi.TABLE_DTYPES[which])

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros(rays.shape[0], abi.HIT_DT)
        lib().orc_accel_trace(self._h, rays.ctypes.data, rays.shape[0], int(any_hit), hits.ctypes.data)
        return hits

    def __del__(self):
        try:
            lib().orc_scene_destroy(self._h)
        except Exception:
            pass


class OracleEnv:
    def __init__(self, rgba):
        a = np.ascontiguousarray(rgba, np.float32)
        self.h, self.w = a.shape[0], a.shape[1]
        self._h = C.c_void_p(lib().orc_env_create(a.ctypes.data, self.w, self.h))

    def get_integral(self):
        return float(lib().orc_env_integral(self._h))

    def get_average(self):
        return float(lib().orc_env_average(self._h))

    def accel(self):
        out = np.zeros(self.w * self.h, abi.IMPT_DT)
        assert lib().orc_env_read_accel(self._h, out.ctypes.data, out.nbytes) == 0
        return out

    def texture(self, uv):
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), np.float32)
        lib().orc_env_texture(self._h, uv.ctypes.data, uv.shape[0], out.ctypes.data)
        return out

    def __del__(self):
        try:
            lib().orc_env_destroy(self._h)
        except Exception:
            pass


class OracleRenderer:
    def __init__(self, scene, size):
        self._scene = scene
        self.size = tuple(size)
        self._h = C.c_void_p(lib().orc_renderer_create(scene._h, size[0], size[1]))

    def set_env_constant(self, rgb):
        lib().orc_renderer_set_env_constant(self._h, _f3(rgb))

    def set_sun_and_sky(self, ss):
        lib().orc_renderer_set_sun_and_sky(self._h, C.byref(ss))

    def set_variant(self, flags):
        lib().orc_renderer_set_variant(self._h, int(flags))

    def run_output(self, tm, state):
        out = np.zeros((self.size[1], self.size[0], 4), np.float32)
        lib().orc_renderer_run_output(self._h, C.byref(tm), C.byref(state), out.ctypes.data)
        return out

    def set_env(self, env):
        lib().orc_renderer_set_env(self._h, env._h if env is not None else None)
        self._env = env

    def run(self, state, frames):
        lib().orc_renderer_run(self._h, C.byref(state), frames)

    def run_trace(self, state, frames, y0, y1):
        lib().orc_renderer_run_trace(self._h, C.byref(state), frames, y0, y1)

    def run_post(self, state, frames):
        lib().orc_renderer_run_post(self._h, C.byref(state), frames)

    def stats(self):
        s = abi.FrameStats()
        lib().orc_renderer_get_stats(self._h, C.byref(s))
        return s

    def read(self, which):
        n = lib().orc_renderer_buffer_bytes(self._h, which)
        assert n >= 0
        buf = np.empty(n, np.uint8)
        assert lib().orc_renderer_read(self._h, which, buf.ctypes.data, n) == 0
        return buf.view(abi.BUFFER_DTYPES[which])

    def write(self, which, arr):
        a = np.ascontiguousarray(arr)
        assert lib().orc_renderer_write(self._h, which, a.ctypes.data, a.nbytes) == 0

    def __del__(self):
        try:
            lib().orc_renderer_destroy(self._h)
        except Exception:
            pass
