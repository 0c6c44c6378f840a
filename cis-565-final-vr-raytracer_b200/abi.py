"""ctypes / numpy mirrors of include/host_device.h and include/eidola.h.

Layouts follow the reference's shaders/host_device.h (sizes asserted at import time against the
table in SURVEY.md §4).  Pure Python: importing this module needs neither CUDA nor the .so.
"""
import ctypes as C
import numpy as np

c_float_p = C.POINTER(C.c_float)
c_uint32_p = C.POINTER(C.c_uint32)


class Vec2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Vec4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class IVec2(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32)]


class Mat4(C.Structure):
    _fields_ = [("m", C.c_float * 16)]


class SceneCamera(C.Structure):  # host_device.h:153-165
    _fields_ = [("viewInverse", Mat4), ("projInverse", Mat4), ("projView", Mat4), ("lastView", Mat4),
                ("lastProjView", Mat4), ("lastPosition", Vec3), ("nbLights", C.c_int32)]


class RtxState(C.Structure):  # host_device.h:207-238
    _fields_ = [("frame", C.c_int32), ("maxDepth", C.c_int32), ("modulate", C.c_int32),
                ("fireflyClampThreshold", C.c_float), ("hdrMultiplier", C.c_float),
                ("debugging_mode", C.c_int32), ("environmentProb", C.c_float), ("time", C.c_uint32),
                ("ReSTIRState", C.c_int32), ("RISSampleNum", C.c_int32), ("reservoirClamp", C.c_int32),
                ("accumulate", C.c_int32), ("size", IVec2), ("envMapLuminIntegInv", C.c_float),
                ("lightLuminIntegInv", C.c_float), ("MIS", C.c_int32), ("sigLuminDirect", C.c_float),
                ("sigNormalDirect", C.c_float), ("sigDepthDirect", C.c_float), ("denoise", C.c_int32),
                ("sigLuminIndirect", C.c_float), ("sigNormalIndirect", C.c_float),
                ("sigDepthIndirect", C.c_float), ("denoiseLevel", C.c_int32)]


# ReSTIRState / DebugMode enums (host_device.h:128-148)
eNone, eRIS, eSpatial, eTemporal, eSpatiotemporal = range(5)
(eNoDebug, eDirectStage, eIndirectStage, eBaseColor, eNormal, eDepth, eMetallic, eEmissive, eRoughness,
 eTexcoord) = range(10)


def default_rtx_state(width, height, **over):
    """SampleExample::m_rtxState defaults (sample_example.hpp:154-184)."""
    s = RtxState(frame=0, maxDepth=4, modulate=1, fireflyClampThreshold=1.0, hdrMultiplier=1.0,
                 debugging_mode=0, environmentProb=0.25, time=0, ReSTIRState=eTemporal, RISSampleNum=4,
                 reservoirClamp=80, accumulate=0, size=IVec2(width, height), envMapLuminIntegInv=0.0,
                 lightLuminIntegInv=0.0, MIS=1, sigLuminDirect=0.4, sigNormalDirect=0.1,
                 sigDepthDirect=0.02, denoise=1, sigLuminIndirect=4.0, sigNormalIndirect=0.4,
                 sigDepthIndirect=1.0, denoiseLevel=0)
    for k, v in over.items():
        setattr(s, k, v)
    return s


class Vec2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class Tonemapper(C.Structure):  # host_device.h:336-351
    _fields_ = [("brightness", C.c_float), ("contrast", C.c_float), ("saturation", C.c_float), ("vignette", C.c_float),
                ("avgLum", C.c_float), ("zoom", C.c_float), ("renderingRatio", Vec2), ("autoExposure", C.c_int32),
                ("Ywhite", C.c_float), ("key", C.c_float), ("pad", C.c_int32)]


def default_tonemapper(**over):
    """RenderOutput::m_tm defaults (render_output.hpp:44-55)."""
    t = Tonemapper(brightness=1.0, contrast=1.0, saturation=1.0, vignette=0.0, avgLum=1.0, zoom=1.0, renderingRatio=Vec2(1.0, 1.0),
                   autoExposure=0, Ywhite=0.5, key=0.5, pad=0)
    for k, v in over.items():
        setattr(t, k, v)
    return t


class SunAndSky(C.Structure):  # host_device.h:353-376
    _fields_ = [("rgb_unit_conversion", Vec3), ("multiplier", C.c_float), ("haze", C.c_float), ("redblueshift", C.c_float),
                ("saturation", C.c_float), ("horizon_height", C.c_float), ("ground_color", Vec3), ("horizon_blur", C.c_float),
                ("night_color", Vec3), ("sun_disk_intensity", C.c_float), ("sun_direction", Vec3), ("sun_disk_scale", C.c_float),
                ("sun_glow_intensity", C.c_float), ("y_is_up", C.c_int32), ("physically_scaled_sun", C.c_int32), ("in_use", C.c_int32)]


def default_sun_and_sky(**over):
    """SampleExample::m_sunAndSky defaults (sample_example.hpp:186-203); in_use = 0."""
    s = SunAndSky(rgb_unit_conversion=Vec3(1, 1, 1), multiplier=0.0000101320, haze=0.0, redblueshift=0.0, saturation=1.0,
                  horizon_height=0.0, ground_color=Vec3(0.4, 0.4, 0.4), horizon_blur=0.1, night_color=Vec3(0.0, 0.0, 0.01),
                  sun_disk_intensity=0.8, sun_direction=Vec3(0.0, 0.78, 0.62), sun_disk_scale=5.0, sun_glow_intensity=1.0,
                  y_is_up=1, physically_scaled_sun=1, in_use=0)
    for k, v in over.items():
        setattr(s, k, v)
    return s


class PrimMesh(C.Structure):
    _fields_ = [("firstIndex", C.c_uint32), ("indexCount", C.c_uint32), ("vertexOffset", C.c_uint32),
                ("vertexCount", C.c_uint32), ("materialIndex", C.c_int32)]


class Node(C.Structure):
    _fields_ = [("worldMatrix", C.c_float * 16), ("primMesh", C.c_int32)]


class MaterialDesc(C.Structure):
    _fields_ = [("baseColorFactor", C.c_float * 4), ("baseColorTexture", C.c_int32),
                ("metallicFactor", C.c_float), ("roughnessFactor", C.c_float),
                ("metallicRoughnessTexture", C.c_int32), ("emissiveTexture", C.c_int32),
                ("emissiveFactor", C.c_float * 3), ("alphaMode", C.c_int32), ("alphaCutoff", C.c_float),
                ("doubleSided", C.c_int32), ("normalTexture", C.c_int32), ("normalTextureScale", C.c_float),
                ("transmissionFactor", C.c_float), ("transmissionTexture", C.c_int32), ("ior", C.c_float)]


class LightDesc(C.Structure):
    _fields_ = [("worldMatrix", C.c_float * 16), ("type", C.c_int32), ("color", C.c_float * 3),
                ("intensity", C.c_float), ("range", C.c_float), ("innerConeAngle", C.c_float),
                ("outerConeAngle", C.c_float)]


class ImageDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgba8", C.c_void_p)]


class TextureDesc(C.Structure):
    _fields_ = [("image", C.c_int32), ("hasSampler", C.c_int32), ("magFilter", C.c_int32), ("minFilter", C.c_int32),
                ("wrapS", C.c_int32), ("wrapT", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("positions", c_float_p), ("normals", c_float_p), ("tangents", c_float_p),
                ("texcoords0", c_float_p), ("colors0", c_float_p), ("vertexCount", C.c_uint32),
                ("indices", c_uint32_p), ("indexCount", C.c_uint32),
                ("primMeshes", C.POINTER(PrimMesh)), ("primMeshCount", C.c_uint32),
                ("nodes", C.POINTER(Node)), ("nodeCount", C.c_uint32),
                ("materials", C.POINTER(MaterialDesc)), ("materialCount", C.c_uint32),
                ("lights", C.POINTER(LightDesc)), ("lightCount", C.c_uint32),
                ("hasCamera", C.c_int32), ("camEye", C.c_float * 3), ("camCenter", C.c_float * 3),
                ("camUp", C.c_float * 3), ("camYfovRad", C.c_float),
                ("images", C.POINTER(ImageDesc)), ("imageCount", C.c_uint32),
                ("textures", C.POINTER(TextureDesc)), ("textureCount", C.c_uint32)]


class SceneInfo(C.Structure):
    _fields_ = [("primMeshCount", C.c_uint32), ("nodeCount", C.c_uint32), ("materialCount", C.c_uint32),
                ("puncLightCount", C.c_uint32), ("trigLightCount", C.c_uint32), ("vertexCount", C.c_uint32),
                ("indexCount", C.c_uint32), ("triangleInstances", C.c_uint64), ("trigLightWeight", C.c_float),
                ("puncLightWeight", C.c_float), ("bboxMin", C.c_float * 3), ("bboxMax", C.c_float * 3)]


class AccelInfo(C.Structure):
    _fields_ = [("triangleCount", C.c_uint64), ("nodeCount", C.c_uint32), ("maxDepth", C.c_uint32),
                ("nodeBytes", C.c_uint64), ("triBytes", C.c_uint64), ("buildMs", C.c_float),
                ("twoLevel", C.c_int32), ("blasCount", C.c_uint32), ("tlasNodeCount", C.c_uint32), ("instanceCount", C.c_uint32),
                ("fastTrace", C.c_int32)]


ACCEL_AUTO, ACCEL_FLAT, ACCEL_TWO_LEVEL = 0, 1, 2
ACCEL_FAST_TRACE, ACCEL_FAST_BUILD = 0x100, 0x200    # build quality, or-ed into the mode (default: FAST_TRACE, the reference's flag)


EID_K_COUNT = 5
KERNEL_NAMES = ["direct_stage", "indirect_stage", "denoise_direct", "denoise_indirect", "compose"]


class FrameStats(C.Structure):
    _fields_ = [("closestHitRays", C.c_uint64), ("anyHitRays", C.c_uint64), ("primaryHits", C.c_uint64),
                ("launches", C.c_uint32), ("kernelMs", C.c_float * EID_K_COUNT),
                ("kernelLaunches", C.c_uint32 * EID_K_COUNT), ("nodeVisits", C.c_uint64),
                ("triangleTests", C.c_uint64), ("totalClosestHitRays", C.c_uint64), ("totalAnyHitRays", C.c_uint64),
                ("exchangeMs", C.c_float), ("maxNodeVisitsPerThread", C.c_uint64), ("maxNodeVisitsPerQueuedRay", C.c_uint64 * 2)]


VARIANT_DIRECT_BILATERAL, VARIANT_INDIRECT_BILATERAL, VARIANT_FETCH_4_SUBPIXELS, VARIANT_DIRECT_SPLIT = 1, 2, 4, 8     # eid_renderer_set_variant


class GroupInfo(C.Structure):   # eid_group_info
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("y0", C.c_uint32), ("y1", C.c_uint32), ("bandRows", C.c_uint32),
                ("ncclVersion", C.c_int32), ("collectives", C.c_uint64),
                ("stages", C.c_int32), ("nDirect", C.c_int32), ("nIndirect", C.c_int32), ("nPost", C.c_int32),
                ("streamMemOps", C.c_int32), ("pad_", C.c_int32), ("peerCopies", C.c_uint64), ("peerBytes", C.c_uint64)]


class PipelineLayout(C.Structure):   # eid_pipeline_layout
    _fields_ = [("nDirect", C.c_int32), ("nIndirect", C.c_int32), ("nPost", C.c_int32), ("stages", C.c_int32),
                ("index", C.c_int32), ("count", C.c_int32), ("y0", C.c_uint32), ("y1", C.c_uint32), ("paddedHeight", C.c_uint32)]


STAGE_DIRECT, STAGE_INDIRECT, STAGE_POST = 1, 2, 4


# eid_scene_table
(TABLE_MATERIALS, TABLE_PUNC_LIGHTS, TABLE_TRIG_LIGHTS, TABLE_LIGHT_INFO, TABLE_INSTANCE_DATA, TABLE_VERTICES,
 TABLE_INDICES, TABLE_CAMERA, TABLE_TEXELS) = range(9)
# eid_buffer
(BUF_THIS_GBUFFER, BUF_LAST_GBUFFER, BUF_MOTION, BUF_THIS_DIRECT_RESV, BUF_LAST_DIRECT_RESV,
 BUF_THIS_INDIRECT_RESV, BUF_LAST_INDIRECT_RESV, BUF_DIRECT, BUF_INDIRECT, BUF_DENOISE_DIR_A,
 BUF_DENOISE_DIR_B, BUF_DENOISE_IND_A, BUF_DENOISE_IND_B, BUF_DISPLAY_F32, BUF_DISPLAY_RGBA8, BUF_TEMP_DIRECT_RESV) = range(16)

# ---- numpy views of the device tables --------------------------------------------------------------
IMPT_DT = np.dtype([("alias", "<i4"), ("q", "<f4"), ("pdf", "<f4"), ("aliasPdf", "<f4")])
VERTEX_DT = np.dtype([("position", "<f4", 3), ("normal", "<u4"), ("texcoord", "<f4", 2), ("tangent", "<u4"),
                      ("color", "<u4")])
MATERIAL_DT = np.dtype([("pbrBaseColorFactor", "<f4", 4), ("pbrBaseColorTexture", "<i4"),
                        ("pbrMetallicFactor", "<f4"), ("pbrRoughnessFactor", "<f4"),
                        ("pbrMetallicRoughnessTexture", "<i4"), ("emissiveTexture", "<i4"),
                        ("emissiveFactor", "<f4", 3), ("normalTexture", "<i4"), ("normalTextureScale", "<f4"),
                        ("transmissionFactor", "<f4"), ("transmissionTexture", "<i4"), ("ior", "<f4"),
                        ("alphaMode", "<i4"), ("alphaCutoff", "<f4"), ("pad", "<i4")])
PUNC_DT = np.dtype([("type", "<i4"), ("direction", "<f4", 3), ("intensity", "<f4"), ("color", "<f4", 3),
                    ("position", "<f4", 3), ("range", "<f4"), ("outerConeCos", "<f4"), ("innerConeCos", "<f4"),
                    ("padding", "<f4", 2), ("impSamp", IMPT_DT)])
TRIG_DT = np.dtype([("matIndex", "<u4"), ("transformIndex", "<u4"), ("v0", "<f4", 3), ("v1", "<f4", 3),
                    ("v2", "<f4", 3), ("uv0", "<f4", 2), ("uv1", "<f4", 2), ("uv2", "<f4", 2),
                    ("impSamp", IMPT_DT), ("pad", "<f4", 3)])
LIGHTINFO_DT = np.dtype([("puncLightSize", "<u4"), ("trigLightSize", "<u4"), ("trigSampProb", "<f4"),
                         ("pad", "<i4")])
INSTANCE_DT = np.dtype([("vertexAddress", "<u8"), ("indexAddress", "<u8"), ("materialIndex", "<i4"),
                        ("_pad", "<i4")])
LIGHTSAMPLE_DT = np.dtype([("Li", "<f4", 3), ("wi", "<f4", 3), ("dist", "<f4")])
DIRECT_RESV_DT = np.dtype([("lightSample", LIGHTSAMPLE_DT), ("num", "<u4"), ("weight", "<f4")])
GISAMPLE_DT = np.dtype([("L", "<f4", 3), ("xv", "<f4", 3), ("nv", "<f4", 3), ("xs", "<f4", 3),
                        ("ns", "<f4", 3), ("pHat", "<f4")])
INDIRECT_RESV_DT = np.dtype([("giSample", GISAMPLE_DT), ("num", "<u4"), ("weight", "<f4"), ("bigW", "<f4")])
HIT_DT = np.dtype([("hitT", "<f4"), ("primitiveID", "<i4"), ("instanceID", "<i4"),
                   ("instanceCustomIndex", "<i4"), ("baryU", "<f4"), ("baryV", "<f4")])

_SIZES = {SceneCamera: 336, RtxState: 100}
for _t, _n in _SIZES.items():
    assert C.sizeof(_t) == _n, (_t, C.sizeof(_t), _n)
for _dt, _n in ((VERTEX_DT, 32), (MATERIAL_DT, 80), (PUNC_DT, 80), (TRIG_DT, 96), (LIGHTINFO_DT, 16),
                (INSTANCE_DT, 24), (LIGHTSAMPLE_DT, 28), (DIRECT_RESV_DT, 36), (GISAMPLE_DT, 64),
                (INDIRECT_RESV_DT, 76), (IMPT_DT, 16), (HIT_DT, 24)):
    assert _dt.itemsize == _n, (_dt, _dt.itemsize, _n)

TABLE_DTYPES = {TABLE_MATERIALS: MATERIAL_DT, TABLE_PUNC_LIGHTS: PUNC_DT, TABLE_TRIG_LIGHTS: TRIG_DT,
                TABLE_LIGHT_INFO: LIGHTINFO_DT, TABLE_INSTANCE_DATA: INSTANCE_DT, TABLE_VERTICES: VERTEX_DT,
                TABLE_INDICES: np.dtype("<u4"), TABLE_CAMERA: np.dtype("<f4"), TABLE_TEXELS: np.dtype("u1")}
BUFFER_DTYPES = {BUF_THIS_GBUFFER: np.dtype("<u4"), BUF_LAST_GBUFFER: np.dtype("<u4"),
                 BUF_MOTION: np.dtype("<i2"), BUF_THIS_DIRECT_RESV: DIRECT_RESV_DT,
                 BUF_LAST_DIRECT_RESV: DIRECT_RESV_DT, BUF_THIS_INDIRECT_RESV: INDIRECT_RESV_DT,
                 BUF_LAST_INDIRECT_RESV: INDIRECT_RESV_DT, BUF_DIRECT: np.dtype("<f4"),
                 BUF_INDIRECT: np.dtype("<f4"), BUF_DENOISE_DIR_A: np.dtype("<f4"),
                 BUF_DENOISE_DIR_B: np.dtype("<f4"), BUF_DENOISE_IND_A: np.dtype("<f4"),
                 BUF_DENOISE_IND_B: np.dtype("<f4"), BUF_DISPLAY_F32: np.dtype("<f4"), BUF_DISPLAY_RGBA8: np.dtype("u1"),
                 BUF_TEMP_DIRECT_RESV: DIRECT_RESV_DT}


class SceneArrays:
    """Flat scene in the nvh::GltfScene shape (what importDrawableNodes leaves behind) as numpy arrays.

    Keeps every array alive so the ctypes SceneDesc built by .desc() stays valid.
    """

    def __init__(self, positions, normals, tangents, texcoords0, colors0, indices, prim_meshes, nodes,
                 materials, lights=(), camera=None, name="scene", images=(), textures=()):
        f32 = lambda a, k: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, k))
        self.positions, self.normals, self.tangents = f32(positions, 3), f32(normals, 3), f32(tangents, 4)
        self.texcoords0, self.colors0 = f32(texcoords0, 2), f32(colors0, 4)
        self.indices = np.ascontiguousarray(np.asarray(indices, dtype=np.uint32).reshape(-1))
        self.prim_meshes = list(prim_meshes)   # dicts: firstIndex,indexCount,vertexOffset,vertexCount,materialIndex
        self.nodes = list(nodes)               # dicts: worldMatrix(16, column-major), primMesh
        self.materials = list(materials)       # dicts of MaterialDesc fields
        self.lights = list(lights)             # dicts of LightDesc fields
        self.camera = camera                   # dict eye, center, up, yfov (rad) or None
        self.name = name
        self.images = [np.ascontiguousarray(i, np.uint8) for i in images]     # (h, w, 4) uint8 each
        self.textures = list(textures)         # dicts: image, and optionally magFilter/minFilter/wrapS/wrapT (sampler present)
        n = self.positions.shape[0]
        assert self.normals.shape[0] == n and self.tangents.shape[0] == n
        assert self.texcoords0.shape[0] == n and self.colors0.shape[0] == n

    @staticmethod
    def material(base=(1, 1, 1, 1), metallic=1.0, roughness=1.0, emissive=(0, 0, 0), double_sided=0,
                 ior=1.5, transmission=0.0, alpha_mode=0, alpha_cutoff=0.5, base_tex=-1, mr_tex=-1, emissive_tex=-1,
                 normal_tex=-1, normal_scale=1.0, transmission_tex=-1):
        return dict(baseColorFactor=tuple(base), baseColorTexture=base_tex, metallicFactor=metallic,
                    roughnessFactor=roughness, metallicRoughnessTexture=mr_tex, emissiveTexture=emissive_tex,
                    emissiveFactor=tuple(emissive), alphaMode=alpha_mode, alphaCutoff=alpha_cutoff,
                    doubleSided=double_sided, normalTexture=normal_tex, normalTextureScale=normal_scale,
                    transmissionFactor=transmission, transmissionTexture=transmission_tex, ior=ior)

    def desc(self):
        d = SceneDesc()
        self._keep = []
        fp = lambda a: a.ctypes.data_as(c_float_p)
        d.positions, d.normals, d.tangents = fp(self.positions), fp(self.normals), fp(self.tangents)
        d.texcoords0, d.colors0 = fp(self.texcoords0), fp(self.colors0)
        d.vertexCount = self.positions.shape[0]
        d.indices = self.indices.ctypes.data_as(c_uint32_p)
        d.indexCount = self.indices.shape[0]
        pm = (PrimMesh * max(1, len(self.prim_meshes)))()
        for i, p in enumerate(self.prim_meshes):
            pm[i] = PrimMesh(p["firstIndex"], p["indexCount"], p["vertexOffset"], p["vertexCount"], p["materialIndex"])
        nd = (Node * max(1, len(self.nodes)))()
        for i, n in enumerate(self.nodes):
            nd[i].worldMatrix = (C.c_float * 16)(*[float(x) for x in n["worldMatrix"]])
            nd[i].primMesh = n["primMesh"]
        mt = (MaterialDesc * max(1, len(self.materials)))()
        for i, m in enumerate(self.materials):
            x = mt[i]
            x.baseColorFactor = (C.c_float * 4)(*m["baseColorFactor"])
            x.emissiveFactor = (C.c_float * 3)(*m["emissiveFactor"])
            for k in ("baseColorTexture", "metallicFactor", "roughnessFactor", "metallicRoughnessTexture",
                      "emissiveTexture", "alphaMode", "alphaCutoff", "doubleSided", "normalTexture",
                      "normalTextureScale", "transmissionFactor", "transmissionTexture", "ior"):
                setattr(x, k, m[k])
        lt = (LightDesc * max(1, len(self.lights)))()
        for i, l in enumerate(self.lights):
            lt[i].worldMatrix = (C.c_float * 16)(*[float(x) for x in l["worldMatrix"]])
            lt[i].type = l["type"]
            lt[i].color = (C.c_float * 3)(*l["color"])
            lt[i].intensity = l["intensity"]
            lt[i].range = l.get("range", 0.0)
            lt[i].innerConeAngle = l.get("innerConeAngle", 0.0)
            lt[i].outerConeAngle = l.get("outerConeAngle", 0.7853981633974483)
        d.primMeshes, d.primMeshCount = pm, len(self.prim_meshes)
        d.nodes, d.nodeCount = nd, len(self.nodes)
        d.materials, d.materialCount = mt, len(self.materials)
        d.lights, d.lightCount = lt, len(self.lights)
        if self.camera is not None:
            d.hasCamera = 1
            d.camEye = (C.c_float * 3)(*self.camera["eye"])
            d.camCenter = (C.c_float * 3)(*self.camera["center"])
            d.camUp = (C.c_float * 3)(*self.camera["up"])
            d.camYfovRad = self.camera["yfov"]
        im = (ImageDesc * max(1, len(self.images)))()
        for i, a in enumerate(self.images):
            im[i] = ImageDesc(a.shape[1], a.shape[0], a.ctypes.data if a.size else None)
        tx = (TextureDesc * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            has = any(k in t for k in ("magFilter", "minFilter", "wrapS", "wrapT"))
            tx[i] = TextureDesc(t["image"], int(has), t.get("magFilter", -1), t.get("minFilter", -1), t.get("wrapS", 10497), t.get("wrapT", 10497))
        d.images, d.imageCount = im, len(self.images)
        d.textures, d.textureCount = tx, len(self.textures)
        self._keep = [pm, nd, mt, lt, im, tx]
        return d
