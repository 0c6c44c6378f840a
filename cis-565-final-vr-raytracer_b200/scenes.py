"""Synthetic scene generators for the configs of BASELINE.json / SURVEY.md §8(d), and a glTF 2.0 writer.

All geometry is opaque, single-sided, counter-clockwise, with explicit NORMAL / TEXCOORD_0 / TANGENT,
no textures and material ids < 200 (ids whose 8-bit hash is 0xff would alias the sky id,
reference common.glsl:141-143).  Generators are deterministic in their seed (numpy PCG64).

  cube_scene()        C1  unit cube + one KHR_lights_punctual point light
  cornell_scene()     C2  32-triangle Cornell-style box with two single-triangle area lights
  heightfield_room()  C3  closed room, 707x707-quad noise height-field floor (999 698 tris) + 10 wall/ceiling
                          tris + 1000 emissive tris ; C5 = same generator at 2236 quads / 10 000 emissive tris
"""
import base64
import json
import os
import struct

import numpy as np

from .abi import SceneArrays

IDENTITY = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]


def _norm(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


class _Builder:
    """Accumulates prim meshes; each add_* call makes one prim mesh + one node with an identity matrix."""

    def __init__(self):
        self.pos, self.nrm, self.tan, self.uv = [], [], [], []
        self.idx = []
        self.prims, self.nodes, self.materials, self.lights = [], [], [], []
        self.nv = 0
        self.ni = 0

    def add_material(self, **kw):
        self.materials.append(SceneArrays.material(**kw))
        return len(self.materials) - 1

    def add_mesh(self, pos, nrm, tan, uv, idx, material, matrix=None):
        pos = np.asarray(pos, np.float32).reshape(-1, 3)
        idx = np.asarray(idx, np.uint32).reshape(-1)
        self.pos.append(pos)
        self.nrm.append(np.asarray(nrm, np.float32).reshape(-1, 3))
        self.tan.append(np.asarray(tan, np.float32).reshape(-1, 4))
        self.uv.append(np.asarray(uv, np.float32).reshape(-1, 2))
        self.idx.append(idx)
        self.prims.append(dict(firstIndex=self.ni, indexCount=int(idx.size), vertexOffset=self.nv,
                               vertexCount=int(pos.shape[0]), materialIndex=material))
        self.nodes.append(dict(worldMatrix=list(matrix) if matrix is not None else list(IDENTITY),
                               primMesh=len(self.prims) - 1))
        self.nv += pos.shape[0]
        self.ni += idx.size

    def add_quads(self, quads, material, matrix=None):
        """quads: (n,4,3) corners, counter-clockwise seen from the front side."""
        q = np.asarray(quads, np.float64).reshape(-1, 4, 3)
        n = q.shape[0]
        nrm = _norm(np.cross(q[:, 1] - q[:, 0], q[:, 3] - q[:, 0]))
        tan = _norm(q[:, 1] - q[:, 0])
        pos = q.reshape(-1, 3)
        nrm4 = np.repeat(nrm, 4, axis=0)
        tan4 = np.concatenate([np.repeat(tan, 4, axis=0), np.ones((4 * n, 1))], axis=1)
        uv = np.tile(np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float64), (n, 1))
        base = (np.arange(n, dtype=np.uint32) * 4)[:, None]
        idx = (base + np.array([0, 1, 2, 0, 2, 3], np.uint32)[None, :]).reshape(-1)
        self.add_mesh(pos, nrm4, tan4, uv, idx, material, matrix)

    def add_tris(self, tris, material, matrix=None):
        t = np.asarray(tris, np.float64).reshape(-1, 3, 3)
        n = t.shape[0]
        nrm = _norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]))
        tan = _norm(t[:, 1] - t[:, 0])
        pos = t.reshape(-1, 3)
        uv = np.tile(np.array([[0, 0], [1, 0], [0, 1]], np.float64), (n, 1))
        self.add_mesh(pos, np.repeat(nrm, 3, axis=0),
                      np.concatenate([np.repeat(tan, 3, axis=0), np.ones((3 * n, 1))], axis=1), uv,
                      np.arange(3 * n, dtype=np.uint32), material, matrix)

    def build(self, camera, name, images=(), textures=()):
        pos = np.concatenate(self.pos)
        return SceneArrays(pos, np.concatenate(self.nrm), np.concatenate(self.tan), np.concatenate(self.uv),
                           np.ones((pos.shape[0], 4), np.float32), np.concatenate(self.idx), self.prims,
                           self.nodes, self.materials, self.lights, camera, name, images=images, textures=textures)


def _box_quads(lo, hi, faces="xXyYzZ", inward=False):
    """Axis-aligned box faces as CCW quads seen from outside (or from inside when inward)."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    f = {
        "x": [(x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)],   # -x face, outward normal -x
        "X": [(x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1)],   # +x
        "y": [(x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)],   # -y
        "Y": [(x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0)],   # +y
        "z": [(x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0)],   # -z
        "Z": [(x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)],   # +z
    }
    out = []
    for c in faces:
        q = f[c]
        out.append(q[::-1] if inward else q)
    return np.array(out, np.float64)


def cube_scene():
    """C1: unit cube (12 tris / 24 verts), grey dielectric, point light intensity 10 at (2,3,2)."""
    b = _Builder()
    grey = b.add_material(base=(0.8, 0.8, 0.8, 1.0), metallic=0.0, roughness=1.0)
    b.add_quads(_box_quads((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5)), grey)
    m = list(IDENTITY)
    m[12], m[13], m[14] = 2.0, 3.0, 2.0
    b.lights.append(dict(worldMatrix=m, type=1, color=(1.0, 1.0, 1.0), intensity=10.0))
    cam = dict(eye=(2.0, 2.0, -5.0), center=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), yfov=float(np.deg2rad(60.0)))
    return b.build(cam, "c1_cube")


def cornell_scene():
    """C2: 5 walls (10 tris), two open-bottom boxes (20 tris), two single-triangle area lights = 32 tris."""
    b = _Builder()
    white = b.add_material(base=(0.73, 0.73, 0.73, 1), metallic=0.0, roughness=1.0)
    red = b.add_material(base=(0.65, 0.05, 0.05, 1), metallic=0.0, roughness=1.0)
    green = b.add_material(base=(0.12, 0.45, 0.15, 1), metallic=0.0, roughness=1.0)
    metal = b.add_material(base=(0.9, 0.85, 0.6, 1), metallic=1.0, roughness=0.3)
    l0 = b.add_material(base=(0, 0, 0, 1), metallic=0.0, roughness=1.0, emissive=(15.0, 15.0, 15.0))
    l1 = b.add_material(base=(0, 0, 0, 1), metallic=0.0, roughness=1.0, emissive=(10.0, 8.0, 6.0))
    lo, hi = (-1.0, 0.0, -1.0), (1.0, 2.0, 1.0)
    b.add_quads(_box_quads(lo, hi, "yYZ", inward=True), white)     # floor, ceiling, back wall
    b.add_quads(_box_quads(lo, hi, "x", inward=True), red)
    b.add_quads(_box_quads(lo, hi, "X", inward=True), green)
    b.add_quads(_box_quads((-0.7, 0.0, -0.1), (-0.1, 1.2, 0.5), "xXYzZ"), white)   # tall box, no bottom
    b.add_quads(_box_quads((0.15, 0.0, -0.6), (0.7, 0.6, -0.05), "xXYzZ"), metal)  # short box, no bottom
    # area lights just under the ceiling, facing down (normal -y)
    b.add_tris([[(-0.35, 1.98, -0.3), (0.25, 1.98, -0.3), (-0.35, 1.98, 0.3)]], l0)
    b.add_tris([[(0.35, 1.98, 0.45), (0.75, 1.98, 0.45), (0.35, 1.98, 0.85)]], l1)
    cam = dict(eye=(0.0, 1.0, -3.6), center=(0.0, 1.0, 0.0), up=(0.0, 1.0, 0.0), yfov=float(np.deg2rad(45.0)))
    return b.build(cam, "c2_cornell")


def _trs(translate=(0, 0, 0), rot_y_deg=0.0, rot_x_deg=0.0, scale=(1, 1, 1)):
    """Column-major 4x4 = T * Ry * Rx * S (float64 maths, stored as the float32 glTF matrix)."""
    cy, sy = np.cos(np.deg2rad(rot_y_deg)), np.sin(np.deg2rad(rot_y_deg))
    cx, sx = np.cos(np.deg2rad(rot_x_deg)), np.sin(np.deg2rad(rot_x_deg))
    ry = np.array([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1]])
    rx = np.array([[1, 0, 0, 0], [0, cx, -sx, 0], [0, sx, cx, 0], [0, 0, 0, 1]])
    sc = np.diag([scale[0], scale[1], scale[2], 1.0])
    t = np.eye(4)
    t[:3, 3] = translate
    m = t @ ry @ rx @ sc
    return [float(np.float32(v)) for v in m.T.reshape(-1)]


def instanced_scene():
    """Node transforms: a room at identity, ONE unit-box prim mesh instanced by three nodes (rotation + non-uniform scale, a second
    placement, a mirroring transform = negative determinant), and an emissive triangle on a rotated, scaled node."""
    b = _Builder()
    white = b.add_material(base=(0.75, 0.75, 0.75, 1), metallic=0.0, roughness=1.0)
    blue = b.add_material(base=(0.2, 0.3, 0.8, 1), metallic=0.3, roughness=0.5)
    lamp = b.add_material(base=(0, 0, 0, 1), metallic=0.0, roughness=1.0, emissive=(12.0, 11.0, 9.0))
    b.add_quads(_box_quads((-1.5, 0.0, -1.5), (1.5, 2.2, 1.5), "yYZxX", inward=True), white)
    b.add_quads(_box_quads((-0.5, 0.0, -0.5), (0.5, 1.0, 0.5), "xXYzZ"), blue, matrix=_trs((-0.6, 0.0, 0.3), 30.0, 0.0, (0.6, 1.2, 0.5)))
    box = len(b.prims) - 1
    b.nodes.append(dict(worldMatrix=_trs((0.7, 0.0, -0.2), -20.0, 0.0, (0.5, 0.6, 0.7)), primMesh=box))
    b.nodes.append(dict(worldMatrix=_trs((0.1, 0.9, 0.9), 10.0, 25.0, (-0.4, 0.4, 0.4)), primMesh=box))      # mirrored instance
    b.add_tris([[(-0.5, 0.0, -0.5), (0.5, 0.0, -0.5), (-0.5, 0.0, 0.5)]], lamp, matrix=_trs((0.0, 2.15, 0.0), 45.0, 180.0, (0.8, 1.0, 1.3)))
    m = list(IDENTITY)
    m[12], m[13], m[14] = -1.0, 1.6, -1.0
    b.lights.append(dict(worldMatrix=m, type=1, color=(1.0, 0.9, 0.8), intensity=6.0))
    cam = dict(eye=(0.0, 1.1, -4.2), center=(0.0, 0.9, 0.0), up=(0.0, 1.0, 0.0), yfov=float(np.deg2rad(45.0)))
    return b.build(cam, "instanced_box")


def instanced_grid(n=7, seed=11, offset=(0.0, 0.0, 0.0)):
    """Heavy instancing: a room plus ONE noisy-blob prim mesh (12 x 12 x 2 triangles) instanced n x n times with random rotations,
    non-uniform scales and every fourth one mirrored (negative determinant); `offset` moves the whole scene away from the origin (large
    coordinates are the hard case for an object-space walk).  The flat BVH holds n * n copies of the blob, the two-level one a single tree."""
    rng = np.random.Generator(np.random.PCG64(seed))
    b = _Builder()
    ox, oy, oz = offset
    white = b.add_material(base=(0.8, 0.8, 0.8, 1), metallic=0.0, roughness=1.0)
    red = b.add_material(base=(0.8, 0.25, 0.2, 1), metallic=0.2, roughness=0.6)
    lamp = b.add_material(base=(0, 0, 0, 1), metallic=0.0, roughness=1.0, emissive=(14.0, 13.0, 11.0))
    half = 0.9 * n
    b.add_quads(_box_quads((ox - half, oy, oz - half), (ox + half, oy + 3.0, oz + half), "yYZxX", inward=True), white)
    # the blob: a latitude / longitude grid on a unit sphere with radial noise, as independent triangles
    m = 12
    th = np.linspace(0.0, np.pi, m + 1)
    ph = np.linspace(0.0, 2 * np.pi, m + 1)
    rad = 0.35 + 0.12 * rng.random((m + 1, m + 1))
    rad[:, -1] = rad[:, 0]
    rad[0, :] = rad[0, 0]
    rad[-1, :] = rad[-1, 0]
    P = np.stack([rad * np.sin(th)[:, None] * np.cos(ph)[None, :], rad * np.cos(th)[:, None] + 0.5, rad * np.sin(th)[:, None] * np.sin(ph)[None, :]], axis=-1)
    tris = []
    for i in range(m):
        for j in range(m):
            a, bb, c, d = P[i, j], P[i + 1, j], P[i + 1, j + 1], P[i, j + 1]
            for t in ([a, c, bb], [a, d, c]):
                if np.linalg.norm(np.cross(t[1] - t[0], t[2] - t[0])) > 1e-9:       # (the pole rows collapse one edge)
                    tris.append([tuple(v) for v in t])
    first = True
    blob = None
    for i in range(n):
        for j in range(n):
            sx, sy, sz = 0.6 + 0.8 * rng.random(3)
            if (i * n + j) % 4 == 3:
                sx = -sx
            mtx = _trs((ox + (i - (n - 1) / 2) * 1.6, oy + 0.05, oz + (j - (n - 1) / 2) * 1.6), float(rng.uniform(0, 360)), float(rng.uniform(-25, 25)), (sx, sy, sz))
            if first:
                b.add_tris(tris, red, matrix=mtx)
                blob = len(b.prims) - 1
                first = False
            else:
                b.nodes.append(dict(worldMatrix=mtx, primMesh=blob))
    b.add_quads(_box_quads((ox - 0.8, oy + 2.9, oz - 0.8), (ox + 0.8, oy + 2.95, oz + 0.8), "y"), lamp)
    cam = dict(eye=(ox + 0.2, oy + 2.2, oz - 0.85 * n), center=(ox, oy + 0.6, oz), up=(0.0, 1.0, 0.0), yfov=float(np.deg2rad(50.0)))
    return b.build(cam, "instanced_grid")


def _procedural_images(seed=21):
    """Five small RGBA8 images: colour checker, metallic-roughness map, tangent-space normal map, emissive pattern, transmission."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 32
    yy, xx = np.mgrid[0:n, 0:n]
    checker = np.zeros((n, n, 4), np.uint8)
    c = ((xx // 4 + yy // 4) % 2).astype(bool)
    checker[c] = (220, 60, 40, 255)
    checker[~c] = (40, 90, 220, 255)
    checker[..., :3] = np.clip(checker[..., :3].astype(np.int32) + rng.integers(-20, 20, (n, n, 3)), 0, 255)
    mr = np.zeros((n, n, 4), np.uint8)
    mr[..., 1] = (60 + 180 * (xx / (n - 1))).astype(np.uint8)          # roughness ramp in g
    mr[..., 2] = np.where(yy < n // 2, 255, 20)                          # metallic split in b
    mr[..., 3] = 255
    h = np.sin(xx * 0.8) * np.cos(yy * 0.6)
    gx, gy = np.gradient(h)
    nm = np.stack([-gy * 0.8, -gx * 0.8, np.ones_like(h)], axis=-1)
    nm /= np.linalg.norm(nm, axis=-1, keepdims=True)
    normal = np.concatenate([(nm * 0.5 + 0.5) * 255, np.full((n, n, 1), 255.0)], axis=-1).astype(np.uint8)
    emis = np.zeros((16, 16, 4), np.uint8)
    emis[..., 0] = rng.integers(120, 255, (16, 16))
    emis[..., 1] = rng.integers(60, 255, (16, 16))
    emis[..., 2] = rng.integers(30, 200, (16, 16))
    emis[..., 3] = 255
    trans = np.zeros((8, 8, 4), np.uint8)
    trans[..., 0] = rng.integers(0, 255, (8, 8))
    trans[..., 3] = 255
    return [checker, mr, normal, emis, trans]


def textured_scene():
    """Cornell-style box exercising every texture tap of the live path: base colour (sRGB), metallic-roughness, normal map,
    emissive map on an area light, transmission map, with LINEAR/NEAREST filters and REPEAT/MIRRORED/CLAMP wrap modes."""
    b = _Builder()
    images = _procedural_images()
    textures = [dict(image=0),                                                        # 0: no sampler -> LINEAR / REPEAT
                dict(image=1, magFilter=9729, minFilter=9987, wrapS=33648, wrapT=10497),  # 1: LINEAR, mirrored / repeat
                dict(image=2, magFilter=9729, minFilter=9729, wrapS=10497, wrapT=10497),  # 2: normal map, LINEAR / REPEAT
                dict(image=3, magFilter=9728, minFilter=9728, wrapS=33071, wrapT=33071),  # 3: NEAREST, clamp
                dict(image=4, magFilter=-1, minFilter=-1, wrapS=10497, wrapT=33648),      # 4: sampler without filters -> NEAREST
                dict(image=7)]                                                            # 5: bad source -> white default texture
    floor = b.add_material(base=(1, 1, 1, 1), metallic=1.0, roughness=1.0, base_tex=0, mr_tex=1, normal_tex=2, normal_scale=0.8)
    wall = b.add_material(base=(0.9, 0.9, 0.9, 1), metallic=0.0, roughness=0.9, base_tex=0)
    bumpy = b.add_material(base=(0.8, 0.7, 0.3, 1), metallic=0.6, roughness=0.4, normal_tex=2, normal_scale=1.5, transmission=0.7, transmission_tex=4)
    plain = b.add_material(base=(0.7, 0.7, 0.7, 1), metallic=0.0, roughness=1.0, base_tex=5)
    lamp = b.add_material(base=(0, 0, 0, 1), metallic=0.0, roughness=1.0, emissive=(14.0, 14.0, 14.0), emissive_tex=3)
    lamp2 = b.add_material(base=(0, 0, 0, 1), metallic=0.0, roughness=1.0, emissive=(6.0, 7.0, 9.0))
    lo, hi = (-1.0, 0.0, -1.0), (1.0, 2.0, 1.0)
    b.add_quads(_box_quads(lo, hi, "y", inward=True), floor)
    b.add_quads(_box_quads(lo, hi, "YZ", inward=True), plain)
    b.add_quads(_box_quads(lo, hi, "xX", inward=True), wall)
    b.add_quads(_box_quads((-0.6, 0.0, -0.2), (-0.05, 1.0, 0.4), "xXYzZ"), bumpy)
    b.add_quads(_box_quads((0.2, 0.0, -0.6), (0.7, 0.5, -0.1), "xXYzZ"), floor)
    b.add_quads([[(-0.4, 1.97, -0.3), (0.3, 1.97, -0.3), (0.3, 1.97, 0.3), (-0.4, 1.97, 0.3)]], lamp)      # faces down, textured emitter
    b.add_tris([[(0.4, 1.98, 0.45), (0.8, 1.98, 0.45), (0.4, 1.98, 0.85)]], lamp2)
    # scale uvs of the floor so the wrap modes are exercised (values outside [0,1] and negative)
    b.uv[0] = b.uv[0] * 3.0 - 1.0
    cam = dict(eye=(0.0, 1.0, -3.6), center=(0.0, 1.0, 0.0), up=(0.0, 1.0, 0.0), yfov=float(np.deg2rad(45.0)))
    return b.build(cam, "textured_box", images=images, textures=textures)


def alpha_scene():
    """textured_scene() plus alpha-tested (MASK, texture alpha) and blended (BLEND, factor + texture alpha) panes in front of the
    boxes, one of them double sided: exercises HitTest (stochastic alpha) for primary, shadow and bounce rays."""
    base = textured_scene()
    b = _Builder()
    rng = np.random.Generator(np.random.PCG64(77))
    leaf = np.zeros((16, 16, 4), np.uint8)
    leaf[..., :3] = (60, 170, 70)
    yy, xx = np.mgrid[0:16, 0:16]
    leaf[..., 3] = np.where(((xx - 8) ** 2 + (yy - 8) ** 2) < 40, 255, 0)      # disc cut-out
    leaf[..., 3] = np.where(rng.random((16, 16)) < 0.1, 128, leaf[..., 3])
    smoke = np.zeros((8, 8, 4), np.uint8)
    smoke[..., :3] = (200, 200, 210)
    smoke[..., 3] = rng.integers(30, 230, (8, 8))
    images = list(base.images) + [leaf, smoke]
    textures = list(base.textures) + [dict(image=5, magFilter=9728, minFilter=9728, wrapS=10497, wrapT=10497),   # 6: leaf, NEAREST
                                      dict(image=6)]                                                              # 7: smoke, LINEAR
    b.materials = list(base.materials)
    mask = b.add_material(base=(1, 1, 1, 1), metallic=0.0, roughness=0.8, base_tex=6, alpha_mode=1, alpha_cutoff=0.5, double_sided=1)
    blend = b.add_material(base=(0.9, 0.9, 1.0, 0.6), metallic=0.0, roughness=0.5, base_tex=7, alpha_mode=2)
    blend_flat = b.add_material(base=(1.0, 0.8, 0.8, 0.35), metallic=0.0, roughness=0.9, alpha_mode=2, double_sided=1)
    # panes between the camera and the boxes (camera looks down +z from z = -3.6)
    b.add_quads([[(-0.9, 0.1, -0.7), (-0.9, 1.3, -0.7), (-0.1, 1.3, -0.7), (-0.1, 0.1, -0.7)]], mask)          # faces the camera (-z)
    b.add_quads([[(0.1, 0.2, -0.8), (0.1, 1.0, -0.8), (0.85, 1.0, -0.8), (0.85, 0.2, -0.8)]], blend)
    b.add_quads([[(-0.5, 1.4, -0.3), (0.5, 1.4, -0.3), (0.5, 1.4, 0.5), (-0.5, 1.4, 0.5)]], blend_flat)         # horizontal, under the lamp: shadow rays cross it
    extra = b.build(base.camera, "alpha_box", images=images, textures=textures)
    # merge: base geometry first, then the panes (vertex / index offsets shift)
    nv, ni = base.positions.shape[0], base.indices.shape[0]
    prims = list(base.prim_meshes) + [dict(p, firstIndex=p["firstIndex"] + ni, vertexOffset=p["vertexOffset"] + nv) for p in extra.prim_meshes]
    nodes = list(base.nodes) + [dict(n, primMesh=n["primMesh"] + len(base.prim_meshes)) for n in extra.nodes]
    cat = lambda a, c: np.concatenate([a, c])
    return SceneArrays(cat(base.positions, extra.positions), cat(base.normals, extra.normals), cat(base.tangents, extra.tangents),
                       cat(base.texcoords0, extra.texcoords0), cat(base.colors0, extra.colors0), cat(base.indices, extra.indices),
                       prims, nodes, extra.materials, base.lights, base.camera, "alpha_box", images=images, textures=textures)


def _value_noise(x, z, seed, octaves=4, base_cells=8):
    """Sum of `octaves` smooth value-noise layers over the unit square (x,z in [0,1])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.zeros_like(x, dtype=np.float64)
    amp, cells, total = 1.0, base_cells, 0.0
    for _ in range(octaves):
        lat = rng.random((cells + 2, cells + 2))
        fx, fz = x * cells, z * cells
        ix = np.minimum(fx.astype(np.int64), cells - 1)
        iz = np.minimum(fz.astype(np.int64), cells - 1)
        tx, tz = fx - ix, fz - iz
        sx, sz = tx * tx * (3 - 2 * tx), tz * tz * (3 - 2 * tz)
        v = (lat[iz, ix] * (1 - sx) + lat[iz, ix + 1] * sx) * (1 - sz) + (lat[iz + 1, ix] * (1 - sx) + lat[iz + 1, ix + 1] * sx) * sz
        out += amp * (v - 0.5)
        total += amp
        amp *= 0.5
        cells *= 2
    return out / total


def heightfield_room(quads=707, n_light_quads=500, seed=565, light_seed=566, room=(40.0, 12.0, 40.0),
                     amplitude=0.5, patches=(4, 3)):
    """C3 (defaults) / C5 (quads=2236, n_light_quads=5000, light_seed=567)."""
    b = _Builder()
    W, H, D = room
    x0, z0 = -W / 2, -D / 2
    # 12 floor materials (dielectric / rough metal mix), then wall material, then 8 emissive materials
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    floor_mats = []
    for i in range(patches[0] * patches[1]):
        base = 0.25 + 0.6 * rng.random(3)
        floor_mats.append(b.add_material(base=(float(base[0]), float(base[1]), float(base[2]), 1.0),
                                         metallic=float(i % 3 == 2) * 0.8, roughness=float(0.35 + 0.6 * rng.random())))
    wall = b.add_material(base=(0.7, 0.7, 0.72, 1.0), metallic=0.0, roughness=0.9)
    emis = []
    for i in range(8):
        e = 2.0 + 18.0 * rng.random()
        tint = 0.7 + 0.3 * rng.random(3)
        emis.append(b.add_material(base=(0, 0, 0, 1), metallic=0.0, roughness=1.0,
                                   emissive=(float(e * tint[0]), float(e * tint[1]), float(e * tint[2]))))
    # height field on a (quads+1)^2 lattice
    g = np.linspace(0.0, 1.0, quads + 1)
    gx, gz = np.meshgrid(g, g)                      # [iz, ix]
    hgt = amplitude * 2.0 * _value_noise(gx, gz, seed)
    px = x0 + gx * W
    pz = z0 + gz * D
    # normals from central differences of the displaced surface
    dx, dz = W / quads, D / quads
    hx = np.gradient(hgt, dx, axis=1)
    hz = np.gradient(hgt, dz, axis=0)
    nrm = _norm(np.stack([-hx, np.ones_like(hx), -hz], axis=-1))
    tan = np.stack([np.ones_like(hx), hx, np.zeros_like(hx)], axis=-1)
    tan = _norm(tan - nrm * np.sum(tan * nrm, axis=-1, keepdims=True))
    xs = np.linspace(0, quads, patches[0] + 1).astype(int)
    zs = np.linspace(0, quads, patches[1] + 1).astype(int)
    k = 0
    for pz_i in range(patches[1]):
        for px_i in range(patches[0]):
            ix0, ix1, iz0, iz1 = xs[px_i], xs[px_i + 1], zs[pz_i], zs[pz_i + 1]
            sl = (slice(iz0, iz1 + 1), slice(ix0, ix1 + 1))
            nxv = ix1 - ix0 + 1
            pos = np.stack([px[sl], hgt[sl], pz[sl]], axis=-1).reshape(-1, 3)
            uv = np.stack([gx[sl], gz[sl]], axis=-1).reshape(-1, 2)
            tn = np.concatenate([tan[sl].reshape(-1, 3), np.ones((pos.shape[0], 1))], axis=1)
            jz, jx = np.meshgrid(np.arange(iz1 - iz0), np.arange(ix1 - ix0), indexing="ij")
            v00 = (jz * nxv + jx).reshape(-1)
            v10 = v00 + 1
            v01 = v00 + nxv
            v11 = v01 + 1
            idx = np.stack([v00, v01, v10, v10, v01, v11], axis=1).reshape(-1).astype(np.uint32)
            b.add_mesh(pos, nrm[sl].reshape(-1, 3), tn, uv, idx, floor_mats[k])
            k += 1
    # walls + ceiling (10 tris), extending below the floor so the room is closed
    b.add_quads(_box_quads((x0, -amplitude - 0.5, z0), (x0 + W, H, z0 + D), "xXzZY", inward=True), wall)
    # emissive quads 0.2 x 0.2 facing down, just under the ceiling, grouped by material
    lrng = np.random.Generator(np.random.PCG64(light_seed))
    cx = x0 + 1.0 + (W - 2.0) * lrng.random(n_light_quads)
    cz = z0 + 1.0 + (D - 2.0) * lrng.random(n_light_quads)
    cy = H - 0.05 - 1.5 * lrng.random(n_light_quads)
    which = lrng.integers(0, 8, n_light_quads)
    h = 0.1
    for m in range(8):
        sel = np.nonzero(which == m)[0]
        if sel.size == 0:
            continue
        q = np.stack([
            np.stack([cx[sel] - h, cy[sel], cz[sel] - h], axis=-1),
            np.stack([cx[sel] + h, cy[sel], cz[sel] - h], axis=-1),
            np.stack([cx[sel] + h, cy[sel], cz[sel] + h], axis=-1),
            np.stack([cx[sel] - h, cy[sel], cz[sel] + h], axis=-1)], axis=1)
        b.add_quads(q, emis[m])   # (p1-p0)x(p3-p0) = (+x) x (+z) = -y : faces down
    cam = dict(eye=(0.0, 5.0, -17.5), center=(0.0, 1.5, 4.0), up=(0.0, 1.0, 0.0), yfov=float(np.deg2rad(60.0)))
    return b.build(cam, "c3_room_%dq_%dl" % (quads, 2 * n_light_quads))


def stress_room():
    """C5: 2236x2236 quads (~10.0 M tris) + 10 000 emissive tris."""
    return heightfield_room(quads=2236, n_light_quads=5000, seed=565, light_seed=567)


def small_room(quads=48, n_light_quads=12, seed=7):
    """Miniature of C3 for parity tests the oracle finishes in seconds."""
    return heightfield_room(quads=quads, n_light_quads=n_light_quads, seed=seed, light_seed=seed + 1,
                            room=(8.0, 4.0, 8.0), amplitude=0.25, patches=(2, 2))


# ------------------------------------------------------------------------------------------------
# glTF 2.0 writer (.gltf + external .bin, or a single file with a base64 data: URI)
# ------------------------------------------------------------------------------------------------
def _camera_matrix(eye, center, up):
    eye, center, up = (np.asarray(v, np.float64) for v in (eye, center, up))
    f = _norm(center - eye)
    s = _norm(np.cross(f, up))
    u = np.cross(s, f)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = s, u, -f, eye
    return [float(x) for x in m.T.reshape(-1)]   # column-major


def write_gltf(scene, path, embed=False):
    """Write `scene` (SceneArrays) as glTF 2.0; one mesh + node per prim mesh, KHR_lights_punctual nodes,
    one camera node.  Returns the path."""
    chunks, views, accessors = [], [], []
    off = 0

    def add(arr, target, typ, ctype, minmax=False):
        nonlocal off
        data = np.ascontiguousarray(arr).tobytes()
        pad = (-len(data)) % 4
        views.append(dict(buffer=0, byteOffset=off, byteLength=len(data), target=target))
        acc = dict(bufferView=len(views) - 1, componentType=ctype, count=int(arr.shape[0]), type=typ)
        if minmax:
            acc["min"] = [float(x) for x in arr.min(axis=0)]
            acc["max"] = [float(x) for x in arr.max(axis=0)]
        accessors.append(acc)
        chunks.append(data + b"\0" * pad)
        off += len(data) + pad
        return len(accessors) - 1

    meshes, nodes = [], []
    vert_cache = {}
    for pi, pm in enumerate(scene.prim_meshes):
        key = (pm["vertexOffset"], pm["vertexCount"])
        if key not in vert_cache:
            sl = slice(pm["vertexOffset"], pm["vertexOffset"] + pm["vertexCount"])
            vert_cache[key] = dict(
                POSITION=add(scene.positions[sl], 34962, "VEC3", 5126, True),
                NORMAL=add(scene.normals[sl], 34962, "VEC3", 5126),
                TANGENT=add(scene.tangents[sl], 34962, "VEC4", 5126),
                TEXCOORD_0=add(scene.texcoords0[sl], 34962, "VEC2", 5126))
        ia = add(scene.indices[pm["firstIndex"]:pm["firstIndex"] + pm["indexCount"]], 34963, "SCALAR", 5125)
        meshes.append(dict(primitives=[dict(attributes=dict(vert_cache[key]), indices=ia,
                                            material=pm["materialIndex"], mode=4)]))
    for nd in scene.nodes:
        n = dict(mesh=nd["primMesh"])
        if list(nd["worldMatrix"]) != IDENTITY:
            n["matrix"] = [float(x) for x in nd["worldMatrix"]]
        nodes.append(n)
    mats = []
    for m in scene.materials:
        g = dict(pbrMetallicRoughness=dict(baseColorFactor=[float(x) for x in m["baseColorFactor"]],
                                           metallicFactor=float(m["metallicFactor"]),
                                           roughnessFactor=float(m["roughnessFactor"])),
                 emissiveFactor=[float(x) for x in m["emissiveFactor"]],
                 doubleSided=bool(m["doubleSided"]),
                 alphaMode=["OPAQUE", "MASK", "BLEND"][m["alphaMode"]], alphaCutoff=float(m["alphaCutoff"]))
        ext = {}
        if m["ior"] != 1.5:
            ext["KHR_materials_ior"] = dict(ior=float(m["ior"]))
        if m["transmissionFactor"] != 0.0:
            ext["KHR_materials_transmission"] = dict(transmissionFactor=float(m["transmissionFactor"]))
        if ext:
            g["extensions"] = ext
        mats.append(g)
    doc = dict(asset=dict(version="2.0", generator="eidola-b200 scenes.py"), scene=0, meshes=meshes,
               materials=mats, accessors=accessors, bufferViews=views)
    ext_used = []
    if scene.lights:
        ext_used.append("KHR_lights_punctual")
        ls = []
        for li, l in enumerate(scene.lights):
            ls.append(dict(type=["directional", "point", "spot"][l["type"]], color=[float(x) for x in l["color"]],
                           intensity=float(l["intensity"])))
            nodes.append(dict(matrix=[float(x) for x in l["worldMatrix"]],
                              extensions=dict(KHR_lights_punctual=dict(light=li))))
        doc["extensions"] = dict(KHR_lights_punctual=dict(lights=ls))
    if scene.camera is not None:
        c = scene.camera
        doc["cameras"] = [dict(type="perspective", perspective=dict(yfov=float(c["yfov"]), znear=0.001, zfar=1000.0,
                                                                    aspectRatio=16.0 / 9.0))]
        nodes.append(dict(camera=0, matrix=_camera_matrix(c["eye"], c["center"], c["up"])))
    if any("KHR_materials_ior" in m.get("extensions", {}) for m in mats):
        ext_used.append("KHR_materials_ior")
    if any("KHR_materials_transmission" in m.get("extensions", {}) for m in mats):
        ext_used.append("KHR_materials_transmission")
    if ext_used:
        doc["extensionsUsed"] = ext_used
    doc["nodes"] = nodes
    doc["scenes"] = [dict(nodes=list(range(len(nodes))))]
    blob = b"".join(chunks)
    if embed:
        doc["buffers"] = [dict(byteLength=len(blob), uri="data:application/octet-stream;base64," + base64.b64encode(blob).decode())]
    else:
        bin_name = os.path.splitext(os.path.basename(path))[0] + ".bin"
        with open(os.path.join(os.path.dirname(path) or ".", bin_name), "wb") as f:
            f.write(blob)
        doc["buffers"] = [dict(byteLength=len(blob), uri=bin_name)]
    with open(path, "w") as f:
        json.dump(doc, f)
    return path


def synthetic_sky(width=64, height=32, seed=9, sun=True):
    """Synthetic RGBA32F lat-long environment: blue-ish gradient sky, dark ground, a small very bright 'sun' block."""
    rng = np.random.Generator(np.random.PCG64(seed))
    v = (np.arange(height, dtype=np.float64) + 0.5) / height          # 0 = +Y pole
    sky = np.clip(1.0 - v * 1.6, 0.02, 1.0)[:, None]
    img = np.zeros((height, width, 4), np.float32)
    img[..., 0] = 0.35 * sky + 0.05
    img[..., 1] = 0.55 * sky + 0.06
    img[..., 2] = 0.95 * sky + 0.08
    img[..., :3] *= (0.8 + 0.4 * rng.random((height, width, 1))).astype(np.float32)
    if sun:
        y0, x0 = height // 5, (2 * width) // 3
        img[y0:y0 + 2, x0:x0 + 2, :3] = (220.0, 200.0, 160.0)
    img[..., 3] = 1.0
    return img


def write_radiance_hdr(path, rgba, rle=True):
    """Write RGBA32F (h, w, 4) as a Radiance .hdr (RGBE); returns the texels a decoder will reproduce."""
    rgb = np.asarray(rgba, np.float32)[..., :3]
    h, w, _ = rgb.shape
    m = rgb.max(axis=-1)
    e = np.where(m > 1e-32, np.ceil(np.log2(np.maximum(m, 1e-38))).astype(np.int32), -128)
    e = np.where(np.ldexp(1.0, e) <= m, e + 1, e)                       # ensure m / 2^e < 1
    scale = np.where(m > 1e-32, np.ldexp(256.0, -e), 0.0)
    mant = np.clip((rgb * scale[..., None]).astype(np.int32), 0, 255).astype(np.uint8)
    rgbe = np.concatenate([mant, np.where(m > 1e-32, e + 128, 0).astype(np.uint8)[..., None]], axis=-1)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w))
        for y in range(h):
            if rle and 8 <= w < 32768:
                f.write(bytes([2, 2, (w >> 8) & 0xff, w & 0xff]))
                for c in range(4):
                    row = rgbe[y, :, c]
                    x = 0
                    while x < w:                                        # literal runs only (valid, simple)
                        n = min(128, w - x)
                        f.write(bytes([n]) + row[x:x + n].tobytes())
                        x += n
            else:
                f.write(rgbe[y].tobytes())
    dec = np.zeros((h, w, 4), np.float32)
    s = np.where(rgbe[..., 3] > 0, np.ldexp(1.0, rgbe[..., 3].astype(np.int32) - 136), 0.0).astype(np.float32)
    dec[..., :3] = rgbe[..., :3].astype(np.float32) * s[..., None]
    dec[..., 3] = 1.0
    return dec
