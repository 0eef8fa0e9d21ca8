"""Synthetic scenes for the BASELINE.json configs.

A scene is pure input data (numpy buffers + pipeline state + a draw list) that
is replayed, call for call, through any library exporting the C ABI of
include/slcu.h: the CUDA product, the compiled reference or the CPU
restatement.  All randomness is a seeded numpy Generator; all matrices are
column-major float32 like ls::math::mat4.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

from . import capi as A


# ------------------------------------------------------------------ matrices
def perspective(fovy_deg: float, aspect: float, znear: float, zfar: float) -> np.ndarray:
    f = 1.0 / np.tan(np.radians(fovy_deg) / 2.0)
    m = np.zeros((4, 4), dtype=np.float64)  # m[col][row]
    m[0][0] = f / aspect
    m[1][1] = f
    m[2][2] = (zfar + znear) / (znear - zfar)
    m[2][3] = -1.0
    m[3][2] = (2.0 * zfar * znear) / (znear - zfar)
    return m.astype(np.float32)


def infinite_perspective(fovy_deg: float, aspect: float, znear: float) -> np.ndarray:
    """Reversed-Z style infinite projection (z_ndc -> 0 at infinity, 1 at near)."""
    f = 1.0 / np.tan(np.radians(fovy_deg) / 2.0)
    m = np.zeros((4, 4), dtype=np.float64)
    m[0][0] = f / aspect
    m[1][1] = f
    m[2][2] = 0.0
    m[2][3] = -1.0
    m[3][2] = znear
    return m.astype(np.float32)


def look_at(eye, target, up) -> np.ndarray:
    eye = np.asarray(eye, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    up = np.asarray(up, dtype=np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    s = np.cross(fwd, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, fwd)
    m = np.identity(4, dtype=np.float64)
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -fwd
    m[3][0] = -np.dot(s, eye)
    m[3][1] = -np.dot(u, eye)
    m[3][2] = np.dot(fwd, eye)
    return m.astype(np.float32)


def matmul_cm(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Column-major product a*b (both stored m[col][row])."""
    return (b.astype(np.float64) @ a.astype(np.float64)).astype(np.float32)


def translate_scale(t, s) -> np.ndarray:
    m = np.identity(4, dtype=np.float32)
    m[0][0] = m[1][1] = m[2][2] = s
    m[3][0:3] = t
    return m


# ------------------------------------------------------------------ geometry
def uv_sphere(stacks: int, slices: int, radius: float = 1.0):
    """(stacks+1)*(slices+1) vertices, stacks*slices*2 triangles (CCW seen from outside)."""
    v = np.linspace(0.0, np.pi, stacks + 1)
    u = np.linspace(0.0, 2.0 * np.pi, slices + 1)
    vv, uu = np.meshgrid(v, u, indexing="ij")
    x = np.sin(vv) * np.cos(uu)
    y = np.cos(vv)
    z = np.sin(vv) * np.sin(uu)
    nrm = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    pos = nrm * radius
    uv = np.stack([uu / (2.0 * np.pi), vv / np.pi], axis=-1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(stacks), np.arange(slices), indexing="ij")
    a = (i * (slices + 1) + j).reshape(-1)
    b = a + 1
    c = a + (slices + 1)
    d = c + 1
    tris = np.stack([a, b, c, b, d, c], axis=-1).reshape(-1, 3)
    return pos.astype(np.float32), nrm.astype(np.float32), uv.astype(np.float32), tris.astype(np.uint32)


def value_noise_texture(seed: int, w: int, h: int, channels: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h // 8 + 1, w // 8 + 1, channels)).astype(np.float32)
    ys = np.arange(h) / 8.0
    xs = np.arange(w) / 8.0
    y0 = ys.astype(int)
    x0 = xs.astype(int)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    t = (base[y0][:, x0] * (1 - fx) * (1 - fy) + base[y0][:, x0 + 1] * fx * (1 - fy)
         + base[y0 + 1][:, x0] * (1 - fx) * fy + base[y0 + 1][:, x0 + 1] * fx * fy)
    t = t + rng.integers(-12, 13, size=t.shape)
    return np.clip(t, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------- scene
@dataclass
class Scene:
    name: str
    width: int
    height: int
    color_type: int
    depth_type: int
    clear_color: tuple
    clear_depth: float
    setup: Callable  # (ctx) -> state dict
    frame: Callable  # (ctx, state) -> None   (clear + all draws of one frame)
    tris_per_frame: int = 0
    draws_per_frame: int = 1
    note: str = ""
    upload_bytes_per_frame: int = 0
    per_frame_update: Optional[Callable] = None  # (ctx, state, frame_idx) -> bytes uploaded


class SceneInstance:
    """A scene bound to one context (one library)."""

    def __init__(self, scene: Scene, ctx: A.Context):
        self.scene = scene
        self.ctx = ctx
        self.state = scene.setup(ctx)

    def render(self):
        self.scene.frame(self.ctx, self.state)

    def read_color(self) -> np.ndarray:
        s = self.scene
        return self.ctx.texture_read(self.state["color"], s.color_type, s.width, s.height)

    def read_depth(self) -> np.ndarray:
        s = self.scene
        return self.ctx.texture_read(self.state["depth"], s.depth_type, s.width, s.height)[..., 0]

    def read_colors(self):
        """Every colour attachment (render targets 1.. exist for MRT scenes only)."""
        s = self.scene
        out = [self.read_color()]
        for tex, t in zip(self.state.get("extra", []), self.state.get("extra_types", [])):
            out.append(self.ctx.texture_read(tex, t, s.width, s.height))
        return out


def _make_targets(ctx: A.Context, scene_w, scene_h, color_type, depth_type):
    color = ctx.create_texture(color_type, scene_w, scene_h, 1)
    depth = ctx.create_texture(depth_type, scene_w, scene_h, 1)
    fbo = ctx.create_framebuffer([color], depth)
    assert ctx.fbo_valid(fbo) == 0
    return color, depth, fbo


# ---- draw-test triangle (softlight/tests/sl_draw_test.cpp:85-197) -----------
def draw_test_scene(width=640, height=480, depth_type=A.SL_COLOR_R_FLOAT) -> Scene:
    tri = np.array([[-0.5, -0.5, 0, 1, 1, 0, 0, 1],
                    [0.0, 0.5, 0, 1, 0, 1, 0, 1],
                    [0.5, -0.5, 0, 1, 0, 0, 1, 1]], dtype=np.float32)
    idx = np.array([0, 2, 1], dtype=np.uint8)

    def setup(ctx):
        color, depth, fbo = _make_targets(ctx, width, height, A.SL_COLOR_RGBA_8U, depth_type)
        vbo = ctx.create_vbo(tri)
        ibo = ctx.create_ibo(idx)
        vao = ctx.create_vao(vbo, ibo, [(0, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT),
                                        (16, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT)])
        ubo = ctx.create_ubo(b"\0" * 16)
        sh = ctx.create_shader("vcolor", "vcolor", ubo, vs_varyings=1, fs_varyings=1, cull=A.SL_CULL_BACK_FACE,
                               blend=A.SL_BLEND_OFF, depth_test=A.SL_DEPTH_TEST_OFF, depth_mask=A.SL_DEPTH_MASK_OFF)
        assert sh != A.INVALID_ID
        return dict(color=color, depth=depth, fbo=fbo, shader=sh, mesh=ctx.mesh(vao, 0, 3))

    def frame(ctx, st):
        ctx.clear_framebuffer(st["fbo"], 0, (1.0, 0.0, 1.0, 1.0), 0.0)
        ctx.draw(st["mesh"], st["shader"], st["fbo"])

    return Scene("draw_test", width, height, A.SL_COLOR_RGBA_8U, depth_type, (1, 0, 1, 1), 0.0, setup, frame,
                 tris_per_frame=1)


# ---- C1: 100k-triangle UV sphere, Lambert -----------------------------------
def sphere_scene(width=1024, height=768, stacks=224, slices=224, depth_type=A.SL_COLOR_R_FLOAT,
                 color_type=A.SL_COLOR_RGBA_8U, depth_test=A.SL_DEPTH_TEST_LESS_THAN, cull=A.SL_CULL_BACK_FACE,
                 blend=A.SL_BLEND_OFF, eye=(0.0, 0.0, 2.2), name=None, clear_depth=1.0, reps=1) -> Scene:
    pos, nrm, _, tris = uv_sphere(stacks, slices)
    verts = np.zeros((pos.shape[0], 8), dtype=np.float32)
    verts[:, 0:3] = pos
    verts[:, 3] = 1.0
    verts[:, 4:7] = nrm
    mvp = matmul_cm(perspective(60.0, width / height, 0.1, 100.0), look_at(eye, (0, 0, 0), (0, 1, 0)))
    light = np.array([0.3, 0.5, 1.0, 0.0], dtype=np.float64)
    light = (light / np.linalg.norm(light)).astype(np.float32)
    blob = mvp.tobytes() + light.tobytes()
    idx = tris.reshape(-1)

    def setup(ctx):
        color, depth, fbo = _make_targets(ctx, width, height, color_type, depth_type)
        vbo = ctx.create_vbo(verts)
        ibo = ctx.create_ibo(idx)
        vao = ctx.create_vao(vbo, ibo, [(0, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT),
                                        (16, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT)])
        ubo = ctx.create_ubo(blob)
        sh = ctx.create_shader("lambert", "lambert", ubo, vs_varyings=1, fs_varyings=1, cull=cull, blend=blend,
                               depth_test=depth_test, depth_mask=A.SL_DEPTH_MASK_ON)
        assert sh != A.INVALID_ID
        return dict(color=color, depth=depth, fbo=fbo, shader=sh, mesh=ctx.mesh(vao, 0, idx.size))

    def frame(ctx, st):
        ctx.clear_framebuffer(st["fbo"], 0, (0.0, 0.0, 0.0, 1.0), clear_depth)
        for _ in range(reps):
            ctx.draw(st["mesh"], st["shader"], st["fbo"])

    return Scene(name or f"sphere{tris.shape[0] // 1000}k_{width}x{height}", width, height, color_type, depth_type,
                 (0, 0, 0, 1), clear_depth, setup, frame, tris_per_frame=tris.shape[0] * reps, draws_per_frame=reps)


# ---- C2: instanced textured spheres (sl_instancing_test.cpp:80-171, 286-310) -
def instancing_scene(width=1920, height=1080, grid=(25, 20, 20), stacks=16, slices=32, depth_type=A.SL_COLOR_R_HALF,
                     name=None) -> Scene:
    pos, nrm, uv, tris = uv_sphere(stacks, slices)
    nverts = pos.shape[0]
    verts = np.zeros((nverts, 8), dtype=np.float32)
    verts[:, 0:3] = pos
    verts[:, 3:5] = uv * 4.0  # a few repeats so REPEAT wrapping is exercised
    verts[:, 5:8] = nrm
    idx = tris.reshape(-1).astype(np.uint16)
    gx, gy, gz = grid
    ninst = gx * gy * gz
    spacing = 25.0
    scale = 0.125 * 40.0  # spheres of radius 5 on a 25-unit lattice
    mats = np.zeros((ninst, 4, 4), dtype=np.float32)
    n = 0
    for z in range(gz):
        for y in range(gy):
            for x in range(gx):
                t = (spacing * (x - gx / 2.0 + 0.5), spacing * (y - gy / 2.0 + 0.5), -spacing * (z + 1.5))
                mats[n] = translate_scale(t, scale)
                n += 1
    tex = value_noise_texture(2, 256, 256, 3)
    proj = infinite_perspective(60.0, width / height, 0.01)
    cam = (0.0, 0.0, 30.0)
    view = look_at(cam, (0, 0, -100.0), (0, 1, 0))
    vp = matmul_cm(proj, view)
    model = np.identity(4, dtype=np.float32)
    # InstTexUniforms layout (shaders/sl_shaders_scene.inl): ptr@0, ambient@16, diffuse@32, ptr@48, model@64, vp@128, cam@192
    blob = bytearray(208)
    blob[16:32] = np.array([0.2, 0.2, 0.2, 1.0], dtype=np.float32).tobytes()
    blob[32:48] = np.array([0.8, 0.8, 0.8, 0.0], dtype=np.float32).tobytes()
    blob[64:128] = model.tobytes()
    blob[128:192] = vp.tobytes()
    blob[192:208] = np.array([cam[0], cam[1], cam[2], 1.0], dtype=np.float32).tobytes()

    def setup(ctx):
        color, depth, fbo = _make_targets(ctx, width, height, A.SL_COLOR_RGBA_8U, depth_type)
        vbo = ctx.create_vbo(verts)
        ibo = ctx.create_ibo(idx)
        vao = ctx.create_vao(vbo, ibo, [(0, 32, A.VERTEX_DIMENSION_3, A.VERTEX_DATA_FLOAT),
                                        (12, 32, A.VERTEX_DIMENSION_2, A.VERTEX_DATA_FLOAT),
                                        (20, 32, A.VERTEX_DIMENSION_3, A.VERTEX_DATA_FLOAT)])
        t = ctx.create_texture(A.SL_COLOR_RGB_8U, 256, 256, 1, tex)
        inst = ctx.create_vbo(mats)
        ubo = ctx.create_ubo(bytes(blob))
        ctx.ubo_bind_texture(ubo, 0, t)
        ctx.ubo_bind_buffer(ubo, 48, inst, 0)
        sh = ctx.create_shader("inst_tex", "inst_tex", ubo, vs_varyings=3, fs_varyings=3, cull=A.SL_CULL_BACK_FACE,
                               blend=A.SL_BLEND_PREMULTIPLED_ALPHA, depth_test=A.SL_DEPTH_TEST_GREATER_EQUAL,
                               depth_mask=A.SL_DEPTH_MASK_ON)
        assert sh != A.INVALID_ID
        return dict(color=color, depth=depth, fbo=fbo, shader=sh, mesh=ctx.mesh(vao, 0, idx.size), ubo=ubo, inst=inst,
                    per_frame={"inst": mats})

    def frame(ctx, st):
        ctx.clear_framebuffer(st["fbo"], 0, (0.6, 0.6, 0.6, 1.0), 0.0)
        ctx.draw_instanced(st["mesh"], ninst, st["shader"], st["fbo"])

    return Scene(name or f"instancing{ninst}_{width}x{height}", width, height, A.SL_COLOR_RGBA_8U, depth_type,
                 (0.6, 0.6, 0.6, 1.0), 0.0, setup, frame, tris_per_frame=tris.shape[0] * ninst, draws_per_frame=1)


# ---- C3: indoor scene, 16x16 inward-facing tessellated rooms, one textured PBR draw per room --------------------
def _room_mesh(quads: int, half: float):
    """Inward-facing box: 6 faces x quads^2 x 2 triangles, CCW seen from inside; vertex = (vec3 pos, vec2 uv, packed normal)."""
    verts, tris = [], []
    axes = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)), ((0, 0, 1), (1, 0, 0), (0, 1, 0))]
    g = np.linspace(-1.0, 1.0, quads + 1)
    for n_ax, u_ax, v_ax in axes:
        for sign in (1.0, -1.0):
            n = np.array(n_ax, dtype=np.float64) * sign          # inward normal of the wall at -sign*half
            u = np.array(u_ax, dtype=np.float64)
            v = np.array(v_ax, dtype=np.float64) * sign          # keeps u x v = n
            base = len(verts) and sum(len(x) for x in verts)
            aa, bb = np.meshgrid(g, g, indexing="ij")
            p = (-n[None, None, :] + aa[..., None] * u + bb[..., None] * v) * half
            uv = np.stack([(aa + 1.0) * 1.5, (bb + 1.0) * 1.5], axis=-1)   # three repeats per wall
            packed = ((int(n[0] * 511) & 0x3FF) | ((int(n[1] * 511) & 0x3FF) << 10) | ((int(n[2] * 511) & 0x3FF) << 20))
            rec = np.zeros((quads + 1, quads + 1), dtype=[("pos", "<f4", 3), ("uv", "<f4", 2), ("nrm", "<i4")])
            rec["pos"], rec["uv"], rec["nrm"] = p, uv, np.int32(packed)
            verts.append(rec.reshape(-1))
            i, j = np.meshgrid(np.arange(quads), np.arange(quads), indexing="ij")
            a = (base + i * (quads + 1) + j).reshape(-1)
            b = a + (quads + 1)
            c = a + 1
            d = b + 1
            tris.append(np.stack([a, b, c, b, d, c], axis=-1).reshape(-1, 3))
    return np.concatenate(verts), np.concatenate(tris).astype(np.uint32)


def _room_texture(i: int, size: int) -> np.ndarray:
    """Cheap, distinct, deterministic RGBA8 texture per room: one value-noise base, rolled and tinted by the room index."""
    base = _room_texture.cache.get(size)
    if base is None:
        base = value_noise_texture(3, size, size, 4)
        _room_texture.cache[size] = base
    t = np.roll(base, (37 * i % size, 91 * i % size), axis=(0, 1)).astype(np.int32)
    tint = np.array([(53 * i) % 97, (29 * i) % 89, (71 * i) % 83, 0], dtype=np.int32)
    t = np.clip(t + tint - 40, 0, 255).astype(np.uint8)
    t[..., 3] = 255
    return t


_room_texture.cache = {}


def indoor_scene(width=3840, height=2160, rooms=(16, 16), quads=26, tex_size=512, depth_type=A.SL_COLOR_R_HALF, name=None) -> Scene:
    half = 10.0
    verts, tris = _room_mesh(quads, half)
    idx = tris.reshape(-1)
    rx, rz = rooms
    nrooms = rx * rz
    spacing = 2.2 * half
    proj = infinite_perspective(60.0, width / height, 0.01)
    # just outside the grid, looking down the rows: the near wall of every room is back-facing (culled), so each room shows
    # its floor, ceiling, side walls and far wall, and the rooms of a row hide one another (early-Z rejects, overdraw)
    cam = np.array([-(rx / 2.0) * spacing - 1.2 * half, 0.25 * half, 0.15 * half])
    target = cam + np.array([1.0, -0.03, 0.22])
    view = look_at(cam, target, (0, 1, 0))
    vp = matmul_cm(proj, view)
    centers = [np.array([(x - rx / 2.0 + 0.5) * spacing, 0.0, (z - rz / 2.0) * spacing]) for z in range(rz) for x in range(rx)]
    light_pos = np.array([cam[0] + 1.5 * half, cam[1] + 0.4 * half, cam[2] + 0.2 * half, 1.0], dtype=np.float32)

    def blob_for(i):
        model = translate_scale(centers[i].astype(np.float32), 1.0)
        mvp = matmul_cm(vp, model)
        # MeshUniforms: ptr@0, camPos@16, light{pos@32, ambient@48, diffuse@64}, point@80, model@96, mvp@160
        b = bytearray(224)
        b[16:32] = np.array([cam[0], cam[1], cam[2], 1.0], dtype=np.float32).tobytes()
        b[32:48] = light_pos.tobytes()
        b[48:64] = np.array([0.25, 0.25, 0.25, 1.0], dtype=np.float32).tobytes()
        b[64:80] = np.array([0.09, 0.085, 0.08, 1.0], dtype=np.float32).tobytes()
        b[80:96] = np.array([1.0, 0.09, 0.032, 0.0], dtype=np.float32).tobytes()
        b[96:160] = model.tobytes()
        b[160:224] = mvp.tobytes()
        return bytes(b)

    blobs = [blob_for(i) for i in range(nrooms)]

    def setup(ctx):
        color, depth, fbo = _make_targets(ctx, width, height, A.SL_COLOR_RGBA_8U, depth_type)
        vbo = ctx.create_vbo(verts)
        ibo = ctx.create_ibo(idx)
        vao = ctx.create_vao(vbo, ibo, [(0, 24, A.VERTEX_DIMENSION_3, A.VERTEX_DATA_FLOAT), (12, 24, A.VERTEX_DIMENSION_2, A.VERTEX_DATA_FLOAT),
                                        (20, 24, A.VERTEX_DIMENSION_1, A.VERTEX_DATA_INT)])
        texs = [ctx.create_texture(A.SL_COLOR_RGBA_8U, tex_size, tex_size, 1, _room_texture(i, tex_size), order=A.TEXEL_SWIZZLED)
                for i in range(nrooms)]
        ubo = ctx.create_ubo(blobs[0])
        ctx.ubo_bind_texture(ubo, 0, texs[0])
        sh = ctx.create_shader("mesh_pbr", "mesh_pbr", ubo, vs_varyings=3, fs_varyings=3, cull=A.SL_CULL_BACK_FACE, blend=A.SL_BLEND_OFF,
                               depth_test=A.SL_DEPTH_TEST_GREATER_EQUAL, depth_mask=A.SL_DEPTH_MASK_ON)
        assert sh != A.INVALID_ID
        return dict(color=color, depth=depth, fbo=fbo, shader=sh, mesh=ctx.mesh(vao, 0, idx.size), ubo=ubo, texs=texs)

    def frame(ctx, st):
        ctx.clear_framebuffer(st["fbo"], 0, (0.0, 0.0, 0.0, 1.0), 0.0)
        if "calls" not in st:
            # the demo's pattern: per-mesh uniforms, then draw (sl_large_scene_test.cpp:700-730); recorded once, replayed from C
            ubo, texs, mesh, sh, fbo = st["ubo"], st["texs"], st["mesh"], st["shader"], st["fbo"]
            calls = A.CommandList(ctx)
            for i in range(nrooms):
                calls.ubo_assign(ubo, blobs[i])
                calls.ubo_bind_texture(ubo, 0, texs[i])
                calls.draw(mesh, sh, fbo)
            st["calls"] = calls
        st["calls"].replay()

    return Scene(name or f"indoor{nrooms}x{tris.shape[0]}_{width}x{height}", width, height, A.SL_COLOR_RGBA_8U, depth_type,
                 (0, 0, 0, 1), 0.0, setup, frame, tris_per_frame=tris.shape[0] * nrooms, draws_per_frame=nrooms,
                 upload_bytes_per_frame=1024 * nrooms)


# ---- torture scene: random clip-space triangles (overlap, near-plane clipping, tiny + huge) ----
def random_tris_scene(seed=1, n=400, width=200, height=150, blend=A.SL_BLEND_OFF, depth_test=A.SL_DEPTH_TEST_LESS_THAN,
                      depth_mask=A.SL_DEPTH_MASK_ON, depth_type=A.SL_COLOR_R_FLOAT, color_type=A.SL_COLOR_RGBA_8U,
                      cull=A.SL_CULL_OFF, index_dtype=np.uint16, clip_fraction=0.3, clear_depth=1.0, viewport=None,
                      scissor=None, split_draws=1, multiple=False, name=None, empty_draw=False, mrt_types=None, points=False, instances=1, wire=False,
                      lines=False) -> Scene:
    """points: the same vertices drawn as RENDER_MODE_POINTS / INDEXED_POINTS (one fragment per visible vertex,
    SL_PointProcessor / SL_PointRasterizer); instances > 1 issues draw_instanced (the vertex shader ignores the id, so every
    instance lands on the same pixels: a pure ordering test).
    lines: the same vertices drawn as RENDER_MODE_LINES / INDEXED_LINES (SL_LineProcessor / SL_LineRasterizer): consecutive
    element pairs are segments.
    mrt_types: colour types of render targets 1..3 (sl_mrt_test.cpp:150-200 attaches four); the "vcolor_mrt" fragment shader
    then writes four outputs and SceneInstance.read_colors() returns all attachments."""
    rng = np.random.default_rng(seed)
    verts = np.zeros((n * 3, 8), dtype=np.float32)
    centers = rng.uniform(-1.1, 1.1, size=(n, 1, 2))
    size = np.where(rng.random((n, 1, 1)) < 0.15, 1.2, 0.12) * rng.uniform(0.2, 1.0, size=(n, 1, 1))
    xy = centers + rng.uniform(-1, 1, size=(n, 3, 2)) * size
    w = rng.uniform(0.5, 2.0, size=(n, 3))
    clipped = rng.random(n) < clip_fraction
    w[clipped, 0] = rng.uniform(-0.5, 0.2, size=int(clipped.sum()))  # one vertex behind / on the near side
    z = rng.uniform(-0.9, 0.9, size=(n, 3)) * np.abs(w)
    verts[:, 0:2] = (xy * w[..., None]).reshape(-1, 2)
    verts[:, 2] = z.reshape(-1)
    verts[:, 3] = w.reshape(-1)
    verts[:, 4:7] = rng.uniform(0, 1, size=(n * 3, 3))
    verts[:, 7] = rng.uniform(0.2, 1.0, size=n * 3) if blend != A.SL_BLEND_OFF else 1.0
    if points:
        # vertices exactly on the clip planes: ndc 1.0 maps to x == width (the texel wraps into the next row) or y == height
        # (the reference's allocation padding), ndc -1.0 to column / row 0
        edge = rng.random(n * 3) < 0.15
        side = rng.integers(0, 4, size=n * 3)
        verts[edge & (side == 0), 0] = verts[edge & (side == 0), 3]
        verts[edge & (side == 1), 0] = -verts[edge & (side == 1), 3]
        verts[edge & (side == 2), 1] = verts[edge & (side == 2), 3]
        verts[edge & (side == 3), 1] = -verts[edge & (side == 3), 3]
    idx = None
    if index_dtype is not None:
        idx = rng.permutation(n * 3).astype(np.int64)
        # keep triangles intact but reference vertices through a shuffled index buffer
        order = np.argsort(idx)
        verts = verts[idx]          # vertex i now lives at position order[i]
        idx = order.astype(index_dtype)
        if index_dtype == np.uint8:
            assert n * 3 <= 256
    count = n * 3

    def setup(ctx):
        color, depth, fbo = _make_targets(ctx, width, height, color_type, depth_type)
        extra = []
        if mrt_types:
            extra = [ctx.create_texture(t, width, height, 1) for t in mrt_types]
            fbo = ctx.create_framebuffer([color] + extra, depth)
            assert ctx.fbo_valid(fbo) == 0
        vbo = ctx.create_vbo(verts)
        ibo = ctx.create_ibo(idx) if idx is not None else None
        vao = ctx.create_vao(vbo, ibo, [(0, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT),
                                        (16, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT)])
        ubo = ctx.create_ubo(b"\0" * 16)
        sh = ctx.create_shader("vcolor", "vcolor_mrt" if mrt_types else "vcolor", ubo, vs_varyings=1, fs_varyings=1, cull=cull, blend=blend,
                               depth_test=depth_test, depth_mask=depth_mask, outputs=1 + len(extra))
        assert sh != A.INVALID_ID
        mode = A.RENDER_MODE_INDEXED_TRIANGLES if idx is not None else A.RENDER_MODE_TRIANGLES
        if points:
            mode = A.RENDER_MODE_INDEXED_POINTS if idx is not None else A.RENDER_MODE_POINTS
        if wire:  # SL_TriRasterizer::render_wireframe: the same triangles, edges only
            mode = A.RENDER_MODE_INDEXED_TRI_WIRE if idx is not None else A.RENDER_MODE_TRI_WIRE
        if lines:
            mode = A.RENDER_MODE_INDEXED_LINES if idx is not None else A.RENDER_MODE_LINES
        per = (n // split_draws) * 3
        meshes = []
        for d in range(split_draws):
            b = d * per
            e = count if d == split_draws - 1 else b + per
            if empty_draw:
                e = b
            meshes.append(ctx.mesh(vao, b, e, mode))
        if viewport is not None:
            ctx.viewport(*viewport)
        if scissor is not None:
            ctx.scissor(*scissor)
        return dict(color=color, depth=depth, fbo=fbo, shader=sh, meshes=meshes, extra=extra, extra_types=list(mrt_types or []))

    def frame(ctx, st):
        ctx.clear_framebuffer(st["fbo"], 0, (0.1, 0.2, 0.3, 1.0), clear_depth)
        for i in range(len(st["extra"])):
            ctx.clear_color_buffer(st["fbo"], 1 + i, (0.05 * (i + 1), 0.5, 0.25, 1.0))
        if multiple:
            ctx.draw_multiple(st["meshes"], st["shader"], st["fbo"])
        elif instances > 1:
            for m in st["meshes"]:
                ctx.draw_instanced(m, instances, st["shader"], st["fbo"])
        else:
            for m in st["meshes"]:
                ctx.draw(m, st["shader"], st["fbo"])

    return Scene(name or f"random{n}_s{seed}", width, height, color_type, depth_type, (0.1, 0.2, 0.3, 1.0), clear_depth,
                 setup, frame, tris_per_frame=n, draws_per_frame=split_draws)


# ---- C4: skinned characters (sl_animation_test.cpp:199-280): 64-bone capsule chains, alpha blending -------------------
def _pack_normal(n: np.ndarray) -> np.ndarray:
    """10-10-10-2 signed packing read back by sl_unpack_vec4_10_10_10_2I (SL_PackedVertex.hpp:20-100)."""
    q = np.clip(np.rint(n * 511.0), -511, 511).astype(np.int32)
    return ((q[..., 0] & 0x3FF) | ((q[..., 1] & 0x3FF) << 10) | ((q[..., 2] & 0x3FF) << 20)).astype(np.int32)


def _capsule_chain(bones: int, around: int, radius: float, height: float):
    """A tube of `bones` segments along +Y, `around` quads per ring; every ring follows its own bone and its neighbours.
    Vertex = Tuple<vec3 pos, vec2h uv, int32 packed normal, vec4<u16> bone ids, vec4h bone weights> (36 bytes)."""
    rings = bones + 1
    k = np.arange(rings)
    a = np.arange(around + 1) * (2.0 * np.pi / around)
    taper = np.sin(np.clip((k + 0.5) / rings, 0, 1) * np.pi) ** 0.35          # rounded ends
    rr = (radius * taper)[:, None]
    pos = np.stack([rr * np.cos(a)[None, :], np.broadcast_to((k * (height / bones))[:, None], (rings, around + 1)),
                    rr * np.sin(a)[None, :]], axis=-1)
    nrm = np.stack([np.broadcast_to(np.cos(a)[None, :], (rings, around + 1)), np.zeros((rings, around + 1)),
                    np.broadcast_to(np.sin(a)[None, :], (rings, around + 1))], axis=-1)
    uv = np.stack([np.broadcast_to((np.arange(around + 1) / around * 4.0)[None, :], (rings, around + 1)),
                   np.broadcast_to((k / bones * 6.0)[:, None], (rings, around + 1))], axis=-1)
    ids = np.stack([np.clip(k, 0, bones - 1), np.clip(k - 1, 0, bones - 1), np.clip(k + 1, 0, bones - 1), np.clip(k - 2, 0, bones - 1)], axis=-1)
    wts = np.array([0.5, 0.25, 0.1875, 0.0625])                               # exact in half precision, sum 1
    rec = np.zeros((rings, around + 1), dtype=[("pos", "<f4", 3), ("uv", "<f2", 2), ("nrm", "<i4"), ("ids", "<u2", 4), ("wts", "<f2", 4)])
    rec["pos"], rec["uv"], rec["nrm"] = pos, uv, _pack_normal(nrm)
    rec["ids"] = ids[:, None, :]
    rec["wts"] = wts[None, None, :]
    assert rec.dtype.itemsize == 36
    i, j = np.meshgrid(np.arange(bones), np.arange(around), indexing="ij")
    v00 = (i * (around + 1) + j).reshape(-1)
    v01 = v00 + 1
    v10 = v00 + (around + 1)
    v11 = v10 + 1
    tris = np.stack([v00, v10, v01, v01, v10, v11], axis=-1).reshape(-1, 3)   # CCW seen from outside
    return rec.reshape(-1), tris.astype(np.uint32)


def _rot_z(angle: float) -> np.ndarray:
    c, s_ = np.cos(angle), np.sin(angle)
    m = np.identity(4)
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s_, s_, c
    return m


def _bone_chain(bones: int, seg: float, phase: float, amp: float) -> np.ndarray:
    """Column-major mat4 per bone: a chain that sways; bone b rotates everything above joint b a little further."""
    out = np.zeros((bones, 4, 4), dtype=np.float32)
    cur = np.identity(4)
    for b in range(bones):
        ang = amp * np.sin(phase + 0.21 * b)
        joint = np.identity(4); joint[1, 3] = b * seg
        back = np.identity(4); back[1, 3] = -b * seg
        cur = cur @ joint @ _rot_z(ang) @ back
        out[b] = cur.T.astype(np.float32)   # stored column-major
    return out


def skinned_scene(width=1920, height=1080, grid=(10, 10), bones=64, around=160, tex_size=256, depth_type=A.SL_COLOR_R_HALF, name=None) -> Scene:
    char_h, radius = 7.0, 0.9
    verts, tris = _capsule_chain(bones, around, radius, char_h)
    idx = tris.reshape(-1)
    gx, gz = grid
    nchars = gx * gz
    spacing = 2.6
    all_bones = np.concatenate([_bone_chain(bones, char_h / bones, 0.37 * i, 0.035 + 0.004 * (i % 5)) for i in range(nchars)])
    tex = value_noise_texture(4, tex_size, tex_size, 4)
    tex[..., 3] = 96 + (tex[..., 0] // 2)            # translucent: alpha 96..223 -> SL_BLEND_ALPHA does real work
    proj = infinite_perspective(60.0, width / height, 0.01)
    cam = (0.0, 6.0, 0.5 * gz * spacing + 9.0)
    view = look_at(cam, (0.0, 3.0, 0.0), (0, 1, 0))
    vp = matmul_cm(proj, view)

    def blob_for(i):
        x, z = i % gx, i // gx
        model = translate_scale(((x - gx / 2.0 + 0.5) * spacing, 0.0, (z - gz / 2.0 + 0.5) * spacing), 1.0)
        # AnimUniforms: pTexture@0, pBones@8, model@16, vp@80, camPos@144
        b = bytearray(160)
        b[16:80] = model.tobytes()
        b[80:144] = vp.tobytes()
        b[144:160] = np.array([cam[0], cam[1], cam[2], 1.0], dtype=np.float32).tobytes()
        return bytes(b)

    blobs = [blob_for(i) for i in range(nchars)]

    def setup(ctx):
        color, depth, fbo = _make_targets(ctx, width, height, A.SL_COLOR_RGBA_8U, depth_type)
        vbo = ctx.create_vbo(verts)
        ibo = ctx.create_ibo(idx)
        vao = ctx.create_vao(vbo, ibo, [(0, 36, A.VERTEX_DIMENSION_3, A.VERTEX_DATA_FLOAT), (12, 36, A.VERTEX_DIMENSION_2, A.VERTEX_DATA_SHORT),
                                        (16, 36, A.VERTEX_DIMENSION_1, A.VERTEX_DATA_INT), (20, 36, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_SHORT),
                                        (28, 36, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_SHORT)])
        t = ctx.create_texture(A.SL_COLOR_RGBA_8U, tex_size, tex_size, 1, tex)
        bone_vbo = ctx.create_vbo(all_bones)
        ubo = ctx.create_ubo(blobs[0])
        ctx.ubo_bind_texture(ubo, 0, t)
        ctx.ubo_bind_buffer(ubo, 8, bone_vbo, 0)
        sh = ctx.create_shader("skin_tex", "skin_tex", ubo, vs_varyings=3, fs_varyings=3, cull=A.SL_CULL_BACK_FACE, blend=A.SL_BLEND_ALPHA,
                               depth_test=A.SL_DEPTH_TEST_GREATER_EQUAL, depth_mask=A.SL_DEPTH_MASK_ON)
        assert sh != A.INVALID_ID
        return dict(color=color, depth=depth, fbo=fbo, shader=sh, mesh=ctx.mesh(vao, 0, idx.size), ubo=ubo, tex=t, bones=bone_vbo,
                    per_frame={"bones": all_bones})

    def frame(ctx, st):
        ctx.clear_framebuffer(st["fbo"], 0, (0.2, 0.2, 0.25, 1.0), 0.0)
        if "calls" not in st:
            # the demo's pattern: per-mesh uniforms (model matrix, bone palette), then draw; recorded once, replayed from C
            ubo, tex, bonebuf, mesh, sh, fbo = st["ubo"], st["tex"], st["bones"], st["mesh"], st["shader"], st["fbo"]
            calls = A.CommandList(ctx)
            for i in range(nchars):
                calls.ubo_assign(ubo, blobs[i])
                calls.ubo_bind_texture(ubo, 0, tex)
                calls.ubo_bind_buffer(ubo, 8, bonebuf, i * bones * 64)
                calls.draw(mesh, sh, fbo)
            st["calls"] = calls
        st["calls"].replay()

    return Scene(name or f"skinned{nchars}x{tris.shape[0]}_{width}x{height}", width, height, A.SL_COLOR_RGBA_8U, depth_type,
                 (0.2, 0.2, 0.25, 1.0), 0.0, setup, frame, tris_per_frame=tris.shape[0] * nchars, draws_per_frame=nchars,
                 upload_bytes_per_frame=all_bones.nbytes + 1024 * nchars)


# ---- C5: volume ray march (sl_volume_rendering_test.cpp:72-254, 298-520) ---------------------------------------------
def synthetic_volume(n: int, seed: int = 5) -> np.ndarray:
    """R8 n^3 volume: six Gaussian blobs + low-frequency value noise, empty (0) outside a sphere -- the ray marcher gets
    empty-space rejection, early saturation and long marches, like a CT head."""
    rng = np.random.default_rng(seed)
    g = (np.arange(n, dtype=np.float32) + 0.5) / n * 2.0 - 1.0
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    vol = np.zeros((n, n, n), dtype=np.float32)
    for _ in range(6):
        c = rng.uniform(-0.45, 0.45, size=3)
        sg = rng.uniform(0.18, 0.38)
        amp = rng.uniform(0.45, 1.0)
        vol += amp * np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (2.0 * sg * sg))
    m = max(n // 16, 2)
    coarse = rng.random((m + 1, m + 1, m + 1)).astype(np.float32)
    t = (np.arange(n) + 0.5) / n * m
    i0 = np.minimum(t.astype(int), m - 1)
    f = (t - i0).astype(np.float32)

    def lerp_axis(a, axis):
        lo = np.take(a, i0, axis=axis)
        hi = np.take(a, i0 + 1, axis=axis)
        shape = [1, 1, 1]
        shape[axis] = n
        w = f.reshape(shape)
        return lo * (1 - w) + hi * w

    noise = lerp_axis(lerp_axis(lerp_axis(coarse, 0), 1), 2)
    vol = vol * (0.75 + 0.5 * noise)
    vol[(x * x + y * y + z * z) > 0.92] = 0.0
    return np.clip(vol * 200.0, 0, 255).astype(np.uint8)


def _unit_cube() -> np.ndarray:
    """36 vec3 vertices, 12 triangles, CCW seen from outside, non-indexed."""
    out = []
    for axis in range(3):
        for sign in (1.0, -1.0):
            u, v = (axis + 1) % 3, (axis + 2) % 3
            def P(a, b):
                p = [0.0, 0.0, 0.0]
                p[axis], p[u], p[v] = sign, a, b
                return p
            quad = [P(-1, -1), P(1, -1), P(1, 1), P(-1, 1)]
            if sign < 0:
                quad = quad[::-1]
            out += [quad[0], quad[1], quad[2], quad[2], quad[3], quad[0]]
    return np.array(out, dtype=np.float32)


def volume_scene(width=7680, height=4320, vol_size=256, depth_type=A.SL_COLOR_R_HALF, name=None, blend=A.SL_BLEND_PREMULTIPLED_ALPHA, fs="volume") -> Scene:
    verts = _unit_cube()
    vol = synthetic_volume(vol_size)
    # transfer functions of the demo (sl_volume_rendering_test.cpp:405-520): 16-wide bands; the last entry of each band stays 0
    opacity = np.zeros(256, dtype=np.float32)
    for b, o in enumerate([0.0, 0.1, 0.1, 0.25, 0.5, 0.2] + [0.05] * 10):
        opacity[16 * b:16 * b + 15] = o
    colors = np.zeros((256, 3), dtype=np.float32)
    band_rgb = [(0, 0, 0), (0.2, 0.2, 0.5), (0.4, 0.3, 0.1), (1.0, 0.3, 0.3), (1.0, 1.0, 1.0), (0.6, 0.6, 0.6)] + [(1.0, 1.0, 1.0)] * 10
    for b, c in enumerate(band_rgb):
        colors[16 * b:16 * b + 15] = c
    proj = perspective(60.0, width / height, 0.1, 100.0)
    cam = np.array([1.35, 1.0, 1.9])
    view = look_at(cam, (0, 0, 0), (0, 1, 0))
    vp = matmul_cm(proj, view)
    # VolumeUniforms: pCubeMap@0, pOpacityMap@8, pColorMap@16, spacing@32, camPos@48, mvp@64
    blob = bytearray(128)
    blob[32:48] = np.array([1.0, 1.0, 1.0, 1.0], dtype=np.float32).tobytes()
    blob[48:64] = np.array([cam[0], cam[1], cam[2], 0.0], dtype=np.float32).tobytes()
    blob[64:128] = vp.tobytes()

    def setup(ctx):
        color, depth, fbo = _make_targets(ctx, width, height, A.SL_COLOR_RGBA_FLOAT, depth_type)
        vbo = ctx.create_vbo(verts)
        vao = ctx.create_vao(vbo, None, [(0, 12, A.VERTEX_DIMENSION_3, A.VERTEX_DATA_FLOAT)])
        tv = ctx.create_texture(A.SL_COLOR_R_8U, vol_size, vol_size, vol_size, vol)
        to = ctx.create_texture(A.SL_COLOR_R_FLOAT, 256, 1, 1, opacity)
        tc = ctx.create_texture(A.SL_COLOR_RGB_FLOAT, 256, 1, 1, colors)
        ubo = ctx.create_ubo(bytes(blob))
        ctx.ubo_bind_texture(ubo, 0, tv)
        ctx.ubo_bind_texture(ubo, 8, to)
        ctx.ubo_bind_texture(ubo, 16, tc)
        sh = ctx.create_shader("volume", fs, ubo, vs_varyings=2, fs_varyings=2, cull=A.SL_CULL_BACK_FACE,
                               blend=blend, depth_test=A.SL_DEPTH_TEST_OFF, depth_mask=A.SL_DEPTH_MASK_OFF)
        assert sh != A.INVALID_ID
        return dict(color=color, depth=depth, fbo=fbo, shader=sh, mesh=ctx.mesh(vao, 0, 36, A.RENDER_MODE_TRIANGLES), ubo=ubo)

    def frame(ctx, st):
        ctx.clear_framebuffer(st["fbo"], 0, (0.0, 0.0, 0.0, 0.0), 0.0)
        ctx.draw(st["mesh"], st["shader"], st["fbo"])

    return Scene(name or f"volume{vol_size}_{width}x{height}", width, height, A.SL_COLOR_RGBA_FLOAT, depth_type, (0, 0, 0, 0), 0.0,
                 setup, frame, tris_per_frame=12, draws_per_frame=1)


# ---- sampler matrix: every filter x wrap mode x texel order the shader API offers, one viewport tile per combination ----
def sampler_matrix_scene(tile_w=64, tile_h=48, name=None) -> Scene:
    cols, rows, modes = 4, 3, 12
    width, height = cols * tile_w, rows * tile_h
    rng = np.random.default_rng(4242)
    # a quad over the whole viewport; the "colour" attribute carries (u, v, w): beyond [0, 1) on every side
    def vert(x, y):
        u = -0.6 + 2.3 * (x + 1) / 2
        v = -0.4 + 1.9 * (y + 1) / 2
        return [x, y, 0.0, 1.0, u, v, 0.35 * u + 0.4 * v + 0.1, 1.0]
    quad = np.array([vert(-1, -1), vert(1, -1), vert(1, 1), vert(-1, -1), vert(1, 1), vert(-1, 1)], dtype=np.float32)
    rgba2d = rng.integers(0, 256, size=(12, 20, 4), dtype=np.uint8)
    rgb2d = rng.integers(0, 256, size=(13, 19, 3), dtype=np.uint8)
    r3d = rng.integers(0, 256, size=(8, 8, 12, 1), dtype=np.uint8)
    rgba3d = rng.integers(0, 256, size=(4, 8, 8, 4), dtype=np.uint8)

    def blob(mode):
        b = bytearray(64)
        b[48:52] = np.array([mode], dtype=np.int32).tobytes()
        return bytes(b)

    def setup(ctx):
        color, depth, fbo = _make_targets(ctx, width, height, A.SL_COLOR_RGBA_FLOAT, A.SL_COLOR_R_FLOAT)
        vbo = ctx.create_vbo(quad)
        vao = ctx.create_vao(vbo, None, [(0, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT), (16, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT)])
        texs = [ctx.create_texture(A.SL_COLOR_RGBA_8U, 20, 12, 1, rgba2d, A.TEXEL_ORDERED),
                ctx.create_texture(A.SL_COLOR_RGBA_8U, 20, 12, 1, rgba2d, A.TEXEL_SWIZZLED),
                ctx.create_texture(A.SL_COLOR_RGB_8U, 19, 13, 1, rgb2d, A.TEXEL_ORDERED),
                ctx.create_texture(A.SL_COLOR_R_8U, 12, 8, 8, r3d, A.TEXEL_ORDERED),
                ctx.create_texture(A.SL_COLOR_R_8U, 12, 8, 8, r3d, A.TEXEL_SWIZZLED),
                ctx.create_texture(A.SL_COLOR_RGBA_8U, 8, 8, 4, rgba3d, A.TEXEL_ORDERED)]
        ubo = ctx.create_ubo(blob(0))
        for i, t in enumerate(texs):
            ctx.ubo_bind_texture(ubo, 8 * i, t)
        sh = ctx.create_shader("vcolor", "sampler_matrix", ubo, vs_varyings=1, fs_varyings=1, cull=A.SL_CULL_OFF, blend=A.SL_BLEND_OFF,
                               depth_test=A.SL_DEPTH_TEST_OFF, depth_mask=A.SL_DEPTH_MASK_OFF)
        assert sh != A.INVALID_ID
        return dict(color=color, depth=depth, fbo=fbo, shader=sh, ubo=ubo, mesh=ctx.mesh(vao, 0, 6, A.RENDER_MODE_TRIANGLES))

    def frame(ctx, st):
        ctx.viewport(0, 0, width, height)
        ctx.clear_framebuffer(st["fbo"], 0, (0.0, 0.0, 0.0, 0.0), 0.0)
        for mode in range(modes):
            ctx.viewport((mode % cols) * tile_w, (mode // cols) * tile_h, tile_w, tile_h)
            ctx.ubo_assign(st["ubo"], blob(mode))
            ctx.draw(st["mesh"], st["shader"], st["fbo"])
        ctx.viewport(0, 0, 65535, 65535)

    return Scene(name or f"sampler_matrix_{width}x{height}", width, height, A.SL_COLOR_RGBA_FLOAT, A.SL_COLOR_R_FLOAT, (0, 0, 0, 0), 0.0,
                 setup, frame, tris_per_frame=2 * modes, draws_per_frame=modes)


SCENES: Dict[str, Callable[[], Scene]] = {
    "sampler_matrix": sampler_matrix_scene,
    "draw_test": draw_test_scene,
    "c1_sphere100k": sphere_scene,
    "c1_small": lambda: sphere_scene(256, 192, 48, 48, name="sphere4k_256x192"),
    "c2_instancing10k": instancing_scene,
    "c2_small": lambda: instancing_scene(320, 240, (5, 4, 3), 8, 16, name="instancing60_320x240"),
    # the survey's 4K anchors: the C1 sphere at 3840x2160 with 100k and 1M triangles
    "c3_indoor2m": indoor_scene,
    "c3_small": lambda: indoor_scene(480, 270, (4, 4), 6, 64, name="indoor16x432_480x270"),
    "c4_skinned100": skinned_scene,
    "c4_small": lambda: skinned_scene(480, 270, (3, 2), 64, 24, 64, name="skinned6x3072_480x270"),
    "c5_volume8k": volume_scene,
    "c5_volume4k": lambda: volume_scene(3840, 2160, 256, name="volume256_3840x2160"),
    "c5_volume4k_noblend": lambda: volume_scene(3840, 2160, 256, name="volume256_3840x2160_noblend", blend=A.SL_BLEND_OFF),
    "c5_dbg0": lambda: volume_scene(3840, 2160, 256, name="volume_dbg0", blend=A.SL_BLEND_OFF, fs="volume_dbg0"),
    "c5_dbg1": lambda: volume_scene(3840, 2160, 256, name="volume_dbg1", blend=A.SL_BLEND_OFF, fs="volume_dbg1"),
    "c5_dbg2": lambda: volume_scene(3840, 2160, 256, name="volume_dbg2", blend=A.SL_BLEND_OFF, fs="volume_dbg2"),
    "c5_small": lambda: volume_scene(160, 90, 32, name="volume32_160x90"),
    "c1_sphere100k_4k": lambda: sphere_scene(3840, 2160, 224, 224, name="sphere100k_3840x2160"),
    "c1_sphere1m_4k": lambda: sphere_scene(3840, 2160, 708, 708, name="sphere1m_3840x2160"),
}
