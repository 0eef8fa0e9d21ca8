"""ctypes binding of the C ABI declared in include/slcu.h.

The same binder drives three libraries that export the same entry points under
different prefixes:

* ``slcu_``  -- the product: softlight_b200/libslcu.so (hand-written sm_100a CUDA).
* ``slref_`` -- TEST ONLY: oracle/_ref/libslref_{strict,fast}.so, the unmodified
  reference compiled from /root/reference (oracle/Makefile).
* ``slor_``  -- TEST ONLY: oracle/libslor.so, the CPU restatement.

Only tests/, bench.py's cpu_baseline/reference arm and __graft_entry__.smoke()
load the two checker libraries.  The product path never does, and opening the
product library fails loudly if it has not been built.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LIB_PATHS = {
    "cuda": os.path.join(ROOT, "softlight_b200", "libslcu.so"),
    "ref_strict": os.path.join(ROOT, "oracle", "_ref", "libslref_strict.so"),
    "ref_fast": os.path.join(ROOT, "oracle", "_ref", "libslref_fast.so"),
    "port": os.path.join(ROOT, "oracle", "libslor.so"),
}
if os.environ.get("SLCU_LIB"):  # development aid: A/B a differently compiled product library
    LIB_PATHS["cuda"] = os.path.abspath(os.environ["SLCU_LIB"])
PREFIX = {"cuda": "slcu_", "ref_strict": "slref_", "ref_fast": "slref_", "port": "slor_"}

INVALID_ID = C.c_size_t(-1).value

# ---- reference enum values (passed through unchanged) -----------------------
# SL_ColorDataType (softlight/include/softlight/SL_Color.hpp:31-73)
(SL_COLOR_R_8U, SL_COLOR_R_16U, SL_COLOR_R_32U, SL_COLOR_R_64U, SL_COLOR_R_HALF, SL_COLOR_R_FLOAT, SL_COLOR_R_DOUBLE,
 SL_COLOR_RG_8U, SL_COLOR_RG_16U, SL_COLOR_RG_32U, SL_COLOR_RG_64U, SL_COLOR_RG_HALF, SL_COLOR_RG_FLOAT, SL_COLOR_RG_DOUBLE,
 SL_COLOR_RGB_8U, SL_COLOR_RGB_16U, SL_COLOR_RGB_32U, SL_COLOR_RGB_64U, SL_COLOR_RGB_HALF, SL_COLOR_RGB_FLOAT, SL_COLOR_RGB_DOUBLE,
 SL_COLOR_RGBA_8U, SL_COLOR_RGBA_16U, SL_COLOR_RGBA_32U, SL_COLOR_RGBA_64U, SL_COLOR_RGBA_HALF, SL_COLOR_RGBA_FLOAT,
 SL_COLOR_RGBA_DOUBLE,
 # the packed ("compressed") types of SL_ColorCompressed.hpp: one unsigned integer per texel, r in the lowest bits
 SL_COLOR_RGB_332, SL_COLOR_RGB_565, SL_COLOR_RGBA_5551, SL_COLOR_RGBA_4444, SL_COLOR_RGBA_1010102) = range(33)

_CHAN_BYTES = [1, 2, 4, 8, 2, 4, 8]
_PACKED_BYTES = {SL_COLOR_RGB_332: 1, SL_COLOR_RGB_565: 2, SL_COLOR_RGBA_5551: 2, SL_COLOR_RGBA_4444: 2, SL_COLOR_RGBA_1010102: 4}


def is_packed_color(color_type: int) -> bool:
    return color_type in _PACKED_BYTES


def texel_channels(color_type: int) -> int:
    """Array elements per texel: the component count, 1 for a packed type (the whole texel is one integer)."""
    return 1 if color_type in _PACKED_BYTES else color_type // 7 + 1


def bytes_per_texel(color_type: int) -> int:
    if color_type in _PACKED_BYTES:
        return _PACKED_BYTES[color_type]
    return _CHAN_BYTES[color_type % 7] * (color_type // 7 + 1)


def texel_dtype(color_type: int):
    if color_type in _PACKED_BYTES:
        return {1: np.uint8, 2: np.uint16, 4: np.uint32}[_PACKED_BYTES[color_type]]
    return [np.uint8, np.uint16, np.uint32, np.uint64, np.float16, np.float32, np.float64][color_type % 7]


# SL_DataType / SL_Dimension (SL_Geometry.hpp:17-39)
VERTEX_DATA_BYTE, VERTEX_DATA_SHORT, VERTEX_DATA_INT, VERTEX_DATA_LONG, VERTEX_DATA_FLOAT, VERTEX_DATA_DOUBLE = range(6)
VERTEX_DIMENSION_1, VERTEX_DIMENSION_2, VERTEX_DIMENSION_3, VERTEX_DIMENSION_4 = range(4)
# SL_RenderMode (SL_Mesh.hpp)
RENDER_MODE_POINTS = 0x01
RENDER_MODE_INDEXED_POINTS = 0x03
RENDER_MODE_LINES = 0x04
RENDER_MODE_INDEXED_LINES = 0x0C
RENDER_MODE_TRIANGLES = 0x10
RENDER_MODE_INDEXED_TRIANGLES = 0x30
RENDER_MODE_TRI_WIRE = 0x50
RENDER_MODE_INDEXED_TRI_WIRE = 0x70
# SL_PipelineState.hpp:15-64
SL_CULL_BACK_FACE, SL_CULL_FRONT_FACE, SL_CULL_OFF = range(3)
(SL_DEPTH_TEST_OFF, SL_DEPTH_TEST_LESS_THAN, SL_DEPTH_TEST_LESS_EQUAL, SL_DEPTH_TEST_GREATER_THAN,
 SL_DEPTH_TEST_GREATER_EQUAL, SL_DEPTH_TEST_EQUAL, SL_DEPTH_TEST_NOT_EQUAL) = range(7)
SL_DEPTH_MASK_OFF, SL_DEPTH_MASK_ON = range(2)
SL_BLEND_OFF, SL_BLEND_ALPHA, SL_BLEND_PREMULTIPLED_ALPHA, SL_BLEND_ADDITIVE, SL_BLEND_SCREEN = range(5)
TEXEL_ORDERED, TEXEL_SWIZZLED = 0, 1


class Mesh(C.Structure):
    _fields_ = [("vaoId", C.c_uint64), ("elementBegin", C.c_uint64), ("elementEnd", C.c_uint64),
                ("mode", C.c_uint32), ("materialId", C.c_uint32)]


class VertexShader(C.Structure):
    _fields_ = [("numVaryings", C.c_uint8), ("cullMode", C.c_uint8), ("shader", C.c_char_p)]


class FragmentShader(C.Structure):
    _fields_ = [("numVaryings", C.c_uint8), ("numOutputs", C.c_uint8), ("blend", C.c_uint8),
                ("depthTest", C.c_uint8), ("depthMask", C.c_uint8), ("shader", C.c_char_p)]


class Stats(C.Structure):
    _fields_ = [("trisSubmitted", C.c_uint64), ("trisBinned", C.c_uint64), ("spans", C.c_uint64),
                ("fragsShaded", C.c_uint64), ("kernelLaunches", C.c_uint64), ("drawCalls", C.c_uint64), ("batches", C.c_uint64)]


KERNEL_CLASSES = ["clear", "vertex", "span_count", "scan", "span_fill", "raster", "blit", "composite", "shade", "route", "rendezvous"]


class KernelTimes(C.Structure):
    _fields_ = [("ms", C.c_double * len(KERNEL_CLASSES)), ("launches", C.c_uint64 * len(KERNEL_CLASSES))]


# name -> (restype, argtypes) ; ctx pointer is c_void_p
_P = C.c_void_p
_SZ = C.c_size_t
_U16 = C.c_uint16
_SIGS = {
    "context_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "context_destroy": (None, [_P]),
    "num_threads": (C.c_uint, [_P, C.c_uint]),
    "last_error": (C.c_char_p, [_P]),
    "impl_name": (C.c_char_p, []),
    "set_rcp_table": (C.c_int, [_P, _P]),
    "get_rcp_table": (C.c_int, [_P, _P]),
    "set_rsqrt_table": (C.c_int, [_P, _P]),
    "get_rsqrt_table": (C.c_int, [_P, _P]),
    "create_vbo": (_SZ, [_P]), "create_ibo": (_SZ, [_P]), "create_vao": (_SZ, [_P]),
    "create_ubo": (_SZ, [_P]), "create_texture": (_SZ, [_P]), "create_framebuffer": (_SZ, [_P]),
    "destroy_vbo": (C.c_int, [_P, _SZ]), "destroy_ibo": (C.c_int, [_P, _SZ]), "destroy_vao": (C.c_int, [_P, _SZ]),
    "destroy_ubo": (C.c_int, [_P, _SZ]), "destroy_texture": (C.c_int, [_P, _SZ]), "destroy_framebuffer": (C.c_int, [_P, _SZ]),
    "destroy_shader": (C.c_int, [_P, _SZ]),
    "vbo_init": (C.c_int, [_P, _SZ, _SZ]),
    "vbo_assign": (C.c_int, [_P, _SZ, _P, _SZ, _SZ]),
    "ibo_init": (C.c_int, [_P, _SZ, C.c_uint32, C.c_int, _P]),
    "vao_set_vertex_buffer": (C.c_int, [_P, _SZ, _SZ]),
    "vao_set_index_buffer": (C.c_int, [_P, _SZ, _SZ]),
    "vao_set_num_bindings": (C.c_int, [_P, _SZ, C.c_int]),
    "vao_set_binding": (C.c_int, [_P, _SZ, _SZ, C.c_ssize_t, C.c_ssize_t, C.c_int, C.c_int]),
    "ubo_assign": (C.c_int, [_P, _SZ, _P, _SZ, _SZ]),
    "ubo_bind_texture": (C.c_int, [_P, _SZ, _SZ, _SZ]),
    "ubo_bind_buffer": (C.c_int, [_P, _SZ, _SZ, _SZ, _SZ]),
    "texture_init": (C.c_int, [_P, _SZ, C.c_int, _U16, _U16, _U16]),
    "texture_set_texels": (C.c_int, [_P, _SZ, C.c_int, _U16, _U16, _U16, _U16, _U16, _U16, _P]),
    "texture_read": (C.c_int, [_P, _SZ, _P, _SZ]),
    "fbo_reserve_color_buffers": (C.c_int, [_P, _SZ, C.c_uint]),
    "fbo_attach_color_buffer": (C.c_int, [_P, _SZ, C.c_uint, _SZ]),
    "fbo_attach_depth_buffer": (C.c_int, [_P, _SZ, _SZ]),
    "fbo_valid": (C.c_int, [_P, _SZ]),
    "create_shader": (_SZ, [_P, C.POINTER(VertexShader), C.POINTER(FragmentShader), _SZ]),
    "viewport": (C.c_int, [_P, C.c_int32, C.c_int32, _U16, _U16]),
    "scissor": (C.c_int, [_P, C.c_int32, C.c_int32, _U16, _U16]),
    "clear_color_buffer": (C.c_int, [_P, _SZ, C.c_uint, C.POINTER(C.c_double)]),
    "clear_depth_buffer": (C.c_int, [_P, _SZ, C.c_double]),
    "clear_framebuffer": (C.c_int, [_P, _SZ, C.c_uint, C.POINTER(C.c_double), C.c_double]),
    "draw": (C.c_int, [_P, C.POINTER(Mesh), _SZ, _SZ]),
    "draw_multiple": (C.c_int, [_P, C.POINTER(Mesh), _SZ, _SZ, _SZ]),
    "draw_instanced": (C.c_int, [_P, C.POINTER(Mesh), _SZ, _SZ, _SZ]),
    "blit": (C.c_int, [_P, _SZ, _SZ]),
    "blit_rect": (C.c_int, [_P, _SZ, _SZ] + [_U16] * 8),
    "sync": (C.c_int, [_P]),
    "reset_stats": (C.c_int, [_P]),
    "get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "timer_begin": (C.c_int, [_P]),
    "timer_end": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "profile_begin": (C.c_int, [_P]),
    "profile_end": (C.c_int, [_P, C.POINTER(KernelTimes)]),
}

# product-only entry points (sort-first multi-GPU); the two checker libraries are single-process
_SIGS_PRODUCT = {
    "set_partition": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_uint]),
    "texture_ipc_export": (C.c_int, [_P, _SZ, _P]),
    "texture_ipc_import": (C.c_int, [_P, _SZ, C.c_uint, _P]),
    "composite": (C.c_int, [_P, _SZ, _SZ, C.c_uint]),
    "register_shader_source": (C.c_int, [_P, C.c_char_p]),
    "shader_source_check": (C.c_int, [C.c_char_p, C.c_char_p, _SZ]),
    "geometry_share_create": (C.c_int, [_P, _SZ]),
    "geometry_share_export": (C.c_int, [_P, _P]),
    "geometry_share_import": (C.c_int, [_P, C.c_uint, _P]),
    "geometry_share_address": (C.c_int, [_P, C.POINTER(_P)]),
    "geometry_share_attach": (C.c_int, [_P, C.c_uint, _P]),
    "blit_to_host": (C.c_int, [_P, _P, C.c_int, _U16, _U16, _SZ]),
    "blit_to_host_rect": (C.c_int, [_P, _P, C.c_int, _U16, _U16, _SZ] + [_U16] * 8),
    "present": (C.c_int, [_P, C.c_uint, _P, C.c_int, _U16, _U16, _SZ]),
    "present_wait": (C.c_int, [_P, C.c_uint]),
    "host_alloc": (C.c_int, [_P, C.POINTER(_P), _SZ]),
    "host_free": (C.c_int, [_P, _P]),
}

EXPORTED_SYMBOLS = sorted(list(_SIGS) + list(_SIGS_PRODUCT))

_LIBS = {}


class SlError(RuntimeError):
    pass


def load_library(impl: str):
    """dlopen one of the libraries; raises (never falls back) if it is missing."""
    if impl in _LIBS:
        return _LIBS[impl]
    path = LIB_PATHS[impl]
    if not os.path.exists(path):
        raise SlError(f"{impl}: {path} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)")
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    pre = PREFIX[impl]
    fns = {}
    sigs = dict(_SIGS)
    if impl == "cuda":
        sigs.update(_SIGS_PRODUCT)
    for name, (res, args) in sigs.items():
        fn = getattr(lib, pre + name)
        fn.restype = res
        fn.argtypes = args
        fns[name] = fn
    _LIBS[impl] = fns
    return fns


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _Cmd(C.Structure):
    _fields_ = [("op", C.c_int32), ("pad_", C.c_int32), ("a", C.c_uint64), ("b", C.c_uint64), ("c", C.c_uint64), ("d", C.c_uint64),
                ("p", C.c_void_p)]


class _Api(C.Structure):
    _fields_ = [("ubo_assign", C.c_void_p), ("ubo_bind_texture", C.c_void_p), ("ubo_bind_buffer", C.c_void_p), ("draw", C.c_void_p),
                ("draw_instanced", C.c_void_p)]


_HARNESS = None


def _harness():
    """softlight_b200/libslharness.so: replays a recorded list of per-draw calls from C (measurement harness, see slharness.c)."""
    global _HARNESS
    if _HARNESS is None:
        path = os.path.join(ROOT, "softlight_b200", "libslharness.so")
        if not os.path.exists(path):
            from softlight_b200 import build as B

            B.build_harness()
        lib = C.CDLL(path)
        lib.slh_replay.restype = C.c_size_t
        lib.slh_replay.argtypes = [C.POINTER(_Api), C.c_void_p, C.POINTER(_Cmd), C.c_size_t, C.POINTER(C.c_int)]
        _HARNESS = lib
    return _HARNESS


class CommandList:
    """The per-draw calls of one frame (uniform updates, pointer bindings, draws), recorded once and replayed from C so that a
    measured frame contains what a C/C++ host would execute, not the Python interpreter.  Same call order, same arguments."""

    def __init__(self, ctx: "Context"):
        self.ctx = ctx
        self.cmds = []
        self.keep = []
        self.array = None
        self.api = _Api(*[C.cast(ctx.f[n], C.c_void_p).value for n in ("ubo_assign", "ubo_bind_texture", "ubo_bind_buffer", "draw", "draw_instanced")])

    def ubo_assign(self, uid, blob, offset=0):
        buf = C.create_string_buffer(bytes(blob), len(blob))
        self.keep.append(buf)
        self.cmds.append(_Cmd(0, 0, uid, offset, len(blob), 0, C.cast(buf, C.c_void_p).value))

    def ubo_bind_texture(self, uid, offset, tex):
        self.cmds.append(_Cmd(1, 0, uid, offset, tex, 0, None))

    def ubo_bind_buffer(self, uid, offset, vbo, byte_offset=0):
        self.cmds.append(_Cmd(2, 0, uid, offset, vbo, byte_offset, None))

    def draw(self, mesh, shader, fbo):
        m = Mesh(mesh.vaoId, mesh.elementBegin, mesh.elementEnd, mesh.mode, mesh.materialId)
        self.keep.append(m)
        self.cmds.append(_Cmd(3, 0, shader, fbo, 0, 0, C.addressof(m)))

    def draw_instanced(self, mesh, instances, shader, fbo):
        m = Mesh(mesh.vaoId, mesh.elementBegin, mesh.elementEnd, mesh.mode, mesh.materialId)
        self.keep.append(m)
        self.cmds.append(_Cmd(4, 0, instances, shader, fbo, 0, C.addressof(m)))

    def replay(self):
        if self.array is None or len(self.array) != len(self.cmds):
            self.array = (_Cmd * len(self.cmds))(*self.cmds)
        rc = C.c_int(0)
        bad = _harness().slh_replay(C.byref(self.api), self.ctx.h, self.array, len(self.cmds), C.byref(rc))
        if bad:
            self.ctx._chk(rc.value, f"replayed call #{bad - 1} (op {self.cmds[bad - 1].op})")


class Context:
    """Thin object wrapper over one slcu_context (an SL_Context stand-in)."""

    def __init__(self, impl: str = "cuda", device: int = 0, threads: int = 1):
        self.impl = impl
        self.f = load_library(impl)
        h = C.c_void_p()
        rc = self.f["context_create"](C.byref(h), device)
        if rc != 0 or not h:
            raise SlError(f"{impl}: context_create failed with {rc} (no CPU fallback exists)")
        self.h = h
        self._keep = []
        self._host_allocs = []
        if threads != 1:
            self.num_threads(threads)

    def close(self):
        if self.h:
            for p in self._host_allocs:
                self.f["host_free"](self.h, C.c_void_p(p))
            self._host_allocs = []
            self.f["context_destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc < 0:
            err = self.f["last_error"](self.h)
            raise SlError(f"{self.impl}: {what} failed rc={rc}: {err.decode() if err else ''}")
        return rc

    def impl_name(self) -> str:
        return self.f["impl_name"]().decode()

    def num_threads(self, n: int) -> int:
        return self.f["num_threads"](self.h, n)

    def set_rcp_table(self, table: np.ndarray):
        table = np.ascontiguousarray(table, dtype=np.uint32)
        assert table.size == 2048
        return self.f["set_rcp_table"](self.h, _ptr(table))

    def get_rcp_table(self) -> np.ndarray:
        t = np.zeros(2048, dtype=np.uint32)
        self._chk(self.f["get_rcp_table"](self.h, _ptr(t)), "get_rcp_table")
        return t

    def set_rsqrt_table(self, table: np.ndarray):
        table = np.ascontiguousarray(table, dtype=np.uint32)
        assert table.size == 4096
        return self.f["set_rsqrt_table"](self.h, _ptr(table))

    def get_rsqrt_table(self) -> np.ndarray:
        t = np.zeros(4096, dtype=np.uint32)
        self._chk(self.f["get_rsqrt_table"](self.h, _ptr(t)), "get_rsqrt_table")
        return t

    # -- resources ----------------------------------------------------------
    def create_vbo(self, data: Optional[np.ndarray] = None, nbytes: Optional[int] = None) -> int:
        vid = self.f["create_vbo"](self.h)
        if data is not None:
            data = np.ascontiguousarray(data)
            n = data.nbytes if nbytes is None else nbytes
            self._chk(self.f["vbo_init"](self.h, vid, n), "vbo_init")
            self._chk(self.f["vbo_assign"](self.h, vid, _ptr(data), 0, data.nbytes), "vbo_assign")
        return vid

    def vbo_assign(self, vid: int, data: np.ndarray, offset: int = 0):
        data = np.ascontiguousarray(data)
        self._chk(self.f["vbo_assign"](self.h, vid, _ptr(data), offset, data.nbytes), "vbo_assign")

    def create_ibo(self, indices: np.ndarray) -> int:
        iid = self.f["create_ibo"](self.h)
        indices = np.ascontiguousarray(indices)
        t = {1: VERTEX_DATA_BYTE, 2: VERTEX_DATA_SHORT, 4: VERTEX_DATA_INT}[indices.dtype.itemsize]
        self._chk(self.f["ibo_init"](self.h, iid, indices.size, t, _ptr(indices)), "ibo_init")
        return iid

    def create_vao(self, vbo: int, ibo: Optional[int], bindings) -> int:
        """bindings: list of (offset, stride, dimension, dataType)."""
        vao = self.f["create_vao"](self.h)
        self.f["vao_set_vertex_buffer"](self.h, vao, vbo)
        if ibo is not None:
            self.f["vao_set_index_buffer"](self.h, vao, ibo)
        rc = self.f["vao_set_num_bindings"](self.h, vao, len(bindings))
        if rc != len(bindings):
            raise SlError(f"vao_set_num_bindings -> {rc}")
        for i, (off, stride, dim, typ) in enumerate(bindings):
            self.f["vao_set_binding"](self.h, vao, i, off, stride, dim, typ)
        return vao

    def create_ubo(self, blob: Optional[bytes] = None) -> int:
        uid = self.f["create_ubo"](self.h)
        if blob is not None:
            self.ubo_assign(uid, blob)
        return uid

    def ubo_assign(self, uid: int, blob, offset: int = 0):
        # per-draw call of the multi-draw configs: keep the Python side thin (bytes go straight through ctypes)
        if not isinstance(blob, bytes):
            blob = bytes(blob)
        rc = self.f["ubo_assign"](self.h, uid, blob, offset, len(blob))
        if rc < 0:
            self._chk(rc, "ubo_assign")

    def ubo_bind_texture(self, uid: int, offset: int, tex: int):
        rc = self.f["ubo_bind_texture"](self.h, uid, offset, tex)
        if rc < 0:
            self._chk(rc, "ubo_bind_texture")

    def ubo_bind_buffer(self, uid: int, offset: int, vbo: int, byte_offset: int = 0):
        rc = self.f["ubo_bind_buffer"](self.h, uid, offset, vbo, byte_offset)
        if rc < 0:
            self._chk(rc, "ubo_bind_buffer")

    def create_texture(self, color_type: int, w: int, h: int, d: int = 1, texels: Optional[np.ndarray] = None,
                       order: int = TEXEL_ORDERED) -> int:
        tid = self.f["create_texture"](self.h)
        self._chk(self.f["texture_init"](self.h, tid, color_type, w, h, d), "texture_init")
        if texels is not None:
            texels = np.ascontiguousarray(texels)
            assert texels.nbytes == w * h * d * bytes_per_texel(color_type), (texels.nbytes, w, h, d)
            self._chk(self.f["texture_set_texels"](self.h, tid, order, 0, 0, 0, w, h, d, _ptr(texels)), "set_texels")
        return tid

    def texture_read(self, tid: int, color_type: int, w: int, h: int, d: int = 1) -> np.ndarray:
        nch = texel_channels(color_type)
        out = np.zeros((d, h, w, nch) if d > 1 else (h, w, nch), dtype=texel_dtype(color_type))
        self._chk(self.f["texture_read"](self.h, tid, _ptr(out), out.nbytes), "texture_read")
        return out

    def create_framebuffer(self, color_tex, depth_tex: int) -> int:
        fid = self.f["create_framebuffer"](self.h)
        self._chk(self.f["fbo_reserve_color_buffers"](self.h, fid, len(color_tex)), "reserve_color_buffers")
        for i, t in enumerate(color_tex):
            self._chk(self.f["fbo_attach_color_buffer"](self.h, fid, i, t), "attach_color_buffer")
        self._chk(self.f["fbo_attach_depth_buffer"](self.h, fid, depth_tex), "attach_depth_buffer")
        return fid

    def register_shader_source(self, source: str) -> int:
        """Compile (NVRTC) and register the shaders of a source text at run time; returns how many it registered."""
        return self._chk(self.f["register_shader_source"](self.h, source.encode()), "register_shader_source")

    def destroy(self, kind: str, oid: int) -> int:
        """SL_Context::destroy_<kind>: later ids of that kind move down by one."""
        return self.f["destroy_" + kind](self.h, oid)

    def fbo_valid(self, fid: int) -> int:
        return self.f["fbo_valid"](self.h, fid)

    def create_shader(self, vs_name: str, fs_name: str, ubo: int, *, vs_varyings: int, fs_varyings: int,
                      outputs: int = 1, cull: int = SL_CULL_BACK_FACE, blend: int = SL_BLEND_OFF,
                      depth_test: int = SL_DEPTH_TEST_LESS_THAN, depth_mask: int = SL_DEPTH_MASK_ON) -> int:
        vs = VertexShader(vs_varyings, cull, vs_name.encode())
        fs = FragmentShader(fs_varyings, outputs, blend, depth_test, depth_mask, fs_name.encode())
        self._keep.append((vs, fs))
        return self.f["create_shader"](self.h, C.byref(vs), C.byref(fs), ubo)

    # -- state + work ---------------------------------------------------------
    def viewport(self, x, y, w, h):
        self.f["viewport"](self.h, x, y, w, h)

    def scissor(self, x, y, w, h):
        self.f["scissor"](self.h, x, y, w, h)

    def clear_color_buffer(self, fbo, attachment, rgba):
        arr = (C.c_double * 4)(*rgba)
        self._chk(self.f["clear_color_buffer"](self.h, fbo, attachment, arr), "clear_color_buffer")

    def clear_depth_buffer(self, fbo, depth):
        self._chk(self.f["clear_depth_buffer"](self.h, fbo, depth), "clear_depth_buffer")

    def clear_framebuffer(self, fbo, attachment, rgba, depth):
        arr = (C.c_double * 4)(*rgba)
        self._chk(self.f["clear_framebuffer"](self.h, fbo, attachment, arr, depth), "clear_framebuffer")

    @staticmethod
    def mesh(vao, begin, end, mode=RENDER_MODE_INDEXED_TRIANGLES, material=0) -> Mesh:
        return Mesh(vao, begin, end, mode, material)

    def draw(self, mesh: Mesh, shader: int, fbo: int):
        rc = self.f["draw"](self.h, C.byref(mesh), shader, fbo)
        if rc < 0:
            self._chk(rc, "draw")

    def draw_multiple(self, meshes, shader: int, fbo: int):
        arr = (Mesh * len(meshes))(*meshes)
        self._chk(self.f["draw_multiple"](self.h, arr, len(meshes), shader, fbo), "draw_multiple")

    def draw_instanced(self, mesh: Mesh, instances: int, shader: int, fbo: int):
        self._chk(self.f["draw_instanced"](self.h, C.byref(mesh), instances, shader, fbo), "draw_instanced")

    def blit(self, out_tex, in_tex, rect=None):
        if rect is None:
            self._chk(self.f["blit"](self.h, out_tex, in_tex), "blit")
        else:
            self._chk(self.f["blit_rect"](self.h, out_tex, in_tex, *rect), "blit_rect")

    def sync(self):
        self._chk(self.f["sync"](self.h), "sync")

    # -- sort-first multi-GPU (product only) ---------------------------------
    def set_partition(self, rank: int, world: int, strip_rows: int = 8):
        self._chk(self.f["set_partition"](self.h, rank, world, strip_rows), "set_partition")

    def texture_ipc_export(self, tid: int) -> bytes:
        buf = (C.c_ubyte * 64)()
        self._chk(self.f["texture_ipc_export"](self.h, tid, buf), "texture_ipc_export")
        return bytes(buf)

    def texture_ipc_import(self, tid: int, peer_rank: int, handle: bytes):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._chk(self.f["texture_ipc_import"](self.h, tid, peer_rank, buf), "texture_ipc_import")

    def geometry_share_create(self, arena_bytes: int):
        self._chk(self.f["geometry_share_create"](self.h, arena_bytes), "geometry_share_create")

    def geometry_share_export(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        self._chk(self.f["geometry_share_export"](self.h, buf), "geometry_share_export")
        return bytes(buf)

    def geometry_share_import(self, peer_rank: int, handle: bytes):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._chk(self.f["geometry_share_import"](self.h, peer_rank, buf), "geometry_share_import")

    def geometry_share_address(self) -> int:
        p = C.c_void_p()
        self._chk(self.f["geometry_share_address"](self.h, C.byref(p)), "geometry_share_address")
        return p.value

    def geometry_share_attach(self, peer_rank: int, address: int):
        self._chk(self.f["geometry_share_attach"](self.h, peer_rank, C.c_void_p(address)), "geometry_share_attach")

    def composite(self, src_tex: int, dst_tex: int, root: int = 0):
        self._chk(self.f["composite"](self.h, src_tex, dst_tex, root), "composite")

    def reset_stats(self):
        self.f["reset_stats"](self.h)

    def timer_begin(self):
        self._chk(self.f["timer_begin"](self.h), "timer_begin")

    def timer_end(self) -> float:
        ms = C.c_double(0.0)
        self._chk(self.f["timer_end"](self.h, C.byref(ms)), "timer_end")
        return ms.value

    def vbo_assign_raw(self, vid: int, ptr: int, nbytes: int, offset: int = 0):
        """Upload from a raw host pointer (e.g. pinned memory) without a numpy copy."""
        self._chk(self.f["vbo_assign"](self.h, vid, C.c_void_p(ptr), offset, nbytes), "vbo_assign")

    def host_array(self, shape, dtype) -> np.ndarray:
        """A numpy view of page-locked host memory from slcu_host_alloc (freed with the context; product library only)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._chk(self.f["host_alloc"](self.h, C.byref(p), nbytes), "host_alloc")
        self._host_allocs.append(p.value)
        buf = (C.c_ubyte * max(nbytes, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint8, count=nbytes).view(dtype).reshape(shape)

    def blit_to_host(self, out: np.ndarray, color_type: int, tid: int, src_rect=None, dst_rect=None):
        """SL_Context::blit(SL_TextureView&, textureId[, rects]): `out` is the host view, shape (h, w, channels)."""
        h, w = out.shape[0], out.shape[1]
        if src_rect is None and dst_rect is None:
            self._chk(self.f["blit_to_host"](self.h, out.ctypes.data_as(C.c_void_p), color_type, w, h, tid), "blit_to_host")
        else:
            self._chk(self.f["blit_to_host_rect"](self.h, out.ctypes.data_as(C.c_void_p), color_type, w, h, tid, *src_rect, *dst_rect), "blit_to_host_rect")

    def present(self, slot: int, ptr: int, color_type: int, w: int, h: int, tid: int):
        self._chk(self.f["present"](self.h, slot, C.c_void_p(ptr), color_type, w, h, tid), "present")

    def present_wait(self, slot: int):
        self._chk(self.f["present_wait"](self.h, slot), "present_wait")

    def texture_read_into(self, tid: int, ptr: int, nbytes: int):
        self._chk(self.f["texture_read"](self.h, tid, C.c_void_p(ptr), nbytes), "texture_read")

    def profile_begin(self):
        self._chk(self.f["profile_begin"](self.h), "profile_begin")

    def profile_end(self) -> dict:
        kt = KernelTimes()
        self._chk(self.f["profile_end"](self.h, C.byref(kt)), "profile_end")
        return {k: {"ms": kt.ms[i], "launches": int(kt.launches[i])} for i, k in enumerate(KERNEL_CLASSES)}

    def get_stats(self) -> dict:
        s = Stats()
        self.f["get_stats"](self.h, C.byref(s))
        return {k: getattr(s, k) for k, _ in Stats._fields_}
