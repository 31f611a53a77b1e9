"""ctypes mirror of the reference's ``GSplatRenderer`` call surface over libgsplat_b200.so.

Method names and argument meaning follow /root/reference/gsplat_plugin/include/GSplatRenderer.h:29-56
(``registerUpdate``, ``includeInRenderPass``, ``flushEntriesForMatchingDetail``,
``generateRenderGeometry``, ``render``, ``postRender``, ``setRenderingEnabled``,
``setExplicitCameraPos``, ``setSphericalHarmonicsOrder``) so tests and bench.py read like a caller of
the reference.  This file is plumbing for tests/bench; the product is the C ABI (include/gsplat_b200.h).
There is NO CPU fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libgsplat_b200.so"
ID_MAX = 128
REFERENCE_SPLAT_CAP = 8388607

OPT_SPLAT_CAP, OPT_EPS_T, OPT_STAGE_TIMING, OPT_KEEP_INTERMEDIATES, OPT_DEPTH_CHUNKS, OPT_COMPACT, OPT_CHUNK_SHIFT, OPT_HOST_DIRECT, OPT_LAZY_PROJECT = 1, 2, 3, 4, 5, 6, 7, 8, 9
(DBG_KEYS_UNSORTED, DBG_ORDER, DBG_RECORDS, DBG_RECTS, DBG_TILE_RANGES, DBG_INSTANCES,
 DBG_FRAMEBUFFER, DBG_KEYS_SORTED, DBG_TILE_CONSUMED, DBG_TRECTS) = range(10)

RECORD_DTYPE = np.dtype([("cx", "f4"), ("cy", "f4"), ("m00", "f4"), ("m01", "f4"),
                         ("m10", "f4"), ("m11", "f4"), ("alpha", "f4"), ("pmax", "f4"),
                         ("r", "f4"), ("g", "f4"), ("b", "f4"), ("hpack", "u4")])
RECT_DTYPE = np.dtype([("x0", "u2"), ("x1", "u2"), ("y0", "u2"), ("y1", "u2")])

# every symbol include/gsplat_b200.h declares
EXPORTS = ["gsb_abi_version", "gsb_create", "gsb_destroy", "gsb_last_error", "gsb_register_update",
           "gsb_include_in_render_pass", "gsb_flush_entries_for_matching_detail",
           "gsb_generate_render_geometry", "gsb_render", "gsb_post_render", "gsb_set_rendering_enabled",
           "gsb_set_explicit_camera_pos", "gsb_set_spherical_harmonics_order", "gsb_set_option",
           "gsb_get_stats", "gsb_set_stream", "gsb_synchronize", "gsb_device_framebuffer",
           "gsb_registry_size", "gsb_debug_fetch", "gsb_debug_sort_pairs", "gsb_debug_exclusive_scan",
           "gsb_ipc_export_frame", "gsb_ipc_open", "gsb_ipc_close", "gsb_copy_to_host",
           "gsb_update_from_attributes", "gsb_debug_fetch_entry", "gsb_host_register", "gsb_host_unregister",
           "gsb_render_wireframe", "gsb_wire_device_vertices"]


class GsbError(RuntimeError):
    pass


class PrimKey(C.Structure):
    _fields_ = [("gdp", C.c_uint64), ("vtx0", C.c_int64), ("version", C.c_int64 * 4)]


class FrameC(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("object", C.c_float * 16),
                ("inv_object", C.c_float * 16), ("obj_view", C.c_float * 16),
                ("width", C.c_int32), ("height", C.c_int32), ("is_object_level", C.c_int32),
                ("row_rank", C.c_int32), ("row_world", C.c_int32), ("row_group", C.c_int32),
                ("depth_func", C.c_int32), ("gl_depth_texture", C.c_uint32), ("depth_range", C.c_float * 2),
                ("scene_depth", C.c_void_p)]


DEPTH_NONE, DEPTH_LESS, DEPTH_LEQUAL = 0, 1, 2
WARN_OBJECT_LEVEL = 1
ACT_NONE, ACT_INRIA = 0, 1


class TargetC(C.Structure):
    _fields_ = [("device_rgba", C.c_void_p), ("host_rgba", C.c_void_p),
                ("gl_texture", C.c_uint32), ("flags", C.c_uint32), ("final_rgba", C.c_void_p)]


class WireTargetC(C.Structure):
    _fields_ = [("device_vertices", C.c_void_p), ("device_colors", C.c_void_p), ("host_vertices", C.c_void_p),
                ("host_colors", C.c_void_p), ("overlay_rgba", C.c_void_p), ("overlay_host_rgba", C.c_void_p)]


class RawAttributesC(C.Structure):
    _fields_ = [("count", C.c_int64), ("P", C.c_void_p), ("Cd", C.c_void_p), ("opacity", C.c_void_p), ("Alpha", C.c_void_p),
                ("scale", C.c_void_p), ("orient", C.c_void_p), ("sh_coefficients", C.c_void_p),
                ("sh_coefficients_len", C.c_int32), ("activation", C.c_int32),
                ("sh", C.c_void_p * 15), ("f_rest", C.c_void_p * 45),
                ("has_sh_order", C.c_int32), ("sh_order", C.c_int32), ("has_explicit_camera", C.c_int32),
                ("explicit_camera", C.c_float * 3)]


class UpdateResultC(C.Structure):
    _fields_ = [("id", C.c_char * ID_MAX), ("sh_order", C.c_int32), ("sh_order_invalid", C.c_int32),
                ("sh_data_found", C.c_int32), ("set_explicit_camera", C.c_int32),
                ("explicit_camera", C.c_float * 3), ("barycentre", C.c_float * 3)]


class StatsC(C.Structure):
    _fields_ = [("n_submitted", C.c_int64), ("n_visible", C.c_int64), ("n_instances", C.c_int64),
                ("n_consumed", C.c_int64), ("rendered", C.c_int32), ("repacked", C.c_int32),
                ("sh_order_used", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("tiles_x", C.c_int32), ("tiles_y", C.c_int32), ("launches", C.c_int32),
                ("depth_chunks", C.c_int32), ("warnings", C.c_int32),
                ("camera", C.c_float * 3), ("origin", C.c_float * 3),
                ("ms_project", C.c_float), ("ms_sort", C.c_float), ("ms_bin", C.c_float),
                ("ms_blend", C.c_float), ("ms_copy", C.c_float), ("ms_total", C.c_float),
                ("ms_records", C.c_float), ("reserved1", C.c_float), ("n_live", C.c_int64)]


_lib = None


def load_library() -> C.CDLL:
    """Loads the in-tree CUDA library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise GsbError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C houdini-gsplat-renderer_b200/csrc); there is no CPU fallback")
        lib = C.CDLL(str(LIB_PATH))
        lib.gsb_last_error.restype = C.c_char_p
        lib.gsb_device_framebuffer.restype = C.c_void_p
        lib.gsb_device_framebuffer.argtypes = [C.c_void_p]
        lib.gsb_set_option.argtypes = [C.c_void_p, C.c_int, C.c_double]
        lib.gsb_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.gsb_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        lib.gsb_debug_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p]
        lib.gsb_debug_exclusive_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        lib.gsb_register_update.argtypes = [C.c_void_p, C.POINTER(PrimKey), C.c_int64, C.c_void_p] + [C.c_void_p] * 8 + [C.c_char_p]
        lib.gsb_render.argtypes = [C.c_void_p, C.POINTER(FrameC), C.POINTER(TargetC)]
        lib.gsb_get_stats.argtypes = [C.c_void_p, C.POINTER(StatsC)]
        for name in ("gsb_include_in_render_pass", "gsb_flush_entries_for_matching_detail"):
            getattr(lib, name).argtypes = [C.c_void_p, C.c_char_p]
        for name in ("gsb_destroy", "gsb_generate_render_geometry", "gsb_post_render", "gsb_synchronize",
                     "gsb_registry_size"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.gsb_set_rendering_enabled.argtypes = [C.c_void_p, C.c_int]
        lib.gsb_set_spherical_harmonics_order.argtypes = [C.c_void_p, C.c_int]
        lib.gsb_set_explicit_camera_pos.argtypes = [C.c_void_p, C.c_void_p]
        lib.gsb_ipc_export_frame.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.gsb_ipc_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.gsb_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
        lib.gsb_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.gsb_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        lib.gsb_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
        lib.gsb_render_wireframe.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(FrameC), C.POINTER(WireTargetC)]
        lib.gsb_wire_device_vertices.restype = C.c_void_p
        lib.gsb_wire_device_vertices.argtypes = [C.c_void_p]
        lib.gsb_update_from_attributes.argtypes = [C.c_void_p, C.POINTER(PrimKey), C.POINTER(RawAttributesC), C.POINTER(UpdateResultC)]
        lib.gsb_debug_fetch_entry.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def frame_to_c(frame, row_rank: int = 0, row_world: int = 1, row_group: int = 1, scene_depth: int | None = None,
               depth_func: int = 0, depth_range=(0.0, 1.0)) -> FrameC:
    f = FrameC()
    for name in ("view", "proj", "object", "inv_object", "obj_view"):
        getattr(f, name)[:] = np.asarray(getattr(frame, name), np.float32).reshape(16).tolist()
    f.width, f.height = int(frame.width), int(frame.height)
    f.is_object_level = int(bool(getattr(frame, "is_object_level", False)))
    f.row_rank, f.row_world, f.row_group = int(row_rank), int(row_world), int(row_group)
    f.depth_func, f.gl_depth_texture, f.scene_depth = int(depth_func), 0, scene_depth
    f.depth_range[:] = [float(depth_range[0]), float(depth_range[1])]
    return f


class GSplatRenderer:
    """One context per GPU (the reference is a process-wide singleton, R.h:29-32)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.gsb_create(int(device), C.byref(h))
        if rc != 0:
            raise GsbError(f"gsb_create failed ({rc}): {self._lib.gsb_last_error().decode()}")
        self._h = h
        self._keep = {}

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gsb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise GsbError(f"{what} failed ({rc}): {self._lib.gsb_last_error().decode()}")

    # ---- the reference's surface ---------------------------------------------------------
    def registerUpdate(self, gdp: int, gversion, gvtx: int, cloud, origin=None) -> str:
        """GSplatRenderer::registerUpdate (R.h:34-47).  ``cloud`` carries the SoA arrays of one prim;
        ``origin`` is the prim's baryCenter (GR_GSplat.C:428)."""
        key = PrimKey(int(gdp), int(gvtx), (C.c_int64 * 4)(*[int(v) for v in gversion]))
        origin = cloud.barycentre() if origin is None else np.asarray(origin, np.float32)
        origin = np.ascontiguousarray(origin, np.float32)
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a).view(dt)
        arrs = [np.ascontiguousarray(cloud.pos, np.float32), c(cloud.cd_h, np.uint16),
                np.ascontiguousarray(cloud.alpha, np.float32), c(cloud.scale_h, np.uint16),
                c(cloud.orient_h, np.uint16), c(cloud.shx_h, np.uint16), c(cloud.shy_h, np.uint16),
                c(cloud.shz_h, np.uint16)]
        out = C.create_string_buffer(ID_MAX)
        self._ck(self._lib.gsb_register_update(self._h, C.byref(key), int(cloud.n), _ptr(origin),
                                               *[_ptr(a) for a in arrs], out), "gsb_register_update")
        return out.value.decode()

    def update(self, gdp: int, gversion, gvtx: int, attrs: dict, activation: int = 0) -> dict:
        """GR_PrimGsplat::update (GR_GSplat.C:191-458) on the GPU: ``attrs`` maps Houdini attribute names to fp32 numpy
        arrays (P, Cd, opacity, Alpha, scale, orient, sh_coefficients | sh1..sh15 | f_rest_0..f_rest_44) plus the detail
        attributes gsplat__sh_order (int) and gsplat__explicit_camera_pos (3 floats).  Returns what update() leaves
        for render(): id, sh_order, explicit camera, barycentre."""
        key = PrimKey(int(gdp), int(gvtx), (C.c_int64 * 4)(*[int(v) for v in gversion]))
        keep = []

        def arr(name, shape_tail):
            a = attrs.get(name)
            if a is None:
                return None
            a = np.ascontiguousarray(a, np.float32)
            assert a.shape[1:] == shape_tail, (name, a.shape)
            keep.append(a)
            return a.ctypes.data

        P = np.ascontiguousarray(attrs["P"], np.float32); keep.append(P)
        ra = RawAttributesC()
        ra.count = P.shape[0]; ra.P = P.ctypes.data
        ra.activation = int(activation)                  # ACT_INRIA: attrs are raw INRIA PLY columns (ply.inria_raw_attributes)
        ra.Cd = arr("Cd", (3,)); ra.opacity = arr("opacity", ()); ra.Alpha = arr("Alpha", ())
        ra.scale = arr("scale", (3,)); ra.orient = arr("orient", (4,))
        shc = attrs.get("sh_coefficients")
        if shc is not None:
            shc = np.ascontiguousarray(shc, np.float32); keep.append(shc)
            ra.sh_coefficients = shc.ctypes.data; ra.sh_coefficients_len = shc.shape[1]
        for j in range(15):
            ra.sh[j] = arr(f"sh{j + 1}", (3,))
        for j in range(45):
            ra.f_rest[j] = arr(f"f_rest_{j}", ())
        if "gsplat__sh_order" in attrs:
            ra.has_sh_order = 1; ra.sh_order = int(attrs["gsplat__sh_order"])
        if "gsplat__explicit_camera_pos" in attrs:
            ra.has_explicit_camera = 1
            ra.explicit_camera[:] = [float(v) for v in attrs["gsplat__explicit_camera_pos"]]
        out = UpdateResultC()
        self._ck(self._lib.gsb_update_from_attributes(self._h, C.byref(key), C.byref(ra), C.byref(out)),
                 "gsb_update_from_attributes")
        return dict(id=out.id.decode(), sh_order=out.sh_order, sh_order_invalid=bool(out.sh_order_invalid),
                    sh_data_found=bool(out.sh_data_found), set_explicit_camera=bool(out.set_explicit_camera),
                    explicit_camera=np.array(out.explicit_camera[:], np.float32),
                    barycentre=np.array(out.barycentre[:], np.float32))

    def fetch_entry(self, registry_id: str, which: int) -> np.ndarray:
        need = C.c_uint64(0)
        self._ck(self._lib.gsb_debug_fetch_entry(self._h, registry_id.encode(), which, None, 0, C.byref(need)), "gsb_debug_fetch_entry")
        dt = np.float32 if which in (0, 2) else np.uint16
        out = np.zeros(need.value // np.dtype(dt).itemsize, dt)
        if need.value:
            self._ck(self._lib.gsb_debug_fetch_entry(self._h, registry_id.encode(), which, _ptr(out), need.value, C.byref(need)),
                     "gsb_debug_fetch_entry")
        return out

    def includeInRenderPass(self, registry_id: str):
        self._ck(self._lib.gsb_include_in_render_pass(self._h, registry_id.encode()), "gsb_include_in_render_pass")

    def flushEntriesForMatchingDetail(self, registry_id: str):
        self._ck(self._lib.gsb_flush_entries_for_matching_detail(self._h, registry_id.encode()),
                 "gsb_flush_entries_for_matching_detail")

    def generateRenderGeometry(self):
        self._ck(self._lib.gsb_generate_render_geometry(self._h), "gsb_generate_render_geometry")

    def render(self, frame, host_rgba: np.ndarray | None = None, device_rgba: int | None = None,
               row_rank: int = 0, row_world: int = 1, row_group: int = 1, final_rgba: int | None = None,
               scene_depth: int | None = None, depth_func: int = 0, depth_range=(0.0, 1.0)):
        """GSplatRenderer::render(r, isObjectLevel) (R.C:534-658).  ``frame`` supplies what the reference
        reads from RE_Render / glH_*; the finished frame goes to the library's device buffer, to
        ``device_rgba`` (a CUDA pointer) and/or to ``host_rgba`` ([H,W,4] f32, D2H inside the call).
        ``scene_depth`` (CUDA pointer to [H,W] f32 window depth) + ``depth_func``: the reference's depth test against the
        scene already in the viewport (R.C:608-610)."""
        fc = frame if isinstance(frame, FrameC) else frame_to_c(frame, row_rank, row_world, row_group, scene_depth,
                                                                depth_func, depth_range)
        t = TargetC(device_rgba, None if host_rgba is None else host_rgba.ctypes.data, 0, 0, final_rgba)
        self._ck(self._lib.gsb_render(self._h, C.byref(fc), C.byref(t)), "gsb_render")

    def renderWireframe(self, registry_id: str, frame, count: int, overlay_rgba: int | None = None,
                        overlay_host: np.ndarray | None = None):
        """The wire pass of GR_PrimGsplat::render (GR_GSplat.C:474-483): returns ([8n,4] gl_Position, [8n,3] colour) of the
        outline's line vertices; overlay_rgba (CUDA pointer): the outlines are also drawn into that frame."""
        fc = frame if isinstance(frame, FrameC) else frame_to_c(frame)
        verts = np.zeros((8 * count, 4), np.float32); cols = np.zeros((8 * count, 3), np.float32)
        t = WireTargetC(None, None, verts.ctypes.data, cols.ctypes.data, overlay_rgba,
                        None if overlay_host is None else overlay_host.ctypes.data)
        self._ck(self._lib.gsb_render_wireframe(self._h, registry_id.encode(), C.byref(fc), C.byref(t)), "gsb_render_wireframe")
        return verts, cols

    def postRender(self):
        self._ck(self._lib.gsb_post_render(self._h), "gsb_post_render")

    def setRenderingEnabled(self, enabled: bool):
        self._ck(self._lib.gsb_set_rendering_enabled(self._h, int(bool(enabled))), "gsb_set_rendering_enabled")

    def setExplicitCameraPos(self, pos):
        p = np.ascontiguousarray(pos, np.float32)
        self._ck(self._lib.gsb_set_explicit_camera_pos(self._h, _ptr(p)), "gsb_set_explicit_camera_pos")

    def setSphericalHarmonicsOrder(self, order: int):
        self._ck(self._lib.gsb_set_spherical_harmonics_order(self._h, int(order)), "gsb_set_spherical_harmonics_order")

    # ---- additions -------------------------------------------------------------------------
    def set_option(self, opt: int, value: float):
        self._ck(self._lib.gsb_set_option(self._h, int(opt), float(value)), "gsb_set_option")

    def set_stream(self, cuda_stream: int | None):
        self._ck(self._lib.gsb_set_stream(self._h, cuda_stream), "gsb_set_stream")

    def synchronize(self):
        self._ck(self._lib.gsb_synchronize(self._h), "gsb_synchronize")

    def device_framebuffer(self) -> int:
        return int(self._lib.gsb_device_framebuffer(self._h) or 0)

    def registry_size(self) -> int:
        return int(self._lib.gsb_registry_size(self._h))

    # ---- multi-GPU frame sharing (CUDA IPC; the blend writes finished tiles into the display rank's frame) ----
    def ipc_export_frame(self, width: int, height: int):
        """Display rank: allocate the shared frame; returns (64-byte handle, local device pointer)."""
        h = (C.c_ubyte * 64)(); p = C.c_void_p()
        self._ck(self._lib.gsb_ipc_export_frame(self._h, int(width), int(height), h, C.byref(p)), "gsb_ipc_export_frame")
        return bytes(h), int(p.value)

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_ubyte * 64).from_buffer_copy(handle); p = C.c_void_p()
        self._ck(self._lib.gsb_ipc_open(self._h, buf, C.byref(p)), "gsb_ipc_open")
        return int(p.value)

    def ipc_close(self, ptr: int):
        self._ck(self._lib.gsb_ipc_close(self._h, ptr), "gsb_ipc_close")

    def copy_to_host(self, device_ptr: int, host: np.ndarray):
        self._ck(self._lib.gsb_copy_to_host(self._h, device_ptr, host.ctypes.data, host.nbytes), "gsb_copy_to_host")

    def host_register(self, arr: np.ndarray):
        """Page-lock + map a page-aligned host array (cudaHostRegister) so the blend kernel can store finished tiles into it."""
        self._ck(self._lib.gsb_host_register(self._h, arr.ctypes.data, arr.nbytes), "gsb_host_register")

    def host_unregister(self, arr: np.ndarray):
        self._ck(self._lib.gsb_host_unregister(self._h, arr.ctypes.data), "gsb_host_unregister")

    def stats(self) -> dict:
        s = StatsC()
        self._ck(self._lib.gsb_get_stats(self._h, C.byref(s)), "gsb_get_stats")
        d = {}
        for k, _ in StatsC._fields_:
            v = getattr(s, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d

    def draw(self, ids, frame, **kw):
        """One viewport redraw as DM_GSplatHook.C:30-39 + GR_GSplat.C:485 drive it."""
        for i in ids:
            self.includeInRenderPass(i)
        self.generateRenderGeometry()
        self.render(frame, **kw)
        self.postRender()

    def fetch(self, which: int) -> np.ndarray:
        need = C.c_uint64(0)
        self._ck(self._lib.gsb_debug_fetch(self._h, which, None, 0, C.byref(need)), "gsb_debug_fetch")
        dt = {DBG_RECORDS: RECORD_DTYPE, DBG_RECTS: RECT_DTYPE, DBG_FRAMEBUFFER: np.float32}.get(which, np.uint32)
        out = np.zeros(need.value // np.dtype(dt).itemsize, dt)
        if need.value:
            self._ck(self._lib.gsb_debug_fetch(self._h, which, _ptr(out), need.value, C.byref(need)), "gsb_debug_fetch")
        if which == DBG_TILE_RANGES:
            out = out.reshape(-1, 2)
        return out

    def sort_pairs(self, keys: np.ndarray, vals: np.ndarray, begin_bit: int = 0, end_bit: int = 32):
        k = np.ascontiguousarray(keys, np.uint32); v = np.ascontiguousarray(vals, np.uint32)
        ko = np.empty_like(k); vo = np.empty_like(v)
        self._ck(self._lib.gsb_debug_sort_pairs(self._h, _ptr(k), _ptr(v), k.shape[0], begin_bit, end_bit,
                                                _ptr(ko), _ptr(vo)), "gsb_debug_sort_pairs")
        return ko, vo

    def exclusive_scan(self, x: np.ndarray):
        a = np.ascontiguousarray(x, np.uint32); o = np.empty_like(a); t = C.c_uint64(0)
        self._ck(self._lib.gsb_debug_exclusive_scan(self._h, _ptr(a), a.shape[0], _ptr(o), C.byref(t)),
                 "gsb_debug_exclusive_scan")
        return o, int(t.value)
