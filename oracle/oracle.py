"""ctypes loader for the CPU oracle (oracle/gsplat_oracle.cpp).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libgsplat_oracle.so"

RECORD_DTYPE = np.dtype([("cx", "f4"), ("cy", "f4"), ("m00", "f4"), ("m01", "f4"),
                         ("m10", "f4"), ("m11", "f4"), ("alpha", "f4"), ("pmax", "f4"),
                         ("r", "f4"), ("g", "f4"), ("b", "f4"), ("hpack", "u4")])
RECT_DTYPE = np.dtype([("x0", "u2"), ("x1", "u2"), ("y0", "u2"), ("y1", "u2")])
TILE = 16
KEY_CULLED = 0xFFFFFFFF


class OrcFrame(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("object", C.c_float * 16),
                ("inv_object", C.c_float * 16), ("obj_view", C.c_float * 16),
                ("cam", C.c_float * 3), ("origin", C.c_float * 3),
                ("width", C.c_int32), ("height", C.c_int32), ("sh_order", C.c_int32),
                ("row_rank", C.c_int32), ("row_world", C.c_int32), ("row_group", C.c_int32), ("eps_t", C.c_float)]


class OrcStats(C.Structure):
    _fields_ = [("n_submitted", C.c_int64), ("n_visible", C.c_int64), ("n_instances", C.c_int64),
                ("n_consumed", C.c_int64), ("ms_sort_reference", C.c_double), ("ms_project", C.c_double),
                ("ms_sort", C.c_double), ("ms_bin", C.c_double), ("ms_blend", C.c_double),
                ("ms_total", C.c_double)]


def build(force: bool = False) -> Path:
    src = HERE / "gsplat_oracle.cpp"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.orc_project.restype = C.c_int64
        _lib.orc_bin.restype = C.c_int64
        _lib.orc_blend.restype = C.c_int64
        _lib.orc_blend_depth.restype = C.c_int64
        _lib.orc_det_log.restype = C.c_double
        _lib.orc_det_log.argtypes = [C.c_double]
        _lib.orc_half_to_float.restype = C.c_float
        _lib.orc_half_to_float.argtypes = [C.c_uint16]
        _lib.orc_float_to_half_ru.restype = C.c_uint16
        _lib.orc_float_to_half_ru.argtypes = [C.c_float]
        assert _lib.orc_sizeof_frame() == C.sizeof(OrcFrame)
        assert _lib.orc_sizeof_record() == RECORD_DTYPE.itemsize == 48
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def camera_from_view(view16: np.ndarray) -> np.ndarray:
    v = _c(view16, np.float32)
    out = np.zeros(3, np.float32)
    lib().orc_camera_from_view(_p(v), _p(out))
    return out


def make_frame(frame, cam, origin, sh_order: int, eps_t: float = 1e-5,
               row_rank: int = 0, row_world: int = 1, row_group: int = 1) -> OrcFrame:
    """frame: scene.Frame-like (width,height,view,proj,object,inv_object,obj_view as 16 f32)."""
    f = OrcFrame()
    for name in ("view", "proj", "object", "inv_object", "obj_view"):
        arr = np.asarray(getattr(frame, name), np.float32).reshape(16)
        getattr(f, name)[:] = arr.tolist()
    f.cam[:] = np.asarray(cam, np.float32).tolist()
    f.origin[:] = np.asarray(origin, np.float32).tolist()
    f.width, f.height, f.sh_order = int(frame.width), int(frame.height), int(sh_order)
    f.row_rank, f.row_world, f.row_group, f.eps_t = int(row_rank), int(row_world), int(row_group), float(eps_t)
    return f


def keys(pos: np.ndarray, cam: np.ndarray) -> np.ndarray:
    pos = _c(pos, np.float32); cam = _c(cam, np.float32)
    out = np.empty(pos.shape[0], np.uint32)
    lib().orc_keys(_p(pos), C.c_int64(pos.shape[0]), _p(cam), _p(out))
    return out


def sort(keys_: np.ndarray) -> np.ndarray:
    k = _c(keys_, np.uint32)
    out = np.empty(k.shape[0], np.int32)
    lib().orc_sort(_p(k), C.c_int64(k.shape[0]), _p(out))
    return out


def sort_reference_style(pos: np.ndarray, cam: np.ndarray) -> np.ndarray:
    pos = _c(pos, np.float32); cam = _c(cam, np.float32)
    out = np.empty(pos.shape[0], np.int32)
    lib().orc_sort_reference_style(_p(pos), C.c_int64(pos.shape[0]), _p(cam), _p(out))
    return out


def _cloud_args(cloud):
    pos = _c(cloud.pos, np.float32)
    cd = _c(cloud.cd_h.view(np.uint16), np.uint16)
    al = _c(cloud.alpha, np.float32)
    sc = _c(cloud.scale_h.view(np.uint16), np.uint16)
    orr = _c(cloud.orient_h.view(np.uint16), np.uint16)
    sh = [None if a is None else _c(a.view(np.uint16), np.uint16) for a in (cloud.shx_h, cloud.shy_h, cloud.shz_h)]
    return pos, cd, al, sc, orr, sh[0], sh[1], sh[2]


def project(F: OrcFrame, cloud):
    a = _cloud_args(cloud)
    n = a[0].shape[0]
    k = np.empty(n, np.uint32); recs = np.zeros(n, RECORD_DTYPE); rects = np.zeros(n, RECT_DTYPE)
    vis = np.zeros(n, np.uint8)
    nv = lib().orc_project(C.byref(F), C.c_int64(n), *[_p(x) for x in a], _p(k), _p(recs), _p(rects), _p(vis))
    return dict(keys=k, recs=recs, rects=rects, vis=vis, n_visible=int(nv), _keep=a)


def bin_tiles(F: OrcFrame, order, vis, rects):
    n = order.shape[0]
    tx = (F.width + TILE - 1) // TILE; ty = (F.height + TILE - 1) // TILE
    ts = np.zeros(tx * ty + 1, np.int64)
    order = _c(order, np.int32); vis = _c(vis, np.uint8)
    D = lib().orc_bin(C.byref(F), C.c_int64(n), _p(order), _p(vis), _p(rects), _p(ts), None)
    inst = np.empty(max(int(D), 1), np.int32)
    lib().orc_bin(C.byref(F), C.c_int64(n), _p(order), _p(vis), _p(rects), _p(ts), _p(inst))
    return ts, inst[:int(D)]


def blend(F: OrcFrame, recs, tile_start, inst, rgba=None):
    tx = (F.width + TILE - 1) // TILE; ty = (F.height + TILE - 1) // TILE
    if rgba is None:
        rgba = np.zeros((F.height, F.width, 4), np.float32)
    consumed = np.zeros(tx * ty, np.int64)
    inst = _c(inst, np.int32) if inst.shape[0] else np.zeros(1, np.int32)
    total = lib().orc_blend(C.byref(F), _p(recs), _p(tile_start), _p(inst), _p(rgba), _p(consumed))
    return rgba, consumed, int(total)


DEPTH_NONE, DEPTH_LESS, DEPTH_LEQUAL = 0, 1, 2


def window_depth(F: OrcFrame, cloud, depth_range=(0.0, 1.0)) -> np.ndarray:
    """Window-space depth of every splat's quad (the centre's, SRC.h:278-282) under glDepthRange = depth_range."""
    pos = _c(cloud.pos, np.float32)
    dr = np.asarray(depth_range, np.float32)
    out = np.zeros(pos.shape[0], np.float32)
    lib().orc_window_depth(C.byref(F), C.c_int64(pos.shape[0]), _p(pos), _p(dr), _p(out))
    return out


def blend_depth(F: OrcFrame, recs, tile_start, inst, zw, scene_depth, depth_func, rgba=None):
    """orc_blend with the reference's depth test against an existing scene depth buffer (R.C:608-610)."""
    tx = (F.width + TILE - 1) // TILE; ty = (F.height + TILE - 1) // TILE
    if rgba is None:
        rgba = np.zeros((F.height, F.width, 4), np.float32)
    consumed = np.zeros(tx * ty, np.int64)
    inst = _c(inst, np.int32) if inst.shape[0] else np.zeros(1, np.int32)
    zw = _c(zw, np.float32); sd = _c(scene_depth, np.float32)
    assert sd.shape == (F.height, F.width)
    total = lib().orc_blend_depth(C.byref(F), _p(recs), _p(tile_start), _p(inst), _p(rgba), _p(consumed),
                                  _p(zw), _p(sd), C.c_int(int(depth_func)))
    return rgba, consumed, int(total)


def blend_bruteforce(F: OrcFrame, order, vis, recs):
    rgba = np.zeros((F.height, F.width, 4), np.float32)
    order = _c(order, np.int32)
    lib().orc_blend_bruteforce(C.byref(F), C.c_int64(order.shape[0]), _p(order), _p(vis), _p(recs), _p(rgba))
    return rgba


def render(F: OrcFrame, cloud, time_reference_sort: bool = False):
    a = _cloud_args(cloud)
    rgba = np.zeros((F.height, F.width, 4), np.float32)
    st = OrcStats()
    lib().orc_render(C.byref(F), C.c_int64(a[0].shape[0]), *[_p(x) for x in a], _p(rgba), C.byref(st),
                     C.c_int(1 if time_reference_sort else 0))
    return rgba, {k: getattr(st, k) for k, _ in OrcStats._fields_}


def pipeline(F: OrcFrame, cloud, scene_depth=None, depth_func=DEPTH_NONE, depth_range=(0.0, 1.0)):
    """All intermediates, for stage-by-stage parity checks.  scene_depth ([H,W] f32) + depth_func: SURVEY 8f-3."""
    pr = project(F, cloud)
    order = sort(pr["keys"])
    ts, inst = bin_tiles(F, order, pr["vis"], pr["rects"])
    if scene_depth is not None and depth_func != DEPTH_NONE:
        zw = window_depth(F, cloud, depth_range)
        rgba, consumed, total = blend_depth(F, pr["recs"], ts, inst, zw, scene_depth, depth_func)
        return dict(pr, order=order, tile_start=ts, inst=inst, rgba=rgba, consumed=consumed, n_consumed=total, zw=zw)
    rgba, consumed, total = blend(F, pr["recs"], ts, inst)
    return dict(pr, order=order, tile_start=ts, inst=inst, rgba=rgba, consumed=consumed, n_consumed=total)


def wire_vertices(F: OrcFrame, cloud):
    """the reference's wire vertex shader for the 8 outline vertices of every splat: ([8n,4] gl_Position, [8n,3] Cd)."""
    pos = _c(cloud.pos, np.float32)
    cd = _c(cloud.cd_h.view(np.uint16), np.uint16); sc = _c(cloud.scale_h.view(np.uint16), np.uint16)
    orr = _c(cloud.orient_h.view(np.uint16), np.uint16)
    n = pos.shape[0]
    verts = np.zeros((8 * n, 4), np.float32); cols = np.zeros((8 * n, 3), np.float32)
    lib().orc_wire_vertices(C.byref(F), C.c_int64(n), _p(pos), _p(cd), _p(sc), _p(orr), _p(verts), _p(cols))
    return verts, cols


def wire_overlay(verts: np.ndarray, cloud, width: int, height: int, rgba=None):
    cd = _c(cloud.cd_h.view(np.uint16), np.uint16)
    if rgba is None:
        rgba = np.zeros((height, width, 4), np.float32)
    v = _c(verts, np.float32)
    lib().orc_wire_overlay(C.c_int64(v.shape[0] // 8), _p(v), _p(cd), C.c_int(width), C.c_int(height), _p(rgba))
    return rgba


def build_prim(pos: np.ndarray):
    pos = _c(pos, np.float32); n = pos.shape[0]
    v2p = np.empty(n, np.int32); bary = np.zeros(3, np.float32); bbox = np.zeros(6, np.float32)
    lib().orc_build_prim(_p(pos), C.c_int64(n), _p(v2p), _p(bary), _p(bbox))
    return v2p, bary, bbox


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int | None = None) -> int:
    """Use n host threads (default: every core this process may run on), whatever OMP_NUM_THREADS says."""
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()
