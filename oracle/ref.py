"""ctypes loader for oracle/_ref/libgsplat_ref.so — the reference's own GLSL text compiled for the host
(oracle/ref_harness.cpp, oracle/build_ref.py).

TEST INFRASTRUCTURE ONLY — used by tests/ (to pin the oracle and the CUDA path to the reference's text) and by
bench.py's --impl reference / cpu_baseline legs.  The product package never imports this module.
The library is built where /root/reference exists; elsewhere (the GPU box) the prebuilt .so is loaded.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import build_ref

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_ref" / "libgsplat_ref.so"


class RefFrame(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("object", C.c_float * 16),
                ("inv_object", C.c_float * 16), ("obj_view", C.c_float * 16),
                ("cam", C.c_float * 3), ("origin", C.c_float * 3),
                ("width", C.c_int32), ("height", C.c_int32), ("sh_order", C.c_int32), ("depth_range", C.c_float * 2)]


_lib = None


def available() -> bool:
    try:
        return build_ref.build() is not None
    except Exception:
        return LIB_PATH.exists()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if build_ref.build() is None:
            raise RuntimeError("oracle/_ref/libgsplat_ref.so is not built and /root/reference is not available")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.ref_num_threads.restype = C.c_int
        _lib.ref_closest_sqrt_pow2.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def set_num_threads(n: int | None = None) -> int:
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    lib().ref_set_num_threads(C.c_int(int(n)))
    return int(lib().ref_num_threads())


class Bound:
    """One scene + frame bound to the compiled shaders (their uniforms and samplers are process-wide state, like a GL
    context's): keeps the arrays alive."""

    def __init__(self, cloud, frame, cam, origin, sh_order: int, order=None, depth_range=(0.0, 1.0)):
        self.n = cloud.n
        self.pos = _c(cloud.pos, np.float32)
        self.cd = _c(cloud.cd_h.view(np.uint16), np.uint16)
        self.alpha = _c(cloud.alpha, np.float32)
        self.scale = _c(cloud.scale_h.view(np.uint16), np.uint16)
        self.orient = _c(cloud.orient_h.view(np.uint16), np.uint16)
        self.sh = [None if a is None else _c(a.view(np.uint16), np.uint16) for a in (cloud.shx_h, cloud.shy_h, cloud.shz_h)]
        self.cam = np.asarray(cam, np.float32)
        self.order = argsort_by_distance(self.pos, self.cam) if order is None else _c(order, np.int32)
        f = RefFrame()
        for name in ("view", "proj", "object", "inv_object", "obj_view"):
            getattr(f, name)[:] = np.asarray(getattr(frame, name), np.float32).reshape(16).tolist()
        f.cam[:] = self.cam.tolist()
        f.origin[:] = np.asarray(origin, np.float32).tolist()
        f.width, f.height, f.sh_order = int(frame.width), int(frame.height), int(sh_order)
        f.depth_range[:] = [float(depth_range[0]), float(depth_range[1])]
        self.f = f
        self.W, self.H = int(frame.width), int(frame.height)
        lib().ref_bind(C.c_int64(self.n), _p(self.pos), _p(self.cd), _p(self.alpha), _p(self.scale), _p(self.orient),
                       _p(self.sh[0]), _p(self.sh[1]), _p(self.sh[2]), _p(self.order), C.byref(self.f))

    def vertices(self, i0: int = 0, i1: int | None = None) -> np.ndarray:
        """Vertex shader over instances [i0, i1) (positions in the z-order texture) x 6 vertices:
        [n, 6, 12] = gl_Position(4), vsOut.pos(4), vsOut.color(3), vsOut.opacity."""
        i1 = self.n if i1 is None else i1
        out = np.zeros((i1 - i0, 6, 12), np.float32)
        lib().ref_vs_batch(C.c_int64(i0), C.c_int64(i1), _p(out))
        return out

    def draw(self, rgba=None, scene_depth=None, depth_func: int = 0, unsafe_tol: float | None = None):
        """The instanced draw of R.C:647 into rgba ([H,W,4] f32, row 0 = bottom; zeros if None).
        Returns (rgba, stats dict[, unsafe mask])."""
        if rgba is None:
            rgba = np.zeros((self.H, self.W, 4), np.float32)
        st = np.zeros(2, np.int64)
        sd = None if scene_depth is None else _c(scene_depth, np.float32)
        if unsafe_tol is None:
            lib().ref_draw(_p(rgba), _p(sd), C.c_int(int(depth_func)), _p(st))
            return rgba, {"fragments_shaded": int(st[0]), "fragments_kept": int(st[1])}
        unsafe = np.zeros((self.H, self.W), np.uint8)
        lib().ref_draw_ex(_p(rgba), _p(sd), C.c_int(int(depth_func)), _p(st), _p(unsafe), C.c_float(float(unsafe_tol)))
        return rgba, {"fragments_shaded": int(st[0]), "fragments_kept": int(st[1])}, unsafe.astype(bool)


def argsort_by_distance(pos: np.ndarray, cam: np.ndarray) -> np.ndarray:
    pos = _c(pos, np.float32); cam = _c(cam, np.float32)
    out = np.empty(pos.shape[0], np.int32)
    lib().ref_argsort_by_distance(_p(pos), C.c_int64(pos.shape[0]), _p(cam), _p(out))
    return out


def fragment(pos, color, opacity: float):
    """One fragment-shader invocation: returns (rgba or None if discarded)."""
    p = _c(pos, np.float32); c = _c(color, np.float32); out = np.zeros(4, np.float32)
    d = lib().ref_fs_main(_p(p), _p(c), C.c_float(float(opacity)), _p(out))
    return None if d else out


def wire_vertex(vertex_id: int, P, Cd, scale, orient):
    """One wireframe vertex-shader invocation (SRC.h:22-90): (gl_Position[4], color[3])."""
    a = [_c(x, np.float32) for x in (P, Cd, scale, orient)]
    g = np.zeros(4, np.float32); col = np.zeros(3, np.float32)
    lib().ref_wire_vs(C.c_int(int(vertex_id)), *[_p(x) for x in a], _p(g), _p(col))
    return g, col


def closest_sqrt_pow2(n: int) -> int:
    return int(lib().ref_closest_sqrt_pow2(C.c_int(int(n))))
