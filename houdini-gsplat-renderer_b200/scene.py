"""Synthetic splat clouds and cameras for the BASELINE.json configs.

Attribute conventions are the reference's input contract (what ``GR_PrimGsplat::update`` hands to
``GSplatRenderer::registerUpdate``, /root/reference/gsplat_plugin/src/GR_GSplat.C:302-372,423-436):
``P`` f32x3; ``Cd``/``scale`` half x3; ``orient`` half x4 stored (x,y,z,w); ``opacity`` f32;
SH as three arrays of 4x4 half matrices (R/G/B), coefficient j at (j/4, j%4), [3][3] unused.

Everything here is host-side numpy; it is input generation, not part of the hot path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

CHUNK = 1 << 20  # fixed chunk size => the stream of every attribute is reproducible by construction

# seeds from SURVEY.md §8d
SEEDS = {"config1_10k": 1001, "1M": 1002, "5M": 1003, "20M": 1004}


@dataclass
class SplatCloud:
    pos: np.ndarray        # [N,3] f32
    cd_h: np.ndarray       # [N,3] f16
    alpha: np.ndarray      # [N]   f32
    scale_h: np.ndarray    # [N,3] f16
    orient_h: np.ndarray   # [N,4] f16 (x,y,z,w)
    shx_h: np.ndarray | None = None  # [N,16] f16 (R)
    shy_h: np.ndarray | None = None  # [N,16] f16 (G)
    shz_h: np.ndarray | None = None  # [N,16] f16 (B)

    @property
    def n(self) -> int:
        return int(self.pos.shape[0])

    @property
    def has_sh(self) -> bool:
        return self.shx_h is not None

    def barycentre(self) -> np.ndarray:
        """GEO_PrimGsplat::baryCenter (GEO_GSplat.C:338-351): sequential fp32 sum / N."""
        cached = getattr(self, "_bary", None)           # (id of the position array, value): a replaced array recomputes
        if cached is not None and cached[0] == id(self.pos):
            return cached[1]
        s = np.zeros(3, dtype=np.float32)
        # sequential fp32 accumulation, done per axis with cumsum (which is sequential in numpy)
        for k in range(3):
            s[k] = np.cumsum(self.pos[:, k], dtype=np.float32)[-1] if self.n else 0.0
        b = (s / np.float32(self.n)).astype(np.float32)
        try:
            object.__setattr__(self, "_bary", (id(self.pos), b))   # the prim computes it at cook time too (GEO_GSplat.C:338)
        except Exception:
            pass
        return b

    def subset(self, idx) -> "SplatCloud":
        f = lambda a: None if a is None else np.ascontiguousarray(a[idx])
        return SplatCloud(f(self.pos), f(self.cd_h), f(self.alpha), f(self.scale_h), f(self.orient_h),
                          f(self.shx_h), f(self.shy_h), f(self.shz_h))


def _gen(seed: int, stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def make_cloud(n: int, seed: int, sh: bool = True, scale_mult: float = 1.0) -> SplatCloud:
    """SURVEY.md §8d distributions.  Each attribute has its own Philox stream (key=[seed, k])."""
    s0 = 0.5 * (8.0 / max(n, 1)) ** (1.0 / 3.0) * scale_mult
    pos = np.empty((n, 3), np.float32)
    cd = np.empty((n, 3), np.float16)
    alpha = np.empty((n,), np.float32)
    scale = np.empty((n, 3), np.float16)
    orient = np.empty((n, 4), np.float16)
    shs = [np.zeros((n, 16), np.float16) for _ in range(3)] if sh else [None] * 3
    g_pos, g_scale, g_or, g_a, g_cd, g_sh = (_gen(seed, k) for k in range(6))
    for a in range(0, n, CHUNK):
        b = min(n, a + CHUNK)
        m = b - a
        pos[a:b] = g_pos.random((m, 3), dtype=np.float32) * np.float32(2.0) - np.float32(1.0)
        scale[a:b] = (np.float32(s0) * np.exp(np.float32(0.5) * g_scale.standard_normal((m, 3), dtype=np.float32))).astype(np.float16)
        q = g_or.standard_normal((m, 4), dtype=np.float32)
        q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), np.float32(1e-12))
        orient[a:b] = q.astype(np.float16)
        alpha[a:b] = np.float32(0.05) + np.float32(0.95) * g_a.random((m,), dtype=np.float32)
        cd[a:b] = g_cd.random((m, 3), dtype=np.float32).astype(np.float16)
        if sh:
            c = (np.float32(0.1) * g_sh.standard_normal((m, 45), dtype=np.float32)).astype(np.float16)
            c = c.reshape(m, 15, 3)  # coefficient j, channel
            for ch in range(3):
                shs[ch][a:b, :15] = c[:, :, ch]
    return SplatCloud(pos, cd, alpha, scale, orient, shs[0], shs[1], shs[2])


# ------------------------------------------------------------------------------------ camera
def look_at(eye, target, up) -> np.ndarray:
    """OpenGL view matrix, column-vector convention, returned as a 4x4 (row, col) float64 array."""
    eye = np.asarray(eye, np.float64); target = np.asarray(target, np.float64); up = np.asarray(up, np.float64)
    f = target - eye; f /= np.linalg.norm(f)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3] = s; m[1, :3] = u; m[2, :3] = -f
    m[:3, 3] = -m[:3, :3] @ eye
    return m


def perspective(fovy_deg: float, aspect: float, near: float, far: float) -> np.ndarray:
    f = 1.0 / math.tan(math.radians(fovy_deg) / 2.0)
    m = np.zeros((4, 4))
    m[0, 0] = f / aspect; m[1, 1] = f
    m[2, 2] = (far + near) / (near - far); m[2, 3] = 2.0 * far * near / (near - far)
    m[3, 2] = -1.0
    return m


def colmajor(m: np.ndarray) -> np.ndarray:
    """4x4 (row, col) -> 16 floats, column-major (element (r,c) at c*4+r), f32."""
    return np.ascontiguousarray(np.asarray(m, np.float64).T.reshape(16).astype(np.float32))


@dataclass
class Frame:
    """What the HDK shim reads from RE_Render / the glH_* builtins (SRC.h:153-159) for one redraw."""
    width: int
    height: int
    view: np.ndarray          # 16 f32 column-major
    proj: np.ndarray
    object: np.ndarray = field(default_factory=lambda: colmajor(np.eye(4)))
    inv_object: np.ndarray = field(default_factory=lambda: colmajor(np.eye(4)))
    obj_view: np.ndarray | None = None
    is_object_level: bool = False

    def __post_init__(self):
        if self.obj_view is None:
            v = self.view.reshape(4, 4).T.astype(np.float64)
            o = self.object.reshape(4, 4).T.astype(np.float64)
            self.obj_view = colmajor(v @ o)


def orbit_frame(width: int, height: int, theta_deg: float = 0.0, radius: float = 3.0,
                fovy_deg: float = 50.0, near: float = 0.01, far: float = 100.0) -> Frame:
    """Benchmark camera of SURVEY.md §8d: eye on the XZ circle looking at the origin, up +Y."""
    th = math.radians(theta_deg)
    eye = (radius * math.sin(th), 0.0, radius * math.cos(th))
    view = look_at(eye, (0.0, 0.0, 0.0), (0.0, 1.0, 0.0))
    proj = perspective(fovy_deg, width / height, near, far)
    return Frame(width, height, colmajor(view), colmajor(proj))


# name -> (n, sh, width, height, seed, orbit)
WORKLOADS = {
    "10k_sh3_cook":  dict(n=10_000, sh=True, width=0, height=0, seed=1001, orbit=False),
    "1M_sh0_1080p":  dict(n=1_000_000, sh=False, width=1920, height=1080, seed=1002, orbit=False),
    "5M_sh3_1080p_orbit": dict(n=5_000_000, sh=True, width=1920, height=1080, seed=1003, orbit=True),
    "20M_sh3_1080p": dict(n=20_000_000, sh=True, width=1920, height=1080, seed=1004, orbit=False),
    "20M_sh3_4k":    dict(n=20_000_000, sh=True, width=3840, height=2160, seed=1004, orbit=False),
    "20M_sh3_8k":    dict(n=20_000_000, sh=True, width=7680, height=4320, seed=1004, orbit=False),
}


def make_inria_shell_columns(n: int, seed: int, shells: int = 3) -> dict:
    """Raw INRIA-PLY columns of a SURFACE-like capture: points on a few nested, bumpy spherical shells
    (a real capture is mostly surfaces, unlike the filled cube of §8d), flat splats lying in the tangent plane, opacities
    as logits, scales as logs, quaternions w-first and un-normalised, SH degree 3 in f_dc / f_rest.  Feed it through
    ply.write_ply / ply.load_inria / GSplatRenderer.update(..., activation=ACT_INRIA)."""
    rng = np.random.Generator(np.random.Philox(key=[seed, 77]))
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    shell = rng.integers(0, shells, n)
    radius = (0.45 + 0.25 * shell + 0.03 * np.sin(7.0 * d[:, 0]) * np.cos(5.0 * d[:, 1])).astype(np.float32)
    pos = d * radius[:, None]
    # tangent frame -> quaternion (x, y, z, w) of the rotation taking +z to the normal d
    z = np.array([0.0, 0.0, 1.0], np.float32)
    axis = np.cross(np.broadcast_to(z, d.shape), d)
    s = np.linalg.norm(axis, axis=1, keepdims=True)
    axis = np.where(s > 1e-6, axis / np.maximum(s, 1e-12), np.array([1.0, 0.0, 0.0], np.float32))
    ang = np.arccos(np.clip(d[:, 2], -1, 1))
    qxyz = axis * np.sin(ang / 2)[:, None]
    qw = np.cos(ang / 2)
    gain = (0.5 + rng.random(n)).astype(np.float32)                       # INRIA checkpoints do not store unit quaternions
    s0 = 0.5 * (4.0 * np.pi * 0.7 ** 2 * shells / max(n, 1)) ** 0.5       # surface density -> splat size
    ls = np.log(s0) + 0.4 * rng.standard_normal((n, 3))
    ls[:, 2] -= 2.0                                                       # thin along the normal
    cols = {"x": pos[:, 0], "y": pos[:, 1], "z": pos[:, 2],
            "opacity": rng.normal(1.5, 2.0, n), "scale_0": ls[:, 0], "scale_1": ls[:, 1], "scale_2": ls[:, 2],
            "rot_0": qw * gain, "rot_1": qxyz[:, 0] * gain, "rot_2": qxyz[:, 1] * gain, "rot_3": qxyz[:, 2] * gain}
    for c in range(3):
        cols[f"f_dc_{c}"] = rng.normal(0.3, 1.0, n)
    for j in range(45):
        cols[f"f_rest_{j}"] = rng.normal(0.0, 0.08, n)
    return {k: np.ascontiguousarray(v, np.float32) for k, v in cols.items()}
