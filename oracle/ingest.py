"""CPU restatement of GR_PrimGsplat::update's attribute extraction (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/gsplat_plugin/src/GR_GSplat.C: SH encoding discovery :145-189 (sh_coefficients vec3 array, else
sh1..sh15, else f_rest_0..44 — an encoding counts only when complete), Alpha-over-opacity :240-257, defaults :309-312,
half quantisation :314-318 (UT_Vector3H / UT_Vector4H = IEEE binary16, round to nearest even = numpy astype(float16)),
SH matrix layout :320-368 (coefficient j at (j/4, j%4), f_rest_j / j+15 / j+30 = R / G / B), sh-order rule :444-457,
and GEO_PrimGsplat::baryCenter (GEO_GSplat.C:338-351, sequential fp32 sum / N)."""
from __future__ import annotations

import numpy as np


SH_C0 = np.float32(0.28209479177387814)


def det_exp(x):
    """exp(x) in float64 with + - * and rint only — the operation sequence of csrc/ingest.cu det_exp(), same bits."""
    x = np.asarray(x, np.float64)
    xc = np.clip(x, -745.0, 709.0)
    k = np.rint(xc * 1.4426950408889634)
    r = (xc - k * 6.93147180369123816490e-01) - k * 1.90821492927058770002e-10
    p = np.full_like(r, 1.0 / 6227020800.0)
    for c in (1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0,
              1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0):
        p = p * r + c
    ki = k.astype(np.int64)
    k1 = np.where(ki >= 0, ki // 2, -((-ki) // 2))          # C integer division truncates toward zero
    k2 = ki - k1
    s1 = ((k1 + 1023) << 52).view(np.float64); s2 = ((k2 + 1023) << 52).view(np.float64)
    out = (p * s1) * s2
    out = np.where(x > 709.0, np.inf, np.where(x < -745.0, 0.0, out))
    return np.where(np.isnan(x), x, out)


def activate_inria(raw: dict) -> dict:
    """The example scene's wrangles (SURVEY 8a note N1) on the raw INRIA columns, fp32 results — the restatement of what
    csrc/ingest.cu does for activation = GSB_ACT_INRIA (same operations, same order)."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    out = {k: v for k, v in raw.items() if k not in ("Cd", "opacity", "scale", "orient")}
    if raw.get("Cd") is not None:
        out["Cd"] = (SH_C0 * f32(raw["Cd"]) + np.float32(0.5)).astype(np.float32)
    if raw.get("opacity") is not None:
        out["opacity"] = (1.0 / (1.0 + det_exp(-f32(raw["opacity"]).astype(np.float64)))).astype(np.float32)
    if raw.get("scale") is not None:
        out["scale"] = det_exp(f32(raw["scale"]).astype(np.float64)).astype(np.float32)
    if raw.get("orient") is not None:
        q = f32(raw["orient"])
        w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        nn = np.sqrt(((x * x + y * y) + z * z) + w * w).astype(np.float32)
        d = np.where(nn > 0, nn, np.float32(1)).astype(np.float32)
        out["orient"] = np.stack([x / d, y / d, z / d, w / d], axis=1).astype(np.float32)
    return out


def update(attrs: dict) -> dict:
    P = np.ascontiguousarray(attrs["P"], np.float32)
    n = P.shape[0]
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    cd = f32(attrs["Cd"]) if attrs.get("Cd") is not None else np.zeros((n, 3), np.float32)
    if attrs.get("Alpha") is not None:
        alpha = f32(attrs["Alpha"])
    elif attrs.get("opacity") is not None:
        alpha = f32(attrs["opacity"])
    else:
        alpha = np.ones(n, np.float32)
    scale = f32(attrs["scale"]) if attrs.get("scale") is not None else np.ones((n, 3), np.float32)
    if attrs.get("orient") is not None:
        orient = f32(attrs["orient"])
    else:
        orient = np.tile(np.array([0, 0, 0, 1], np.float32), (n, 1))
    sh = None
    if attrs.get("sh_coefficients") is not None and np.asarray(attrs["sh_coefficients"]).shape[1] > 0:
        c = f32(attrs["sh_coefficients"])[:, :15]
        sh = np.zeros((n, 15, 3), np.float32); sh[:, :c.shape[1]] = c
    elif all(attrs.get(f"sh{j + 1}") is not None for j in range(15)):
        sh = np.stack([f32(attrs[f"sh{j + 1}"]) for j in range(15)], axis=1)
    elif all(attrs.get(f"f_rest_{j}") is not None for j in range(45)):
        r = np.stack([f32(attrs[f"f_rest_{j}"]) for j in range(45)], axis=1)      # [n,45]
        sh = np.stack([r[:, 0:15], r[:, 15:30], r[:, 30:45]], axis=2)             # [n,15,3]
    out = dict(pos=P, cd_h=cd.astype(np.float16), alpha=alpha, scale_h=scale.astype(np.float16),
               orient_h=orient.astype(np.float16), shx_h=None, shy_h=None, shz_h=None)
    if sh is not None and n > 0:
        for ch, name in enumerate(("shx_h", "shy_h", "shz_h")):
            m = np.zeros((n, 16), np.float16); m[:, :15] = sh[:, :, ch].astype(np.float16)
            out[name] = m
    order, invalid = 3, False
    if "gsplat__sh_order" in attrs:
        order = int(attrs["gsplat__sh_order"])
        if order < 0 or order > 3:
            order, invalid = 0, True
    out["sh_order"], out["sh_order_invalid"] = order, invalid
    bary = np.zeros(3, np.float32)
    if n:
        for k in range(3):
            bary[k] = np.cumsum(P[:, k], dtype=np.float32)[-1]
        bary = (bary / np.float32(n)).astype(np.float32)
    out["barycentre"] = bary
    return out
