"""Literal, line-by-line emulation of the reference's GLSL (float64), independent of the spec path.

TEST INFRASTRUCTURE ONLY.  Follows the shader TEXT of
  /root/reference/gsplat_plugin/shaders/GSplatShaderCoreLib.h:10-93,103-179   (library)
  /root/reference/gsplat_plugin/shaders/GSplatShaderSource.h:161-288,304-312  (main VS / FS)
and the ROP state of /root/reference/gsplat_plugin/src/GSplatRenderer.C:613-621, with GLSL's
column-major matrix constructors emulated explicitly, then rasterises the 4 quad corners the VS
emits by inverting the affine corner map per pixel (what the fixed-function interpolator does to the
``pos`` varying; w is constant over the quad so interpolation is affine).  Nothing here is derived:
it exists to check that oracle/gsplat_oracle.cpp (the closed-form spec) means the same thing.
Pure numpy / Python loops — small cases only.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- GLSL helpers
def mat3(*v):
    """GLSL mat3(a,b,c, d,e,f, g,h,i): first three scalars are COLUMN 0."""
    return np.array(v, np.float64).reshape(3, 3).T


def mat4_from_colmajor(a16):
    return np.asarray(a16, np.float64).reshape(4, 4).T


def col(M, c):          # GLSL M[c]
    return M[:, c]


def el(M, c, r):        # GLSL M[c][r]
    return M[r, c]


def normalize(v):
    return v / np.sqrt(np.dot(v, v))


def clamp(x, lo, hi):
    return min(max(x, lo), hi)


# ----------------------------------------------------------------------------- LIB.h
def CalcMatrixFromRotationScale(rot, scale):                     # LIB.h:10-27
    ms = mat3(scale[0], 0, 0, 0, scale[1], 0, 0, 0, scale[2])
    x, y, z, w = rot  # GLSL names rot.x .. rot.w
    mr = mat3(
        1.0 - 2.0 * (z * z + w * w), 2.0 * (y * z - x * w), 2.0 * (y * w + x * z),
        2.0 * (y * z + x * w), 1.0 - 2.0 * (y * y + w * w), 2.0 * (z * w - x * y),
        2.0 * (y * w - x * z), 2.0 * (z * w + x * y), 1.0 - 2.0 * (y * y + z * z))
    return ms @ mr


def CalcCovariance3D(rotMat):                                    # LIB.h:29-35
    sig = rotMat.T @ rotMat
    sigma0 = np.array([el(sig, 0, 0), el(sig, 0, 1), el(sig, 0, 2)])
    sigma1 = np.array([el(sig, 1, 1), el(sig, 1, 2), el(sig, 2, 2)])
    return sigma0, sigma1, sig


def CalcCovariance2D(worldPos, matrixV, matrixP, screenSize, sigma):   # LIB.h:38-76
    viewPos = (matrixV @ np.append(worldPos, 1.0))[:3].copy()
    aspect = el(matrixP, 0, 0) / el(matrixP, 1, 1)
    tanFovX = 1.0 / el(matrixP, 0, 0)
    tanFovY = 1.0 / (el(matrixP, 1, 1) * aspect)
    limX = 1.3 * tanFovX
    limY = 1.3 * tanFovY
    viewPos[0] = clamp(viewPos[0] / viewPos[2], -limX, limX) * viewPos[2]
    viewPos[1] = clamp(viewPos[1] / viewPos[2], -limY, limY) * viewPos[2]
    focal = screenSize[0] * el(matrixP, 0, 0) / 2
    J = mat3(
        focal / viewPos[2], 0, -(focal * viewPos[0]) / (viewPos[2] * viewPos[2]),
        0, focal / viewPos[2], -(focal * viewPos[1]) / (viewPos[2] * viewPos[2]),
        0, 0, 0)
    W = matrixV[:3, :3].copy()      # mat3(viewMatrix)
    W = W.T
    T = W @ J
    cov = T.T @ (sigma.T @ T)
    cov = cov.copy()
    cov[0, 0] += 0.3                # cov[0][0]
    cov[1, 1] += 0.3                # cov[1][1]
    return np.array([el(cov, 0, 0), el(cov, 0, 1), el(cov, 1, 1)])


def DecomposeCovariance(cov2d):                                  # LIB.h:79-93
    diag1, offDiag, diag2 = cov2d[0], cov2d[1], cov2d[2]
    mid = 0.5 * (diag1 + diag2)
    radius = np.sqrt(((diag1 - diag2) / 2.0) ** 2 + offDiag ** 2)
    lambda1 = mid + radius
    lambda2 = max(mid - radius, 0.1)
    diagVec = normalize(np.array([offDiag, lambda1 - diag1]))
    diagVec = np.array([diagVec[0], -diagVec[1]])
    maxSize = 4096.0
    v1 = min(np.sqrt(2.0 * lambda1), maxSize) * diagVec
    v2 = min(np.sqrt(2.0 * lambda2), maxSize) * np.array([diagVec[1], -diagVec[0]])
    return v1, v2


SH_C1 = 0.4886025
SH_C2 = [1.0925484, -1.0925484, 0.3153916, -1.0925484, 0.5462742]
SH_C3 = [-0.5900436, 2.8906114, -0.4570458, 0.3731763, -0.4570458, 1.4453057, -0.5900436]


def ShadeSH(color, sh, dirv, shOrder):                           # LIB.h:117-179 (sh[0] = sh1)
    x, y, z = dirv
    res = np.array(color, np.float64)
    if shOrder >= 1:
        res = res + SH_C1 * (-sh[0] * y + sh[1] * z - sh[2] * x)
        if shOrder >= 2:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = res + ((SH_C2[0] * xy) * sh[3] + (SH_C2[1] * yz) * sh[4]
                         + (SH_C2[2] * (2 * zz - xx - yy)) * sh[5] + (SH_C2[3] * xz) * sh[6]
                         + (SH_C2[4] * (xx - yy)) * sh[7])
            if shOrder >= 3:
                res = res + ((SH_C3[0] * y * (3 * xx - yy)) * sh[8] + (SH_C3[1] * xy * z) * sh[9]
                             + (SH_C3[2] * y * (4 * zz - xx - yy)) * sh[10]
                             + (SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy)) * sh[11]
                             + (SH_C3[4] * x * (4 * zz - xx - yy)) * sh[12]
                             + (SH_C3[5] * z * (xx - yy)) * sh[13]
                             + (SH_C3[6] * x * (xx - 3 * yy)) * sh[14])
    return np.maximum(res, 0.0)


# ----------------------------------------------------------------------------- SRC.h main VS
def CalculateQuadPos(vtx):                                       # SRC.h:168-188
    q = np.array([0.0, 0.0])
    if vtx == 0:
        q = np.array([1.0, 0.0])
    elif vtx == 3:
        q = np.array([0.0, 1.0])
    elif vtx in (1, 5):
        q = np.array([1.0, 1.0])
    q = (q * 2) - 1
    q = q * 2
    return q


def vertex_shader(i, vtx, cloud, frame, cam, origin, sh_order):
    """Returns None if the splat is culled by the VS, else dict(gl_Position, pos, color, opacity)."""
    ObjView = mat4_from_colmajor(frame.obj_view); Obj = mat4_from_colmajor(frame.object)
    InvObj = mat4_from_colmajor(frame.inv_object); View = mat4_from_colmajor(frame.view)
    Proj = mat4_from_colmajor(frame.proj)
    screen = np.array([float(frame.width), float(frame.height)])
    origin32 = np.asarray(origin, np.float32)
    # texel = P - origin in f32 (R.C:459-461); shader adds the origin back in f32 (SRC.h:201-202)
    P = ((cloud.pos[i].astype(np.float32) - origin32) + origin32).astype(np.float64)
    flipY = np.diag([1.0, -1.0, 1.0, 1.0])
    centerViewPos = (ObjView @ np.append(P, 1.0))[:3]
    centerClipPos = (Proj @ flipY) @ np.append(centerViewPos, 1.0)
    if centerClipPos[3] <= 0:
        return None
    color = cloud.cd_h[i].astype(np.float64)
    alpha = float(cloud.alpha[i])
    scale = cloud.scale_h[i].astype(np.float64)
    orient = cloud.orient_h[i].astype(np.float64)            # xyzw
    quadPos = CalculateQuadPos(vtx)
    rs = CalcMatrixFromRotationScale(np.array([orient[3], orient[0], orient[1], orient[2]]), scale)  # orient.wxyz
    rs = rs @ Obj[:3, :3].T
    _, _, sigma = CalcCovariance3D(rs)
    cov2d = CalcCovariance2D(P, View, Proj, screen, sigma)
    v1, v2 = DecomposeCovariance(cov2d)
    if sh_order > 0 and cloud.shx_h is not None:
        sh = [np.array([cloud.shx_h[i, j], cloud.shy_h[i, j], cloud.shz_h[i, j]], np.float64) for j in range(15)]
        worldCamToPoint = P - np.asarray(cam, np.float64)
        objCamToPoint = InvObj[:3, :3] @ worldCamToPoint
        color = ShadeSH(color, sh, normalize(objCamToPoint), sh_order)
    delta = (quadPos[0] * v1 + quadPos[1] * v2) * 2 / screen
    out = centerClipPos.copy()
    out[:2] += delta * centerClipPos[3]
    out[1] = -out[1]
    return dict(gl_Position=out, pos=quadPos, color=color, opacity=alpha)


def splat_quad(i, cloud, frame, cam, origin, sh_order):
    """Window-space corners of splat i plus its varyings, or None if culled / clipped in z."""
    corners = {}
    base = None
    for vtx in range(6):
        o = vertex_shader(i, vtx, cloud, frame, cam, origin, sh_order)
        if o is None:
            return None
        base = o
        g = o["gl_Position"]
        if not (-g[3] <= g[2] <= g[3]):       # GL clip volume, constant z over the quad
            return None
        ndc = g[:3] / g[3]
        win = np.array([(ndc[0] + 1) / 2 * frame.width, (ndc[1] + 1) / 2 * frame.height])
        corners[tuple(o["pos"])] = win
    c00 = corners[(-2.0, -2.0)]; c10 = corners[(2.0, -2.0)]; c01 = corners[(-2.0, 2.0)]
    ax = (c10 - c00) / 4.0       # d(window)/d(qx)
    ay = (c01 - c00) / 4.0       # d(window)/d(qy)
    centre = c00 + 2.0 * ax + 2.0 * ay
    g = base["gl_Position"]
    return dict(centre=centre, ax=ax, ay=ay, color=base["color"], opacity=base["opacity"], ndc_z=g[2] / g[3])


def render(cloud, frame, cam, origin, sh_order, order, edge_tol=1e-4, scene_depth=None, depth_func=0,
           depth_range=(0.0, 1.0), depth_tol=3e-7):
    """Full-frame literal render in the given submission order (front to back, R.C:613-621).
    Returns (rgba [H,W,4] f64, unsafe [H,W] bool) where unsafe marks pixels that came within
    edge_tol of a coverage / discard discontinuity for some splat (excluded from comparisons).
    scene_depth ([H,W]) + depth_func (1 = GL_LESS, 2 = GL_LEQUAL): the fixed-function depth test the reference leaves on
    (R.C:608-610) with depth writes off; a fragment's window depth is the quad's constant z/w mapped by glDepthRange.
    Pixels whose scene depth lies within depth_tol of a splat's depth are marked unsafe as well."""
    H, W = frame.height, frame.width
    dst = np.zeros((H, W, 4))
    unsafe = np.zeros((H, W), bool)
    ys, xs = np.mgrid[0:H, 0:W]
    px = xs + 0.5; py = ys + 0.5
    for i in order:
        q = splat_quad(int(i), cloud, frame, cam, origin, sh_order)
        if q is None:
            continue
        A = np.array([[q["ax"][0], q["ay"][0]], [q["ax"][1], q["ay"][1]]])
        Ai = np.linalg.inv(A)
        dx = px - q["centre"][0]; dy = py - q["centre"][1]
        qx = Ai[0, 0] * dx + Ai[0, 1] * dy
        qy = Ai[1, 0] * dx + Ai[1, 1] * dy
        inside = (np.abs(qx) <= 2.0) & (np.abs(qy) <= 2.0)
        power = -(qx * qx + qy * qy)
        alpha = np.clip(np.exp(power) * q["opacity"], 0.0, 1.0)       # FS SRC.h:306-307
        keep = inside & ~(alpha < 1.0 / 255.0)                        # discard SRC.h:308-309
        if scene_depth is not None and depth_func:
            n_, f_ = depth_range
            zw = q["ndc_z"] * (f_ - n_) / 2.0 + (f_ + n_) / 2.0
            keep &= (zw < scene_depth) if depth_func == 1 else (zw <= scene_depth)
            unsafe |= inside & (np.abs(zw - scene_depth) < depth_tol)
        near_edge = (np.abs(np.abs(qx) - 2.0) < edge_tol) | (np.abs(np.abs(qy) - 2.0) < edge_tol)
        near_disc = inside & (np.abs(alpha - 1.0 / 255.0) < edge_tol * (1.0 / 255.0) * 4)
        unsafe |= near_edge & (alpha >= 0.5 / 255.0) | near_disc
        src = np.concatenate([q["color"][None, None, :] * alpha[..., None], alpha[..., None]], axis=2)
        f = (1.0 - dst[..., 3:4])                                     # ONE_MINUS_DST_ALPHA, ONE
        dst = np.where(keep[..., None], dst + f * src, dst)
    return dst, unsafe
