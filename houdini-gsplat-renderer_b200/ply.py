"""INRIA 3D-Gaussian-splatting ``.ply`` -> the raw attribute arrays of ``gsb_update_from_attributes``.

The reference has no PLY code: its example scene (/root/reference/hip/GSplatPlugin_simpleScene_v001.hip) loads the
file with a File SOP and converts the INRIA columns with point wrangles, then casts to fp16 (SURVEY.md §2 row 12,
§8a note N1):
    Cd      = 0.28209479177387814 * f_dc + 0.5          (SH band 0 folded into the colour)
    opacity = 1 / (1 + exp(-opacity_raw))
    scale   = exp(scale_raw)
    orient  = normalize(rot_1, rot_2, rot_3, rot_0)     (Houdini quaternions are (x, y, z, w); INRIA stores w first)
    f_rest_j pass through (the renderer's f_rest encoding, GR_GSplat.C:173-184,357-366)
That conversion runs ON THE GPU, inside the ingestion kernel (csrc/ingest.cu, ``activation = GSB_ACT_INRIA``), together
with the fp16 cast; this module only reads the file and lines the raw columns up.  Host-side I/O, not hot path.
"""
from __future__ import annotations

import numpy as np

SH_C0 = np.float32(0.28209479177387814)

_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
              "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def read_ply(path) -> dict:
    """Vertex columns of a binary_little_endian (or ascii) PLY as {name: 1-D array}."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    count = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties on vertices are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "binary_little_endian":
            data = np.fromfile(f, dtype=np.dtype(props), count=count)
        elif fmt == "ascii":
            flat = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2)
            data = np.zeros(count, np.dtype(props))
            for k, (name, _) in enumerate(props):
                data[name] = flat[:, k]
        else:
            raise ValueError(f"{path}: unsupported PLY format {fmt!r}")
        if data.shape[0] != count:
            raise ValueError(f"{path}: expected {count} vertices, file holds {data.shape[0]}")
    return {name: np.ascontiguousarray(data[name]) for name, _ in props}


def write_ply(path, cols: dict):
    """Writes columns as float32 binary_little_endian vertices (tests, tools)."""
    names = list(cols)
    n = len(next(iter(cols.values())))
    rec = np.zeros(n, np.dtype([(k, "<f4") for k in names]))
    for k in names:
        rec[k] = np.asarray(cols[k], np.float32)
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n).encode())
        for k in names:
            f.write(f"property float {k}\n".encode())
        f.write(b"end_header\n")
        rec.tofile(f)


def inria_raw_attributes(cols: dict) -> dict:
    """INRIA columns -> the RAW attribute dict for ``GSplatRenderer.update(..., activation=ACT_INRIA)``: nothing is
    converted here (Cd = f_dc, opacity = logit, scale = log scale, orient = (rot_0, rot_1, rot_2, rot_3))."""
    f32 = lambda k: np.ascontiguousarray(cols[k], np.float32)
    attrs = {"P": np.stack([f32("x"), f32("y"), f32("z")], axis=1),
             "Cd": np.stack([f32("f_dc_0"), f32("f_dc_1"), f32("f_dc_2")], axis=1),
             "opacity": f32("opacity"),
             "scale": np.stack([f32("scale_0"), f32("scale_1"), f32("scale_2")], axis=1),
             "orient": np.stack([f32("rot_0"), f32("rot_1"), f32("rot_2"), f32("rot_3")], axis=1)}
    if sum(1 for k in cols if k.startswith("f_rest_")) >= 45:
        for j in range(45):
            attrs[f"f_rest_{j}"] = f32(f"f_rest_{j}")
    return attrs


def load_inria(path) -> dict:
    """Raw attributes of an INRIA .ply; pass them to ``GSplatRenderer.update(..., activation=ACT_INRIA)``."""
    return inria_raw_attributes(read_ply(path))
