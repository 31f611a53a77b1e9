#!/usr/bin/env python
"""build_ref.py — builds oracle/_ref/libgsplat_ref.so from the reference's own shader sources.  TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py [--reference /root/reference] [--force]

1. compiles oracle/ref_dump.cpp with -I<reference>/gsplat_plugin/shaders: the reference's GSplatShaderSource.h /
   GSplatShaderCoreLib.h are included WHERE THEY LIE, and the helper writes their shader strings, byte for byte, into a
   scratch directory (tempfile, deleted afterwards);
2. turns each string into a C++ include file by the token-level rewrites of REWRITES below — the only places where GLSL is
   not already valid C++ once oracle/glsl_cxx.h supplies the types (`out T x` parameter -> `T& x`, interface blocks and
   shader outputs -> thread-local structs / variables, vertex attributes -> thread-local variables); every other byte
   is the reference's;
3. compiles oracle/ref_harness.cpp (which #includes them) into oracle/_ref/libgsplat_ref.so with -ffp-contract=off;
4. writes oracle/_ref/manifest.json: sha256 of every extracted string, how often each rewrite fired, compiler flags.

Only the .so and the manifest are kept (oracle/_ref/ is git-ignored; it travels to the GPU box with gpurun).  No reference
source text is stored in the repository.  Without /root/reference the script does nothing (the prebuilt .so is used).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import re
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT_DIR = HERE / "_ref"
OUT = OUT_DIR / "libgsplat_ref.so"

# (name, regex, replacement): applied to every extracted string, in this order
REWRITES = [
    # `out vec3 sigma0` in a parameter list -> C++ reference parameter
    ("out_parameter", re.compile(r"\bout\s+(vec2|vec3|vec4|mat3|mat4|float)\s+(\w+)\s*(?=[,)])"), r"\1& \2"),
    # `out parms { ... } vsOut;` / `in parms { ... } fsIn;` -> a thread-local struct variable
    ("interface_block", re.compile(r"\b(?:in|out)\s+parms\b"), "thread_local struct parms_t"),
    # fragment shader output
    ("fragment_output", re.compile(r"\bout\s+vec4\s+color_out\s*;"), "thread_local vec4 color_out;"),
    # vertex attributes of the wireframe shader: `in vec3 P;`
    ("vertex_attribute", re.compile(r"^(\s*)in\s+(vec[234])\s+(\w+)\s*;", re.M), r"\1thread_local \2 \3;"),
]
PARTS = ["core_lib", "sh_lib", "main_vs", "main_fs", "wire_vs", "wire_fs"]
CXX = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-ftls-model=initial-exec",
            "-Wno-unused-variable", "-Wno-unused-but-set-variable"]


def build(reference: Path = Path("/root/reference"), force: bool = False) -> Path | None:
    shaders = reference / "gsplat_plugin" / "shaders"
    if not (shaders / "GSplatShaderSource.h").exists():
        return OUT if OUT.exists() else None                 # GPU box: the prebuilt library travels with the snapshot
    srcs = [HERE / "ref_harness.cpp", HERE / "glsl_cxx.h", HERE / "ref_dump.cpp", Path(__file__),
            shaders / "GSplatShaderSource.h", shaders / "GSplatShaderCoreLib.h"]
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return OUT
    OUT_DIR.mkdir(exist_ok=True)
    tmp = Path(tempfile.mkdtemp(prefix="gsb_ref_"))
    try:
        subprocess.run([CXX, "-O0", "-std=c++17", f"-I{shaders}", str(HERE / "ref_dump.cpp"), "-o", str(tmp / "ref_dump")], check=True)
        subprocess.run([str(tmp / "ref_dump"), str(tmp)], check=True)
        manifest = {"reference": str(reference), "strings": {}, "rewrites": {}, "cxxflags": CXXFLAGS}
        for part in PARTS:
            text = (tmp / f"{part}.glsl").read_text()
            manifest["strings"][part] = {"sha256": hashlib.sha256(text.encode()).hexdigest(), "bytes": len(text)}
            fired = {}
            for name, rx, rep in REWRITES:
                text, k = rx.subn(rep, text)
                fired[name] = k
            manifest["rewrites"][part] = fired
            (tmp / f"{part}.inc").write_text(text)
        subprocess.run([CXX, *CXXFLAGS, f"-I{tmp}", f"-I{HERE}", str(HERE / "ref_harness.cpp"), "-o", str(OUT)], check=True)
        (OUT_DIR / "manifest.json").write_text(json.dumps(manifest, indent=1))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    p = build(Path(a.reference), a.force)
    print(p if p else "reference not available and no prebuilt oracle/_ref/libgsplat_ref.so", file=sys.stderr if not p else sys.stdout)
    sys.exit(0 if p else 1)
