"""The HDK side of the drop-in (hdk/GSplatRenderer_b200.C) cannot be built here (no Houdini, no OpenGL); it is syntax-checked
against hdk/stubs/ — stand-in headers declaring exactly the HDK / GL members it uses — together with the REFERENCE'S OWN
include/GSplatRenderer.h and GSplatLogger.h, so every member it defines matches the class it replaces and every helper it
calls exists.  INTEGRATION.md must describe the same binding."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_INC = Path("/root/reference/gsplat_plugin/include")
SHIM = ROOT / "hdk" / "GSplatRenderer_b200.C"


@pytest.mark.skipif(not (REF_INC / "GSplatRenderer.h").exists(), reason="/root/reference not available")
def test_shim_is_syntactically_valid_against_the_reference_header():
    p = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", f"-I{ROOT / 'hdk' / 'stubs'}", f"-I{REF_INC}",
                        f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include", str(SHIM)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_shim_defines_every_public_member_and_every_helper_it_calls():
    txt = SHIM.read_text()
    for m in ("registerUpdate", "includeInRenderPass", "flushEntriesForMatchingDetail", "generateRenderGeometry", "render",
              "postRender", "setRenderingEnabled", "setExplicitCameraPos", "setSphericalHarmonicsOrder"):
        assert re.search(rf"GSplatRenderer::{m}\s*\(", txt), m
    for helper in ("copyBoundDepthAttachmentTo", "drawFullViewportTexture", "builtinMatrix", "toColumnMajorF", "ensureTexture"):
        assert re.search(rf"^void {helper}\s*\(", txt, re.M), f"{helper} is called but not defined"
    # every C-ABI call it makes is declared by the header
    hdr = (ROOT / "include" / "gsplat_b200.h").read_text()
    for call in set(re.findall(r"\b(gsb_[a-z_]+)\s*\(", txt)):
        assert re.search(rf"\b{call}\s*\(", hdr), call


def test_integration_doc_and_shim_fetch_the_matrices_the_same_way():
    doc = (ROOT / "INTEGRATION.md").read_text()
    txt = SHIM.read_text()
    for token in ("RE_UNIFORM_PROJECT_MATRIX", "RE_UNIFORM_OBJECT_MATRIX", "RE_UNIFORM_INV_OBJECT_MATRIX", "RE_UNIFORM_OBJVIEW_MATRIX",
                  "getMatrix(view)", "copyBoundDepthAttachmentTo", "drawFullViewportTexture"):
        assert token in txt and token in doc, token
    assert "getProjectionMatrix" not in txt and "getProjectionMatrix" not in doc
