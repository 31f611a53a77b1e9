"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/gsplat_b200.h declares; the ctypes struct mirrors match the header's layout.  No compute."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    txt = (ROOT / "include" / "gsplat_b200.h").read_text()
    return sorted(set(re.findall(r"GSB_API[^;(]*?\b(gsb_\w+)\s*\(", txt)))


def test_header_declares_the_reference_surface():
    syms = _header_symbols()
    # one entry point per public member of the reference's GSplatRenderer (R.h:34-56)
    for s in ["gsb_register_update", "gsb_include_in_render_pass", "gsb_flush_entries_for_matching_detail",
              "gsb_generate_render_geometry", "gsb_render", "gsb_post_render", "gsb_set_rendering_enabled",
              "gsb_set_explicit_camera_pos", "gsb_set_spherical_harmonics_order"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from houdini_gsplat_renderer_b200 import renderer as R
    lib = R.load_library()
    syms = _header_symbols()
    assert sorted(R.EXPORTS) == syms
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.gsb_abi_version() == 3
    assert lib.gsb_last_error() is not None


def test_struct_layouts_match_header():
    from houdini_gsplat_renderer_b200 import renderer as R
    assert C.sizeof(R.PrimKey) == 48
    assert C.sizeof(R.FrameC) == 5 * 64 + 6 * 4 + 4 + 4 + 8 + 8      # matrices, ints, depth_func, gl tex, range, pointer
    assert C.sizeof(R.TargetC) == 32
    assert R.RECORD_DTYPE.itemsize == 48 and R.RECT_DTYPE.itemsize == 8
    assert C.sizeof(R.StatsC) == 4 * 8 + 10 * 4 + 6 * 4 + 8 * 4 + 8


def test_no_cpu_fallback_without_device():
    """Without a CUDA device gsb_create must fail loudly (status + message), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from houdini_gsplat_renderer_b200 import renderer as R
    with pytest.raises(R.GsbError):
        R.GSplatRenderer(0)


def test_product_does_not_touch_oracle():
    pkg = ROOT / "houdini-gsplat-renderer_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("Makefile")):
        txt = p.read_text()
        code = "\n".join(l for l in txt.splitlines() if not l.lstrip().startswith(("//", "#", "*", "/*")))
        assert "import oracle" not in code and "from oracle" not in code and "gsplat_oracle" not in code, p


def test_python_constants_match_the_header_enums():
    """renderer.py mirrors the option / debug-buffer / depth-function enums of include/gsplat_b200.h by value."""
    import re
    from houdini_gsplat_renderer_b200 import renderer as R
    txt = (ROOT / "include" / "gsplat_b200.h").read_text()
    enums = {m.group(1): int(m.group(2))                    # enumerator lines only (comments quote some of them too)
             for m in re.finditer(r"^\s+(GSB_[A-Z0-9_]+)\s*=\s*(\d+)\s*,?\s*(?:/\*.*)?$", txt, re.M)}
    for name, val in enums.items():
        for prefix in ("GSB_OPT_", "GSB_DBG_", "GSB_DEPTH_"):
            if name.startswith(prefix):
                py = name[4:]                                  # GSB_OPT_EPS_T -> OPT_EPS_T
                assert hasattr(R, py), f"renderer.py lacks {py}"
                assert getattr(R, py) == val, (name, val, getattr(R, py))
    assert enums["GSB_OPT_LAZY_PROJECT"] == 9 and enums["GSB_DBG_TRECTS"] == 9 and enums["GSB_DEPTH_LEQUAL"] == 2
    assert int(re.search(r"#define GSB_ABI_VERSION (\d+)", txt).group(1)) == 3
