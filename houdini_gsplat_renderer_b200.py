"""Import alias: the product package lives in the directory ``houdini-gsplat-renderer_b200/`` (the
name the build contract fixes); a hyphen is not importable, so this module loads that directory as
the package ``houdini_gsplat_renderer_b200``."""
import importlib.util as _u
import pathlib as _p
import sys as _s

_dir = _p.Path(__file__).resolve().parent / "houdini-gsplat-renderer_b200"
_spec = _u.spec_from_file_location(__name__, _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = _u.module_from_spec(_spec)
_s.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
