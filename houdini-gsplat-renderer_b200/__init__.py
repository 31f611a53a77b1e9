"""B200-native drop-in for the splat -> framebuffer hot path of rubendhz/houdini-gsplat-renderer.

The product is ``csrc/`` (hand-written sm_100a CUDA + the C ABI of include/gsplat_b200.h, built
in-tree as ``libgsplat_b200.so``).  ``renderer`` is a thin ctypes mirror of the reference's
``GSplatRenderer`` call surface used by tests and bench.py; ``scene`` generates the synthetic
clouds of SURVEY.md §8d.  There is no CPU fallback: without the CUDA library the renderer raises.
"""
from . import scene  # noqa: F401

__all__ = ["scene", "renderer"]


def __getattr__(name):
    if name == "renderer":
        import importlib
        return importlib.import_module(__name__ + ".renderer")
    raise AttributeError(name)
