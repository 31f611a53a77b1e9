"""Host-side helpers that need no GPU: NUMA placement context manager, launch-period detection of the ncu summary tool."""
import importlib.util
import os
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_near_gpu_is_a_noop_without_a_gpu_and_restores_affinity():
    from houdini_gsplat_renderer_b200 import multigpu as M
    before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    with M.near_gpu(0) as ng:
        assert ng.cpus is None or set(ng.cpus) <= set(before)
    if before is not None:
        assert os.sched_getaffinity(0) == before
    os.environ["GSB_NUMA_AFFINITY"] = "0"
    try:
        with M.near_gpu(0) as ng:
            assert ng.cpus is None and ng.saved is None
    finally:
        del os.environ["GSB_NUMA_AFFINITY"]


def test_launch_period_of_a_static_camera_capture():
    spec = importlib.util.spec_from_file_location("ncu_summary", ROOT / "tools" / "ncu_summary.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    frame = ["cell_project", "choose", "select", "pass", "pass", "pass", "records", "scan", "emit", "pass", "pass", "range", "blend",
             "sat", "select", "pass", "pass", "pass", "records", "scan", "emit", "pass", "pass", "range", "blend"]
    seq = (frame * 3)[7:7 + 40]                       # a capture that starts mid-frame
    assert mod.find_period(seq) == len(frame)
    assert mod.find_period(["a", "b", "c", "d", "e"]) is None
