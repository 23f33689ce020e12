import os, sys, subprocess, pathlib
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: exhaustive sweeps, excluded from the default CPU run")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (they are the parity tests proper and need the B200 box)."""
    try:
        import orb_line_slam_b200 as olf
        ndev = olf.device_count()
    except Exception:
        ndev = 0
    if ndev > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked parity tests run on the B200 box (gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
