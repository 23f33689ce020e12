"""orb_line_slam_b200 -- B200-native (sm_100a) stereo point+line front end behind the ORB_Line_SLAM class surfaces.

The product is the CUDA library `libolf.so` (C-ABI in include/olf_abi.h) plus the C++ shim in shim/.  This Python
package is the thin host mirror used by the tests and the benchmark; it never falls back to a CPU path:
loading fails loudly when the library has not been built, and every call fails when no CUDA device is present.
"""
from __future__ import annotations
import ctypes, functools, os, pathlib
from .abi import FrontEndApi, LineParams, LineMatchParams, Camera, SbpLastArgs, SbpMapArgs, KEYPOINT, KEYLINE  # noqa: F401

_PKG = pathlib.Path(__file__).resolve().parent
# OLF_LIB selects another build of the same library (e.g. the -DOLF_LSD_PROFILE build used by tools/lsd_trace.py)
LIB_PATH = pathlib.Path(os.environ["OLF_LIB"]).resolve() if os.environ.get("OLF_LIB") else _PKG / "libolf.so"


@functools.lru_cache(maxsize=1)
def load_library() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    return ctypes.CDLL(str(LIB_PATH))


def api(device: int = 0) -> FrontEndApi:
    """Bind the olf_* entry points for one CUDA device."""
    lib = load_library()
    lib.olf_device_count.restype = ctypes.c_int
    return FrontEndApi(lib, "olf_", device)


def device_count() -> int:
    lib = load_library()
    lib.olf_device_count.restype = ctypes.c_int
    return lib.olf_device_count()
