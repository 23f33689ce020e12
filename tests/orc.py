"""Loader for the CPU oracle (oracle/_build/liborc.so).  TEST INFRASTRUCTURE: only tests/, smoke() and the
bench's cpu_baseline / --impl reference legs may import this."""
import ctypes, pathlib, subprocess, functools
from orb_line_slam_b200.abi import FrontEndApi

ROOT = pathlib.Path(__file__).resolve().parents[1]


@functools.lru_cache(maxsize=1)
def oracle() -> FrontEndApi:
    so = ROOT / "oracle" / "_build" / "liborc.so"
    srcs = list((ROOT / "oracle").glob("*.cpp")) + list((ROOT / "oracle").glob("*.h*")) + list((ROOT / "include").glob("*.h"))
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(str(so))
    lib.orc_fast_atan2.restype = ctypes.c_float
    return FrontEndApi(lib, "orc_", None)
