"""The reference-shaped C++ class surfaces (orb_line_slam_b200/shim: ORB_SLAM2::ORBextractor, Lineextractor, matchNNR/match,
ComputeStereoMatches*) give the same results as the C-ABI path (and hence the oracle)."""
import pathlib, subprocess, re
import numpy as np
import pytest
import orb_line_slam_b200 as olf
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS

ROOT = pathlib.Path(__file__).resolve().parents[1]
SHIM = ROOT / "orb_line_slam_b200" / "shim"


def build_shim_test():
    exe = ROOT / "tests" / "shim" / "_test_shim"
    srcs = [ROOT / "tests" / "shim" / "test_shim.cpp", SHIM / "ORBextractor.cc", SHIM / "LineExtractor.cc", SHIM / "LineMatcher.cc"]
    cmd = ["g++", "-std=c++17", "-O2", "-o", str(exe), *map(str, srcs), "-L" + str(ROOT / "orb_line_slam_b200"), "-lolf",
           "-Wl,-rpath," + str(ROOT / "orb_line_slam_b200")]
    subprocess.run(cmd, check=True)
    return exe


def fnv(b: bytes, h=1469598103934665603):
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_shim_compiles_and_links():
    olf.load_library()
    assert build_shim_test().exists()


@pytest.mark.gpu
def test_shim_equals_abi(tmp_path):
    exe = build_shim_test()
    sc = Scene("euroc", 3)
    L, R = sc.stereo(0)
    (tmp_path / "l.raw").write_bytes(L.tobytes()); (tmp_path / "r.raw").write_bytes(R.tobytes())
    out = subprocess.run([str(exe), "640", "480", str(tmp_path / "l.raw"), str(tmp_path / "r.raw")], capture_output=True, text=True, check=True).stdout
    assert "EXCEPTION" not in out, out
    v = dict(re.findall(r"(\w+) (\d+)", out))
    fe = FrontEnd(olf.api(0), CAMERAS["euroc"], 1000, 200)
    f = fe.process(L, R)
    m, nm = fe.api.match_lines(f.ldesc, f.ldesc_r, 0.9, True)
    assert int(v["nL"]) == len(f.kps) and int(v["nR"]) == len(f.kps_r) and int(v["mL"]) == len(f.kls) and int(v["mR"]) == len(f.kls_r)
    assert int(v["desc"]) == fnv(f.desc.tobytes()) and int(v["ldesc"]) == fnv(f.ldesc.tobytes())
    assert int(v["uright"]) == fnv(f.u_right.tobytes()) and int(v["disp"]) == fnv(np.ascontiguousarray(f.line_disp).tobytes())
    assert re.search(r"match (\d+) (\d+)", out).groups() == (str(nm), str(fnv(m.tobytes())))
    assert "pyr7 179x134" in out and int(v["levels"]) == 8
    fe.close()
