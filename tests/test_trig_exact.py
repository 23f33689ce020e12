"""csrc/sincosf_exact.h (compiled for the host) == this machine's libm sinf / cosf on EVERY float in [0, 6.5]: the device
evaluates rBRIEF's per-keypoint trig with this source."""
import pathlib, subprocess, tempfile

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_sincosf_exact_equals_libm_exhaustively():
    with tempfile.TemporaryDirectory() as d:
        exe = pathlib.Path(d) / "trig_check"
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-pthread", "-o", str(exe), str(ROOT / "tests/emul/trig_check.cpp"), "-lm"], check=True)
        r = subprocess.run([str(exe), "1"], capture_output=True, text=True)
        total, bad = map(int, r.stdout.split())
        assert r.returncode == 0 and bad == 0 and total == 1087373313, r.stdout
