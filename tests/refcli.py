"""Driver for oracle/_ref/refcli: the reference's own hot-path sources compiled against stand-in OpenCV types
(oracle/Makefile target `ref`, oracle/ref_harness/).  TEST INFRASTRUCTURE: used only to pin the oracle to the reference."""
import pathlib, struct, subprocess, tempfile
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
EXE = ROOT / "oracle" / "_ref" / "refcli"
REFERENCE = pathlib.Path("/root/reference")
_DT = {np.dtype(np.uint8): 0, np.dtype(np.int32): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}
_RT = {0: np.uint8, 1: np.int32, 2: np.float32, 3: np.float64}


def available() -> bool:
    """Build refcli when the reference tree is present (this container); the GPU box has neither and does not need it."""
    if REFERENCE.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], check=True, capture_output=True)
    return EXE.exists() and REFERENCE.exists()


def run(cmd: str, *arrays):
    return run_exe(EXE, cmd, *arrays)


def run_exe(exe, cmd: str, *arrays):
    """The array-file protocol with any executable that speaks it (oracle/_ref/refcli, tests/shim/_test_shim_kf)."""
    with tempfile.TemporaryDirectory() as d:
        fin, fout = pathlib.Path(d) / "in.bin", pathlib.Path(d) / "out.bin"
        with open(fin, "wb") as f:
            f.write(struct.pack("<i", len(arrays)))
            for a in arrays:
                a = np.ascontiguousarray(a)
                f.write(struct.pack("<ii", _DT[a.dtype], a.ndim)); f.write(struct.pack(f"<{a.ndim}q", *a.shape)); f.write(a.tobytes())
        r = subprocess.run([str(exe), cmd, str(fin), str(fout)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{pathlib.Path(exe).name} {cmd} failed ({r.returncode}): {r.stderr[-2000:]}")
        raw = fout.read_bytes()
    n, = struct.unpack_from("<i", raw, 0); off = 4; out = []
    for _ in range(n):
        dt, nd = struct.unpack_from("<ii", raw, off); off += 8
        dims = struct.unpack_from(f"<{nd}q", raw, off); off += 8 * nd
        cnt = int(np.prod(dims)) if nd else 1
        a = np.frombuffer(raw, dtype=_RT[dt], count=cnt, offset=off).reshape(dims).copy(); off += a.nbytes
        out.append(a)
    return out


def keypoints_as_rows(kps):
    """olf_keypoint structured array -> [n,6] float32 rows (octave bit-cast)."""
    return np.ascontiguousarray(kps).view(np.float32).reshape(-1, 6)


def keylines_as_rows(kls):
    return np.ascontiguousarray(kls).view(np.float32).reshape(-1, 17)
