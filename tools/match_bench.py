"""BASELINE.json configs[4] -- descriptor-match micro-benchmark: n x n 256-bit Hamming brute-force kNN(2) (ORB / binary LBD),
n = 512..8192.  Two numbers per n: through the C-ABI with host buffers (what matchNNR costs a caller) and kernel-only on
device-resident descriptors, the latter against the chip's MEASURED xor+popc+add ceiling (the kernel moves only (n1+n2)*32 B:
its roofline is the integer pipe, not HBM -- SURVEY 8d).  One JSON line per n; the last line is the 8192^2 headline.
The float-LBD L2 variant of configs[4] has no counterpart on the reference's hot path (it matches the BINARY descriptor,
src/LineMatcher.cpp:42-62) and is not built."""
import ctypes as C, json, time
import numpy as np


def run_c5(device=0, sizes=(512, 1024, 2048, 4096, 8192)):
    import orb_line_slam_b200 as olf
    lib = olf.load_library(); api = olf.api(device)
    rng = np.random.RandomState(0)
    for n in sizes:
        d1 = rng.randint(0, 256, (n, 32)).astype(np.uint8); d2 = rng.randint(0, 256, (n, 32)).astype(np.uint8)
        for _ in range(3):
            api.knn2_hamming(d1, d2)
        reps = max(3, min(50, (1 << 26) // (n * n)))
        t0 = time.perf_counter()
        for _ in range(reps):
            api.knn2_hamming(d1, d2)
        dt = (time.perf_counter() - t0) / reps
        kms, peak = C.c_double(0), C.c_double(0)
        rc = lib.olf_knn2_bench(C.c_int(n), C.c_int(n), C.c_int(20), C.c_int(device), C.byref(kms), C.byref(peak))
        assert rc == 0
        pairs_k = n * n / (kms.value * 1e-3)
        line = {"metric": "256-bit Hamming pair-distances/s (brute-force kNN2)", "value": pairs_k, "unit": "pairs/s", "n_gpus": 1,
                "higher_is_better": True, "dtype": "u32 popc", "data": "synthetic",
                "config": {"workload": f"{n}x{n} 256-bit Hamming brute-force kNN(2), BASELINE configs[4]", "n1": n, "n2": n},
                "e2e": {"value": n * n / dt, "unit": "pairs/s", "h2d_bytes_per_step": 2 * n * 32, "d2h_bytes_per_step": 4 * n * 4, "ms_per_call": dt * 1e3},
                "kernel_ms": kms.value,
                "roofline": {"bound": "int-pipe (xor+popc+add per 32-bit word pair)", "achieved": pairs_k * 8, "peak": peak.value, "unit": "word-pairs/s",
                             "frac": pairs_k * 8 / peak.value, "hbm_bytes": (2 * n) * 32 + 4 * n * 4,
                             "note": "peak measured live by k_popc_peak (independent xor+popc+add chains on every SM)"}}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    run_c5()
