#!/usr/bin/env python3
"""bench.py -- stereo frames/sec of the ORB+line front end (extract + match) on 1..8 B200.

A "step" = one 1280x720 stereo frame through the whole hot path:
  ExtractORB(L), ExtractORB(R), ExtractLine(L), ExtractLine(R), ComputeStereoMatches, ComputeStereoMatches_Lines,
  ORBmatcher::SearchByProjection(cur,last) and LineMatcher::match(last,cur)          [configs[1] of BASELINE.json]
`value`  : frames/s with the images already resident in HBM (olf_frontend_process, on_device=1);
`e2e`    : the same through the host-buffer entry points (host->device copies inside the timed region);
results always come back to host memory (that is the API: Tracking.cc consumes host vectors).
`--impl reference` times the CPU restatement of the reference path (oracle/) with every host thread it can use (the
reference's 4 extraction threads per frame, cores//4 frames in flight) on a bounded sample of the same workload; the
`cpu_baseline` object of the GPU arm is the north star's single-thread number ("single-thread CPU ExtractORB+ExtractLine+
SearchByProjection").
"""
from __future__ import annotations
import argparse, json, os, pathlib, subprocess, sys, threading, time

# more hardware work queues than the default 8: every rig drives 4 extractor streams (must be set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np

WORKLOAD = dict(workload="1280x720 stereo synthetic seq, 2000 ORB + 500 LBD per eye, extract + stereo match + frame-to-frame match",
                camera="zed720", nfeatures=2000, nlines=500, min_line_length=0.025)
METRIC = "stereo frames/sec (extract+match, 1280x720)"
N_BASE_FRAMES = 8            # distinct rendered poses
N_DISTINCT = 80              # distinct stereo pairs (base frames + per-frame noise): 147 MB of input > 126 MB L2


def make_sequence(n_distinct: int):
    """Seeded synthetic stereo sequence.  Rendering is the slow part, so N_BASE_FRAMES poses are rendered
    and every further pair is a base pair with fresh N(0,2) sensor noise -> all pairs are distinct images."""
    from orb_line_slam_b200.synth import Scene, pose_f32
    sc = Scene(WORKLOAD["camera"], 0)
    base = [sc.stereo(f) for f in range(N_BASE_FRAMES)]
    seq, poses = [], []
    for i in range(n_distinct):
        L, R = base[i % N_BASE_FRAMES]
        if i >= N_BASE_FRAMES:
            rng = np.random.RandomState(7000 + i)
            L = np.clip(L.astype(np.int16) + np.rint(rng.normal(0, 2, L.shape)).astype(np.int16), 0, 255).astype(np.uint8)
            R = np.clip(R.astype(np.int16) + np.rint(rng.normal(0, 2, R.shape)).astype(np.int16), 0, 255).astype(np.uint8)
        seq.append((np.ascontiguousarray(L), np.ascontiguousarray(R)))
        poses.append(pose_f32(i % N_BASE_FRAMES))
    return sc, seq, poses


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def run_steps(pipes, frontends, track_fe, seq_imgs, poses, first, count, on_device, dev_imgs, n_trackers=1, do_track=True):
    """Process frames [first, first+count) on len(pipes) concurrent rigs (Frame::Frame each); one tracker thread consumes
    the finished frames in order and runs the frame-to-frame matchers of TrackWithMotionModelWithLine on (f, f-1)."""
    P = len(pipes)
    done = [threading.Event() for _ in range(count)]
    views = [None] * count
    stats = {"matches": 0, "line_matches": 0, "kps": 0, "lines": 0, "d2h": 0}
    errors = []

    def worker(p):
        try:
            nat = pipes[p]
            B = nat.max_frames
            nb = (count + B - 1) // B                       # batches of B consecutive frames, dealt round-robin to the rigs
            for b in range(p, nb, P):
                ks = list(range(b * B, min(count, (b + 1) * B)))
                idx = [(first + k) % len(seq_imgs) for k in ks]
                blks = [nat.new_block() for _ in ks]
                if on_device:
                    nat.process_batch([dev_imgs[i][0] for i in idx], [dev_imgs[i][1] for i in idx], blks, on_device=True)
                else:
                    nat.process_batch([seq_imgs[i][0] for i in idx], [seq_imgs[i][1] for i in idx], blks, on_device=False)
                for k, i, blk in zip(ks, idx, blks):
                    views[k] = nat.view(blk, poses[i])
                    done[k].set()
        except Exception as e:       # noqa: BLE001
            errors.append(e)
            for d in done:
                d.set()

    lock = threading.Lock()

    def tracker(tid):
        # frame-to-frame matchers of TrackWithMotionModelWithLine on (f, f-1): src/Tracking.cc:1296 (SearchByProjection) and
        # :1308 (line match); the pairs are independent, tracker thread `tid` takes frames tid, tid+T, ...
        try:
            loc = {"matches": 0, "line_matches": 0, "kps": 0, "lines": 0, "d2h": 0}
            for k in range(tid, count, n_trackers):
                done[k].wait()
                if k > 0:
                    done[k - 1].wait()
                if errors:
                    return
                v = views[k]
                if k > 0 and do_track:
                    t = track_fe[tid % len(track_fe)].track(v, views[k - 1])
                    loc["matches"] += t["nmatches"]; loc["line_matches"] += t.get("n_line_matches", 0)
                loc["kps"] += len(v.kps) + len(v.kps_r); loc["lines"] += len(v.kls) + len(v.kls_r)
                loc["d2h"] += (len(v.kps) + len(v.kps_r)) * 56 + len(v.kps) * 8 + (len(v.kls) + len(v.kls_r)) * 100 + len(v.kls) * 36
            with lock:
                for key, val in loc.items():
                    stats[key] += val
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    ths = [threading.Thread(target=worker, args=(p,)) for p in range(P)] + [threading.Thread(target=tracker, args=(i,)) for i in range(n_trackers)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errors:
        raise errors[0]
    return stats


def cpu_reference(frames, seq, poses, warm=1):
    """Single-thread CPU restatement of the reference path (oracle) on `frames` frames; returns (fps, seconds)."""
    from orc import oracle
    from orb_line_slam_b200.frame import FrontEnd
    from orb_line_slam_b200.synth import CAMERAS
    fe = FrontEnd(oracle(), CAMERAS[WORKLOAD["camera"]], WORKLOAD["nfeatures"], WORKLOAD["nlines"], WORKLOAD["min_line_length"])
    last = None
    t0 = None
    for k in range(warm + frames):
        if k == warm:
            t0 = time.perf_counter()
        L, R = seq[k % len(seq)]
        cur = fe.process(L, R, poses[k % len(seq)])
        if last is not None:
            fe.track(cur, last)
        last = cur
    dt = time.perf_counter() - t0
    fe.close()
    return frames / dt, dt


def cpu_reference_all_cores(frames, seq, poses, warm=1, cores=None):
    """The reference's own threading on every host core: each rig runs Frame::Frame's four extraction threads (ORB L/R,
    lines L/R, src/Frame.cc:164-171) followed by the stereo matchers; cores//4 rigs work on independent frames at the
    same time (the same frames-in-flight arrangement as the GPU arm) and one tracker consumes them in order.
    The oracle is a ctypes library, so the GIL is released inside every call.  Returns (fps, seconds, threads)."""
    from concurrent.futures import ThreadPoolExecutor
    from orc import oracle
    from orb_line_slam_b200.frame import FrontEnd, StereoFrame
    from orb_line_slam_b200.synth import CAMERAS
    cores = cores or os.cpu_count() or 4
    rigs = max(1, cores // 4)
    a = oracle()
    fes = [FrontEnd(a, CAMERAS[WORKLOAD["camera"]], WORKLOAD["nfeatures"], WORKLOAD["nlines"], WORKLOAD["min_line_length"]) for _ in range(rigs)]
    pools = [ThreadPoolExecutor(4) for _ in range(rigs)]

    def frame(r, k):
        fe, L, R = fes[r], *seq[k % len(seq)]
        jobs = [pools[r].submit(a.orb_extract, fe.orb_l, L), pools[r].submit(a.orb_extract, fe.orb_r, R),
                pools[r].submit(a.line_extract, fe.line_l, L), pools[r].submit(a.line_extract, fe.line_r, R)]
        (kl, dl), (kr, dr), (kll, dll), (klr, dlr) = [j.result() for j in jobs]
        f = StereoFrame(kl, dl, kr, dr, None, None, kll, dll, klr, dlr)
        f.u_right, f.depth = a.stereo_points(fe.orb_l, fe.orb_r, kl, dl, kr, dr, fe.bf, fe.fx)
        f.line_matches, f.line_disp, f.line_le = a.stereo_lines(kll, dll, klr, dlr, fe.w, fe.h, fe.lmp)
        f.Rcw, f.tcw = poses[k % len(seq)]
        return f

    def run(first, count):
        done = [threading.Event() for _ in range(count)]
        out = [None] * count

        def worker(r):
            for k in range(r, count, rigs):
                out[k] = frame(r, first + k); done[k].set()

        def tracker():
            prev = None
            for k in range(count):
                done[k].wait()
                if prev is not None:
                    fes[0].track(out[k], prev)
                prev = out[k]
                if k > 0:
                    out[k - 1] = None
        ths = [threading.Thread(target=worker, args=(r,)) for r in range(rigs)] + [threading.Thread(target=tracker)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0
    if warm:
        run(0, max(warm, rigs))
    dt = run(warm, frames)
    for fe in fes:
        fe.close()
    for p in pools:
        p.shutdown()
    return frames / dt, dt, rigs * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1664)
    ap.add_argument("--warmup", type=int, default=52)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pipelines", type=int, default=0, help="concurrent stereo rigs per GPU (0 = auto)")
    ap.add_argument("--trackers", type=int, default=1, help="host threads running the frame-to-frame matchers")
    ap.add_argument("--no-track", action="store_true", help="(diagnostic) skip the frame-to-frame matchers")
    ap.add_argument("--batch", type=int, default=4, help="independent stereo frames per olf_frontend_process_batch call (1..4)")
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prewarm-s", type=float, default=6.0, help="seconds of untimed work before the warm-up steps")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        if rank != 0:
            return
        sc, seq, poses = make_sequence(N_BASE_FRAMES)
        cores = os.cpu_count() or 4
        rigs = max(1, cores // 4)
        # one "step" = one stereo frame; the sample is bounded so that K steps + W warm-up end within a few minutes
        frames = max(2 * rigs, min(args.steps, args.cpu_frames * rigs))
        warm = max(rigs, min(args.warmup, 2 * rigs))
        fps, dt, threads = cpu_reference_all_cores(frames, seq, poses, warm=warm, cores=cores)
        fps1, dt1 = cpu_reference(min(frames, 8), seq, poses, warm=1)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": frames, "warmup": warm,
                "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": dict(WORKLOAD, host_cores=cores, rigs=rigs, threads=threads,
                               note="CPU restatement of the reference path (oracle/; the reference itself needs OpenCV C++/Eigen/Pangolin and cannot be "
                                    "built here), run the reference's way: 4 extraction threads per stereo frame (src/Frame.cc:164-171), "
                                    "cores//4 frames in flight, tracking matchers in frame order"),
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                                 "sample": f"{frames} consecutive frames of the bench sequence, {rigs} rigs x 4 threads ({dt:.1f} s)",
                                 "single_thread_value": fps1},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    import orb_line_slam_b200 as olf
    from orb_line_slam_b200.frame import FrontEnd
    from orb_line_slam_b200.synth import CAMERAS
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = olf.load_library()
    import ctypes
    lib.olf_kernel_launch_count.restype = ctypes.c_longlong
    api = olf.api(local)
    cam = CAMERAS[WORKLOAD["camera"]]
    # concurrent stereo rigs per GPU: the LSD grow phases are latency-bound, so frames in flight are what fills the GPU
    # a device has 32 hardware work queues: 2 streams per rig + the tracker's + the call-by-call extractors' (idle) must fit
    P = args.pipelines or max(2, min(13, (os.cpu_count() or 16) // max(1, world)))
    sc, seq, poses = make_sequence(N_DISTINCT)
    # weak scaling: every rank runs the same number of frames of its own slice of the sequence
    shift = rank * 11
    seq = seq[shift:] + seq[:shift]; poses = poses[shift:] + poses[:shift]
    dev_imgs = [(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()) for L, R in seq]
    dev_ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in dev_imgs]
    host_pinned = [(torch.from_numpy(L).pin_memory().numpy(), torch.from_numpy(R).pin_memory().numpy()) for L, R in seq]
    # the call-by-call FrontEnd object only serves the tracking matchers; extraction goes through the native rigs
    fes = [FrontEnd(api, cam, WORKLOAD["nfeatures"], WORKLOAD["nlines"], WORKLOAD["min_line_length"])]
    pipes = [fes[0].native(WORKLOAD["nfeatures"], WORKLOAD["nlines"], max_frames=args.batch) for _ in range(P)]
    gather_buf = None
    if world > 1:
        nbytes = int(pipes[0].off.total)
        gather_buf = (torch.empty(nbytes, dtype=torch.uint8, device="cuda"), torch.empty(world * nbytes, dtype=torch.uint8, device="cuda"))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(count, first, on_device):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.olf_kernel_launch_count()
        ev0.record()
        st = run_steps(pipes, pipes, fes, host_pinned, poses, first, count, on_device, dev_ptrs, args.trackers, not args.no_track)
        if dist is not None:
            # trivial NCCL gather of the fixed-capacity keypoint/descriptor block of the rank's last frame
            dist.all_gather_into_tensor(gather_buf[1], gather_buf[0])
        torch.cuda.synchronize()
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms, st, lib.olf_kernel_launch_count() - l0

    # untimed pre-warm: a fresh box needs a few seconds of GPU work before clocks, lazily loaded modules, pinned-page
    # mappings and the host threads settle (the first seconds show multi-100-ms stalls); then the W warm-up steps
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < args.prewarm_s:
        timed(max(args.warmup, 2 * P * args.batch), 0, True)       # every rig sees at least two batches (buffers sized, modules loaded)
    timed(max(args.warmup, P * args.batch), 0, True)               # the W warm-up steps (rounded up to one batch per rig)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, st, launches = timed(args.steps, args.warmup, True)       # HBM-resident inputs
    clocks = sampler.stop() if sampler else None
    timed(max(min(args.warmup, 8), P * args.batch), 0, False)
    ms_e2e, st_e2e, _ = timed(args.steps, args.warmup, False)     # host buffers in, host results out
    # live timing of the dominant kernel (k_lsd_grow): CUDA events on its own stream inside the library
    grow_us = []
    stats8 = (ctypes.c_int * 8)()
    lib.olf_frontend_line.restype = ctypes.c_void_p
    for nat in pipes:
        for eye in range(1):                           # slot 0 carries the timing of the rig's batched chain
            lh = lib.olf_frontend_line(nat.handle, eye)
            if lh and lib.olf_line_last_stats(ctypes.c_void_p(lh), stats8) == 0:
                grow_us.append(stats8[3] / max(stats8[4], 1))       # device time of the batched chain / images in it
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        S = 1536 * 864
        alg_bytes = 6 * S                                          # read angle+used once per px (5S), write used (S)
        grow_ms = float(np.mean(grow_us)) / 1000.0 if grow_us else None
        achieved = (alg_bytes / (grow_ms * 1e-3) / 1e9) if grow_ms else None
        fps = world * args.steps / (ms / 1000.0)
        fps_e2e = world * args.steps / (ms_e2e / 1000.0)
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": dict(WORKLOAD, pipelines_per_gpu=P, frames_per_call=args.batch, l2="inputs larger than L2: 80 distinct stereo pairs = 147 MB cycled",
                               host_cores=os.cpu_count(), per_frame={k: v / max(args.steps, 1) for k, v in st.items()}),
                "clocks": clocks,
                "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": 4 * 1280 * 720, "d2h_bytes_per_step": int(st_e2e["d2h"] / max(args.steps, 1)),
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches),
                "roofline": {"kernel": "k_lsd_grow (+scan, verify)", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)", "unit": "GB/s",
                             "frac": (achieved / peak) if achieved else None,
                             # DRAM bytes per image of the chain, cold caches, from the ncu launch list of the single-frame path
                             # (profiles/r01d_launches_summary.csv: scan 268 + verify 155 + grow 42 MB); the batched chain of the bench
                             # plans smaller waves and moves less
                             "traffic": 465.0e6, "kernel_ms": grow_ms,
                             "algorithmic_bytes_per_launch": alg_bytes,
                             "note": "dominant chain = the LSD region-growing passes (k_lsd_scan / k_lsd_verify / k_lsd_grow, ~95% of a frame's GPU time): device time of one batched chain / images in it, CUDA events on its stream; 'launch' = one image's chain; latency-bound sequential region growing -- the HBM fraction is honest but not the limiter (DESIGN.md section 5)"}}
        if not args.no_cpu_baseline and world == 1:
            frames = args.cpu_frames
            cfps, dt = cpu_reference(frames, [(a, b) for a, b in host_pinned[:N_BASE_FRAMES]], poses[:N_BASE_FRAMES], warm=1)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": 1, "kind": "port",
                                    "sample": f"{frames} consecutive frames of the same sequence through oracle/ (single thread, {dt:.1f} s)"}
        print(json.dumps(line), flush=True)
    for nat in pipes:
        nat.close()
    for fe in fes:
        fe.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
