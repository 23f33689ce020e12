#!/usr/bin/env python3
"""bench.py -- stereo frames/sec of the ORB+line front end (extract + match) on 1..8 B200.

Workload = BASELINE.json configs[1]: 1280x720 synthetic stereo sequence, 2000 ORB + 500 LBD per eye, every frame through
  ExtractORB(L), ExtractORB(R), ExtractLine(L), ExtractLine(R), ComputeStereoMatches, ComputeStereoMatches_Lines,
  ORBmatcher::SearchByProjection(cur,last) and LineMatcher::match(last,cur)                      

A "step" = ONE olf_frontend_process_batch call (B independent stereo frames) on EVERY stereo rig of a GPU: P rigs x B frames
= `frames_per_step` frames per GPU (52 by default), so the pipeline is full at any --steps and `value` = steps x
frames_per_step x n_gpus / time.  The rigs free-run through the K steps of a timed region (no barrier between steps); the
frame-to-frame matchers run on (f, f-1) in frame order on tracker threads.
`value`  : frames/s with the images already resident in HBM (on_device=1);
`e2e`    : the same through the host-buffer entry point (pinned host images in, host result blocks out, copies inside the
           timed region).  Results always come back to host memory (Tracking.cc consumes host vectors).
Multi-GPU (torchrun, one rank per GPU): frames are independent -> every rank runs its own slice of the sequence (weak
scaling); the one exchange is an NCCL all-gather of the step's fixed-capacity result blocks, issued once per
step by a gather thread, inside the timed region.  Every rank issues exactly the same collective sequence (fixed counts).
`--impl reference` times the CPU restatement of the reference path (oracle/) with every host thread it can use; the
`cpu_baseline` object of the GPU arm is the north star's single-thread number.
`--config c3|c4|c5` report the other BASELINE.json configurations (not the driver's headline line).
"""
from __future__ import annotations
import argparse, json, os, pathlib, subprocess, sys, threading, time

# more hardware work queues than the default 8: every rig drives its own streams (must be set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np

WORKLOADS = {
    "c2": dict(workload="1280x720 stereo synthetic seq, 2000 ORB + 500 LBD per eye, extract + stereo match + frame-to-frame match",
               camera="zed720", nfeatures=2000, nlines=500, min_line_length=0.025, has_lines=True),
    "c3": dict(workload="KITTI-00-shaped 1241x376 stereo synthetic seq, 2000 ORB per eye, line path off, extract + stereo match + frame-to-frame match",
               camera="kitti", nfeatures=2000, nlines=0, min_line_length=0.025, has_lines=False),
    "c4": dict(workload="1280x720 stereo synthetic seq, 4000 ORB + 1000 LBD per eye, extract + stereo match + frame-to-frame match",
               camera="zed720", nfeatures=4000, nlines=1000, min_line_length=0.025, has_lines=True),
}
WORKLOAD = WORKLOADS["c2"]
METRIC = "stereo frames/sec (extract+match, 1280x720)"
N_BASE_FRAMES = 8            # distinct rendered poses
N_DISTINCT = 80              # distinct stereo pairs (base frames + per-frame noise): 147 MB of input > 126 MB L2
RING = 6                     # result-block ring: step slots a rig may run ahead of the consumers (tracker, gather)


def make_sequence(n_distinct: int, wl=None):
    """Seeded synthetic stereo sequence.  Rendering is the slow part, so N_BASE_FRAMES poses are rendered
    and every further pair is a base pair with fresh N(0,2) sensor noise -> all pairs are distinct images."""
    from orb_line_slam_b200.synth import Scene, pose_f32
    wl = wl or WORKLOAD
    sc = Scene(wl["camera"], 0)
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
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


class StepRunner:
    """P stereo rigs x B frames per call on one GPU.  Result blocks live in a pinned ring of RING step slots (the library
    writes them, the trackers and the gather thread read them); a rig starts step s only when slot s % RING is free."""

    def __init__(self, torch, dist, pipes, track_fe, host_imgs, dev_ptrs, poses, batch, n_trackers, do_track, local):
        self.torch, self.dist, self.pipes, self.track_fe = torch, dist, pipes, track_fe
        self.host_imgs, self.dev_ptrs, self.poses = host_imgs, dev_ptrs, poses
        self.P, self.B, self.T, self.do_track, self.local = len(pipes), batch, max(1, n_trackers), do_track, local
        self.fps_step = self.P * self.B
        self.nbytes = int(pipes[0].off.total)
        self.ring_t = torch.zeros((RING, self.fps_step, self.nbytes), dtype=torch.uint8).pin_memory()
        self.ring = self.ring_t.numpy()
        self.world = dist.get_world_size() if dist is not None else 1
        self.send = self.recv = self.gstream = None
        if dist is not None:
            self.send = torch.empty(self.fps_step * self.nbytes, dtype=torch.uint8, device="cuda")
            self.recv = torch.empty(self.world * self.fps_step * self.nbytes, dtype=torch.uint8, device="cuda")
            self.gstream = torch.cuda.Stream()
        self.gather_checked = None
        # persistent host threads (one per rig, the trackers, the gatherer): the matchers keep per-thread scratch on the device,
        # and creating it costs device allocations -- nothing of that kind may happen inside a timed region
        from concurrent.futures import ThreadPoolExecutor
        self.pool_w, self.pool_t, self.pool_g = ThreadPoolExecutor(self.P), ThreadPoolExecutor(self.T), ThreadPoolExecutor(1)

    def block(self, k):
        s, j = divmod(k, self.fps_step)
        return self.ring[s % RING, j]

    def run(self, nsteps, first, on_device):
        P, B, T = self.P, self.B, self.T
        count = nsteps * self.fps_step
        nseq = len(self.host_imgs)
        done = [threading.Event() for _ in range(count)]
        views = [None] * count
        consumers = T + (1 if self.dist is not None else 0)
        left = [consumers] * nsteps
        released = [threading.Event() for _ in range(nsteps)]
        lock = threading.Lock()
        stats = {"matches": 0, "line_matches": 0, "kps": 0, "lines": 0, "d2h": 0}
        errors = []
        calls = [[] for _ in range(P)]                    # per rig: (start, end, wait-for-ring) of every batch call, seconds

        def consumer_done(s):
            with lock:
                left[s] -= 1
                if left[s] == 0:
                    released[s].set()

        def fail(e):
            errors.append(e)
            for ev in done + released:
                ev.set()

        def worker(p):
            try:
                nat = self.pipes[p]
                for s in range(nsteps):
                    tw = time.perf_counter()
                    if s >= RING:
                        released[s - RING].wait()
                    if errors:
                        return
                    tc = time.perf_counter()
                    ks = [(s * P + p) * B + j for j in range(B)]
                    idx = [(first + k) % nseq for k in ks]
                    blks = [self.block(k) for k in ks]
                    if on_device:
                        nat.process_batch([self.dev_ptrs[i][0] for i in idx], [self.dev_ptrs[i][1] for i in idx], blks, on_device=True)
                    else:
                        nat.process_batch([self.host_imgs[i][0] for i in idx], [self.host_imgs[i][1] for i in idx], blks, on_device=False)
                    calls[p].append((tc, time.perf_counter(), tc - tw))
                    for k, i, blk in zip(ks, idx, blks):
                        views[k] = nat.view(blk, self.poses[i])
                        done[k].set()
            except Exception as e:       # noqa: BLE001
                fail(e)

        def tracker(tid):
            # frame-to-frame matchers of TrackWithMotionModelWithLine on (f, f-1): src/Tracking.cc:1296 (SearchByProjection) and
            # :1308 (line match); the pairs are independent, tracker thread `tid` takes frames tid, tid+T, ...
            try:
                loc = {"matches": 0, "line_matches": 0, "kps": 0, "lines": 0, "d2h": 0}
                rel = 0                                                # steps below `rel` are released by this tracker
                for k in range(tid, count, T):
                    done[k].wait()
                    if k > 0:
                        done[k - 1].wait()
                    if errors:
                        return
                    v = views[k]
                    if k > 0 and self.do_track:
                        t = self.track_fe[tid % len(self.track_fe)].track(v, views[k - 1])
                        loc["matches"] += t["nmatches"]; loc["line_matches"] += t.get("n_line_matches", 0)
                    nl = 0 if v.kls is None else len(v.kls) + len(v.kls_r)
                    loc["kps"] += len(v.kps) + len(v.kps_r); loc["lines"] += nl
                    loc["d2h"] += (len(v.kps) + len(v.kps_r)) * 56 + len(v.kps) * 8 + nl * 100 + (0 if v.kls is None else len(v.kls) * 36)
                    # my next frame is k+T and reads frames k+T-1 and k+T: every step that ends before k+T-1 is no longer needed
                    while rel < nsteps and (rel + 1) * self.fps_step - 1 < k + T - 1:
                        consumer_done(rel); rel += 1
                while rel < nsteps:
                    consumer_done(rel); rel += 1
                with lock:
                    for key, val in loc.items():
                        stats[key] += val
            except Exception as e:       # noqa: BLE001
                fail(e)

        def gatherer():
            # the one exchange of the sharded front end: all-gather of the step's fixed-capacity result blocks
            try:
                torch = self.torch
                torch.cuda.set_device(self.local)
                with torch.cuda.stream(self.gstream):
                    for s in range(nsteps):
                        for k in range(s * self.fps_step, (s + 1) * self.fps_step):
                            done[k].wait()
                        if errors:
                            return
                        self.send.copy_(self.ring_t[s % RING].view(-1), non_blocking=True)
                        self.gstream.synchronize()
                        consumer_done(s)
                        self.dist.all_gather_into_tensor(self.recv, self.send)
                    self.gstream.synchronize()
                    if self.gather_checked is None:
                        # my own slice of the gathered buffer is my send buffer; every other rank's blocks carry a sane header
                        rank = self.dist.get_rank()
                        g = self.recv.view(self.world, self.fps_step, self.nbytes)
                        own = bool(torch.equal(g[rank].reshape(-1), self.send))
                        hdr = g[:, :, :32].contiguous().view(torch.int32).cpu().numpy().reshape(self.world, self.fps_step, 8)
                        self.gather_checked = own and bool((hdr[:, :, 0] > 0).all()) and bool((hdr[:, :, 6] == 0).all())
            except Exception as e:       # noqa: BLE001
                fail(e)

        # every pool has exactly as many threads as tasks, so all tasks run concurrently (they wait for each other)
        futs = [self.pool_w.submit(worker, p) for p in range(P)] + [self.pool_t.submit(tracker, i) for i in range(T)]
        if self.dist is not None:
            futs.append(self.pool_g.submit(gatherer))
        for f in futs:
            f.result()
        if errors:
            raise errors[0]
        # how steady the rigs ran: duration of the batch calls, time between calls (host glue), time spent waiting for a ring slot
        dur = np.array([e - b for c in calls for (b, e, w) in c]); wait = np.array([w for c in calls for (b, e, w) in c])
        gap = np.array([c[i + 1][0] - c[i][1] - c[i + 1][2] for c in calls for i in range(len(c) - 1)] or [0.0])
        stats["call_ms"] = dict(p50=round(float(np.percentile(dur, 50)) * 1e3, 1), p99=round(float(np.percentile(dur, 99)) * 1e3, 1), max=round(float(dur.max()) * 1e3, 1))
        stats["gap_ms"] = dict(p50=round(float(np.percentile(gap, 50)) * 1e3, 2), max=round(float(gap.max()) * 1e3, 1))
        stats["ring_wait_ms"] = dict(total=round(float(wait.sum()) * 1e3, 1), max=round(float(wait.max()) * 1e3, 1))
        return stats


def cpu_reference(frames, seq, poses, warm=1, wl=None):
    """Single-thread CPU restatement of the reference path (oracle) on `frames` frames; returns (fps, seconds)."""
    from orc import oracle
    from orb_line_slam_b200.frame import FrontEnd
    from orb_line_slam_b200.synth import CAMERAS
    wl = wl or WORKLOAD
    fe = FrontEnd(oracle(), CAMERAS[wl["camera"]], wl["nfeatures"], wl["nlines"], wl["min_line_length"], has_lines=wl["has_lines"])
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


def cpu_reference_all_cores(frames, seq, poses, warm=1, cores=None, wl=None):
    """The reference's own threading on every host core: each rig runs Frame::Frame's four extraction threads (ORB L/R,
    lines L/R, src/Frame.cc:164-171) followed by the stereo matchers; cores//4 rigs work on independent frames at the
    same time (the same frames-in-flight arrangement as the GPU arm) and one tracker consumes them in order.
    The oracle is a ctypes library, so the GIL is released inside every call.  Returns (fps, seconds, threads)."""
    from concurrent.futures import ThreadPoolExecutor
    from orc import oracle
    from orb_line_slam_b200.frame import FrontEnd, StereoFrame
    from orb_line_slam_b200.synth import CAMERAS
    wl = wl or WORKLOAD
    cores = cores or os.cpu_count() or 4
    rigs = max(1, cores // 4)
    a = oracle()
    fes = [FrontEnd(a, CAMERAS[wl["camera"]], wl["nfeatures"], wl["nlines"], wl["min_line_length"], has_lines=wl["has_lines"]) for _ in range(rigs)]
    pools = [ThreadPoolExecutor(4) for _ in range(rigs)]

    def frame(r, k):
        fe, L, R = fes[r], *seq[k % len(seq)]
        jobs = [pools[r].submit(a.orb_extract, fe.orb_l, L), pools[r].submit(a.orb_extract, fe.orb_r, R)]
        if fe.has_lines:
            jobs += [pools[r].submit(a.line_extract, fe.line_l, L), pools[r].submit(a.line_extract, fe.line_r, R)]
        res = [j.result() for j in jobs]
        (kl, dl), (kr, dr) = res[:2]
        f = StereoFrame(kl, dl, kr, dr, None, None)
        if fe.has_lines:
            (f.kls, f.ldesc), (f.kls_r, f.ldesc_r) = res[2:]
        f.u_right, f.depth = a.stereo_points(fe.orb_l, fe.orb_r, kl, dl, kr, dr, fe.bf, fe.fx)
        if fe.has_lines:
            f.line_matches, f.line_disp, f.line_le = a.stereo_lines(f.kls, f.ldesc, f.kls_r, f.ldesc_r, fe.w, fe.h, fe.lmp)
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


def load_peaks():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"], help="BASELINE.json configuration (c2 = the headline workload)")
    ap.add_argument("--pipelines", type=int, default=0, help="concurrent stereo rigs per GPU (0 = auto)")
    ap.add_argument("--trackers", type=int, default=2, help="host threads running the frame-to-frame matchers")
    ap.add_argument("--no-track", action="store_true", help="(diagnostic) skip the frame-to-frame matchers")
    ap.add_argument("--batch", type=int, default=4, help="independent stereo frames per olf_frontend_process_batch call (1..8)")
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames of the bounded single-thread CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prewarm-steps", type=int, default=60, help="untimed steps before the warm-up steps (fixed count: rank-invariant)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3
    if args.config == "c5":
        if rank == 0:
            from tools.match_bench import run_c5
            run_c5(local)
        return
    wl = WORKLOADS[args.config]
    metric = METRIC if args.config != "c3" else "stereo frames/sec (extract+match, 1241x376, lines off)"

    if args.impl == "reference":
        if rank != 0:
            return
        sc, seq, poses = make_sequence(N_BASE_FRAMES, wl)
        cores = os.cpu_count() or 4
        rigs = max(1, cores // 4)
        # one "step" = one batch of 4 frames on every CPU rig (the GPU arm's step with the CPU's number of rigs); the sample
        # is bounded so that K steps + W warm-up end within a few minutes
        per_step = 4 * rigs
        steps = max(1, min(args.steps, 30)); warm_steps = max(1, min(args.warmup, 5))
        frames = steps * per_step
        fps, dt, threads = cpu_reference_all_cores(frames, seq, poses, warm=warm_steps * per_step, cores=cores, wl=wl)
        fps1, dt1 = cpu_reference(8, seq, poses, warm=1, wl=wl)
        line = {"impl": "reference", "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm_steps,
                "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": dict(wl, frames_per_step=per_step, host_cores=cores, rigs=rigs, threads=threads,
                               note="CPU restatement of the reference path (oracle/; the reference itself needs OpenCV C++/Eigen/Pangolin and cannot be "
                                    "built here), run the reference's way: 4 extraction threads per stereo frame (src/Frame.cc:164-171), "
                                    "cores//4 frames in flight, tracking matchers in frame order; a step = 4 frames on each of the cores//4 rigs"),
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
    lib.olf_alloc_count.restype = ctypes.c_longlong
    api = olf.api(local)
    cam = CAMERAS[wl["camera"]]
    # concurrent stereo rigs per GPU: the LSD grow phases are latency-bound, so frames in flight are what fills the GPU
    # (measured on one B200: 13 rigs 970, 16 rigs 1040, 20 rigs 1120, 22 rigs 1135, 26 rigs 995 frames/s).  A rig costs one mostly
    # sleeping host thread and one CUDA stream, so the number of rigs is the same at every GPU count.
    cores = os.cpu_count() or 16
    P = args.pipelines or 20
    B = max(1, min(8, args.batch))
    sc, seq, poses = make_sequence(N_DISTINCT, wl)
    # weak scaling: every rank runs the same number of frames of its own slice of the sequence
    shift = (rank * 11) % len(seq)
    seq = seq[shift:] + seq[:shift]; poses = poses[shift:] + poses[:shift]
    dev_imgs = [(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()) for L, R in seq]
    dev_ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in dev_imgs]
    host_pinned = [(torch.from_numpy(L).pin_memory().numpy(), torch.from_numpy(R).pin_memory().numpy()) for L, R in seq]
    # the call-by-call FrontEnd objects only serve the tracking matchers; extraction goes through the native rigs
    T = max(1, args.trackers)
    fes = [FrontEnd(api, cam, wl["nfeatures"], wl["nlines"], wl["min_line_length"], has_lines=wl["has_lines"]) for _ in range(T)]
    pipes = [fes[0].native(wl["nfeatures"], wl["nlines"], max_frames=B) for _ in range(P)]
    runner = StepRunner(torch, dist, pipes, fes, host_pinned, dev_ptrs, poses, B, T, not args.no_track, local)
    fps_step = runner.fps_step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, first, on_device):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.olf_kernel_launch_count(); a0 = lib.olf_alloc_count()
        ev0.record()
        st = runner.run(nsteps, first, on_device)
        torch.cuda.synchronize()
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        st["allocs"] = int(lib.olf_alloc_count() - a0)
        return ms, st, lib.olf_kernel_launch_count() - l0

    # untimed pre-warm, a FIXED number of steps on every rank (the collective sequence must be rank-invariant): a fresh box
    # needs a few seconds of GPU work before clocks, lazily loaded modules, pinned-page mappings and host threads settle
    timed(max(2, args.prewarm_steps), 0, True)
    if os.environ.get("OLF_BENCH_BURN"):
        # diagnostic (tools/burn.cu): a persistent kernel takes a known share of every scheduler's issue slots during the timed regions
        chains, secs = os.environ["OLF_BENCH_BURN"].split(",")
        burn = ctypes.CDLL(str(ROOT / "tools" / "libburn.so"))
        burn.burn_start.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double]
        assert burn.burn_start(local, int(chains), float(secs)) == 0
    timed(args.warmup, 0, True)                                      # the W warm-up steps
    sampler = ClockSampler(local) if rank == 0 else None
    ms, st, launches = timed(args.steps, args.warmup * fps_step, True)       # HBM-resident inputs
    clocks = sampler.stop() if sampler else None
    timed(args.warmup, 0, False)
    ms_e2e, st_e2e, _ = timed(args.steps, args.warmup * fps_step, False)     # host buffers in, host results out
    # live timing of the dominant kernels (the LSD region-growing chain): CUDA events on its own stream inside the library
    grow_us = []
    stats8 = (ctypes.c_int * 8)()
    lib.olf_frontend_line.restype = ctypes.c_void_p
    host_us = []
    for nat in pipes:
        t8 = (ctypes.c_int * 8)()
        if lib.olf_frontend_last_timing(nat.handle, t8) == 0:
            host_us.append(list(t8))
    line_us = []
    if wl["has_lines"]:
        for nat in pipes:
            lh = lib.olf_frontend_line(nat.handle, 0)              # slot 0 carries the timing of the rig's batched chain
            if lh and lib.olf_line_last_stats(ctypes.c_void_p(lh), stats8) == 0:
                grow_us.append(stats8[3] / max(stats8[4], 1))       # device time of the batched chain / images in it
                line_us.append([stats8[5], stats8[6], stats8[7]])
    if rank == 0:
        peaks = load_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        nframes = args.steps * fps_step
        fps = world * nframes / (ms / 1000.0)
        fps_e2e = world * nframes / (ms_e2e / 1000.0)
        w, h = cam[0], cam[1]
        line = {"metric": metric, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": dict(wl, step="one olf_frontend_process_batch call on every rig", pipelines_per_gpu=P, frames_per_call=B, frames_per_step=fps_step,
                               frames_timed_per_gpu=nframes, trackers=T, prewarm_steps=args.prewarm_steps,
                               l2=f"inputs larger than L2: {N_DISTINCT} distinct stereo pairs = {N_DISTINCT * 2 * w * h / 1e6:.0f} MB cycled",
                               host_cores=cores, per_frame={k: v / max(nframes, 1) for k, v in st.items() if not isinstance(v, dict) and k != "allocs"},
                               steadiness=dict(resident=dict({k: v for k, v in st.items() if isinstance(v, dict)}, allocs=st["allocs"]),
                                               e2e=dict({k: v for k, v in st_e2e.items() if isinstance(v, dict)}, allocs=st_e2e["allocs"])),
                               exchange=(None if dist is None else f"NCCL all_gather of {fps_step} result blocks x {runner.nbytes} B per rank per step"),
                               gather_verified=runner.gather_checked,
                               rig_call_ms=(None if not host_us else dict(zip(("orb_enqueue", "lines", "orb_wait", "collect", "stereo_lines", "total"),
                                                                              [round(float(v) / 1000, 2) for v in np.mean(np.array(host_us)[:, :6], 0)]))),
                               line_call_ms=(None if not line_us else dict(zip(("enqueue_and_chain", "rect", "keylines_lbd"),
                                                                               [round(float(v) / 1000, 2) for v in np.mean(np.array(line_us), 0)])))),
                "clocks": clocks,
                "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": 2 * w * h * fps_step,
                        "d2h_bytes_per_step": int(st_e2e["d2h"] / max(args.steps, 1)), "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches)}
        if grow_us:
            S = int(round(w * 1.2)) * int(round(h * 1.2))
            alg_bytes = 6 * S                                          # read angle+used once per px (5S), write used (S)
            grow_ms = float(np.mean(grow_us)) / 1000.0
            achieved = alg_bytes / (grow_ms * 1e-3) / 1e9
            traffic = None
            try:
                traffic = json.loads((ROOT / "profiles" / "lsd_chain_traffic.json").read_text()).get("dram_bytes_per_image")
            except Exception:
                pass
            line["roofline"] = {"kernel": "k_lsd_grow (+scan, verify)", "bound": "hbm", "achieved": achieved, "peak": peak,
                                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)", "unit": "GB/s",
                                "frac": achieved / peak, "traffic": traffic, "kernel_ms": grow_ms, "algorithmic_bytes_per_launch": alg_bytes,
                                "note": "dominant chain = the LSD region-growing passes (k_lsd_scan / k_lsd_verify / k_lsd_grow): device time of one "
                                        "batched chain / images in it, CUDA events on its stream, while the other rigs share the GPU; 'launch' = one image's "
                                        "chain; traffic = DRAM bytes per image of the same batched chain from the ncu launch list in profiles/ "
                                        "(profiles/lsd_chain_traffic.json); latency-bound sequential region growing -- the HBM fraction is honest "
                                        "but not the limiter (DESIGN.md section 5)"}
        if not args.no_cpu_baseline and world == 1:
            frames = args.cpu_frames
            cfps, dt = cpu_reference(frames, [(a, b) for a, b in host_pinned[:N_BASE_FRAMES]], poses[:N_BASE_FRAMES], warm=1, wl=wl)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": 1, "kind": "port",
                                    "sample": f"{frames} consecutive frames of the same sequence through oracle/ (single thread, {dt:.1f} s)"}
        print(json.dumps(line), flush=True)
    if os.environ.get("OLF_BENCH_BURN"):
        burn.burn_wait.restype = ctypes.c_ulonglong
        print("burn: outer iterations of one warp (x 256 inner iterations)", burn.burn_wait(), file=sys.stderr, flush=True)
    for nat in pipes:
        nat.close()
    for fe in fes:
        fe.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
