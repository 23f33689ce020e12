"""Seeded synthetic stereo sequence: no dataset ships with the reference, so the
benchmark and the parity tests render their own images.  numpy only (must run on the GPU box).

Scene: mid-grey background + random thick line strokes + filled rectangles, each primitive
fronto-parallel at its own depth Z in [2, 30] m; right eye = same scene shifted by d = bf / Z;
frame f moves the camera 2 cm forward and yaws it 0.1 degree; 3x3 box smoothing + N(0,3) noise.
"""
from __future__ import annotations
import numpy as np

CAMERAS = {
    # name: (width, height, fx, fy, cx, cy, bf)  -- Examples/PL/ZED_HD720.yaml, PL_KITTI00-02.yaml, PL_EuRoC.yaml
    "zed720": (1280, 720, 670.44, 670.44, 640.0, 360.0, 80.4534),
    "kitti": (1241, 376, 718.856, 718.856, 607.1928, 185.2157, 386.1448),
    "euroc": (640, 480, 435.2046959714599, 435.2046959714599, 367.4517211914062 * 640 / 752, 252.2008514404297, 47.90639384423901),
}


def _pose(frame: int):
    """World->camera rotation (yaw about y) and translation for frame index `frame`."""
    yaw = np.deg2rad(0.1 * frame)
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]], dtype=np.float64)
    cam_center = np.array([0.0, 0.0, 0.02 * frame])
    t = -R @ cam_center
    return R, t


def pose_f32(frame: int):
    R, t = _pose(frame)
    return R.astype(np.float32), t.astype(np.float32)


class Scene:
    def __init__(self, camera: str = "zed720", seed: int = 0, n_lines: int = 400, n_rects: int = 120):
        self.camera = camera
        self.w, self.h, self.fx, self.fy, self.cx, self.cy, self.bf = CAMERAS[camera]
        rng = np.random.RandomState(seed)
        w, h = self.w, self.h
        prims = []
        for _ in range(n_rects):
            x0, y0 = rng.uniform(-0.1 * w, w), rng.uniform(-0.1 * h, h)
            rw, rh = rng.uniform(12, 0.2 * w), rng.uniform(12, 0.25 * h)
            prims.append(("rect", x0, y0, x0 + rw, y0 + rh, 0.0, rng.uniform(2, 30), int(rng.randint(20, 236))))
        for _ in range(n_lines):
            x0, y0 = rng.uniform(0, w), rng.uniform(0, h)
            ang, ln = rng.uniform(0, np.pi), rng.uniform(20, 300)
            prims.append(("line", x0, y0, x0 + ln * np.cos(ang), y0 + ln * np.sin(ang), rng.uniform(1, 8),
                          rng.uniform(2, 30), int(rng.randint(10, 246))))
        # far-to-near painter's order
        self.prims = sorted(prims, key=lambda p: -p[6])
        self.seed = seed

    def _project(self, u, v, Z, R, t, eye_shift):
        X = np.array([(u - self.cx) * Z / self.fx, (v - self.cy) * Z / self.fy, Z])
        Xc = R @ X + t
        Xc[0] -= eye_shift
        z = max(Xc[2], 0.1)
        return self.fx * Xc[0] / z + self.cx, self.fy * Xc[1] / z + self.cy, z

    def render(self, frame: int = 0, eye: int = 0) -> np.ndarray:
        """eye 0 = left, 1 = right (baseline b = bf/fx along +x)."""
        R, t = _pose(frame)
        shift = (self.bf / self.fx) if eye else 0.0
        img = np.full((self.h, self.w), 128.0, dtype=np.float32)
        for kind, x0, y0, x1, y1, width, Z, grey in self.prims:
            ax, ay, za = self._project(x0, y0, Z, R, t, shift)
            bx, by, _ = self._project(x1, y1, Z, R, t, shift)
            if kind == "rect":
                xa, xb = sorted((int(round(ax)), int(round(bx))))
                ya, yb = sorted((int(round(ay)), int(round(by))))
                xa, xb, ya, yb = max(xa, 0), min(xb, self.w), max(ya, 0), min(yb, self.h)
                if xa < xb and ya < yb:
                    img[ya:yb, xa:xb] = grey
            else:
                hw = 0.5 * width * Z / za
                xa, xb = int(np.floor(min(ax, bx) - hw - 1)), int(np.ceil(max(ax, bx) + hw + 2))
                ya, yb = int(np.floor(min(ay, by) - hw - 1)), int(np.ceil(max(ay, by) + hw + 2))
                xa, xb, ya, yb = max(xa, 0), min(xb, self.w), max(ya, 0), min(yb, self.h)
                if xa >= xb or ya >= yb:
                    continue
                yy, xx = np.mgrid[ya:yb, xa:xb].astype(np.float32)
                dx, dy = bx - ax, by - ay
                L2 = dx * dx + dy * dy + 1e-9
                tt = np.clip(((xx - ax) * dx + (yy - ay) * dy) / L2, 0, 1)
                dist2 = (xx - (ax + tt * dx)) ** 2 + (yy - (ay + tt * dy)) ** 2
                img[ya:yb, xa:xb][dist2 <= hw * hw] = grey
        # 3x3 box smoothing (edge-replicated) + noise
        p = np.pad(img, 1, mode="edge")
        sm = sum(p[i:i + self.h, j:j + self.w] for i in range(3) for j in range(3)) / 9.0
        rng = np.random.RandomState(1000003 * (self.seed + 1) + 2 * frame + eye)
        sm = sm + rng.normal(0.0, 3.0, sm.shape)
        return np.clip(np.rint(sm), 0, 255).astype(np.uint8)

    def stereo(self, frame: int = 0):
        return self.render(frame, 0), self.render(frame, 1)


def random_image(w: int, h: int, seed: int) -> np.ndarray:
    """Small scene-like test image of arbitrary size (parity tests)."""
    rng = np.random.RandomState(seed)
    img = np.full((h, w), 128.0, dtype=np.float32)
    for _ in range(max(4, (w * h) // 6000)):
        x0, y0 = rng.randint(0, w), rng.randint(0, h)
        x1, y1 = min(w, x0 + rng.randint(6, max(8, w // 3))), min(h, y0 + rng.randint(6, max(8, h // 3)))
        img[y0:y1, x0:x1] = rng.randint(10, 246)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for _ in range(max(4, (w * h) // 4000)):
        ax, ay = rng.uniform(0, w), rng.uniform(0, h)
        ang, ln, hw = rng.uniform(0, np.pi), rng.uniform(10, max(12, w // 2)), rng.uniform(0.5, 4)
        bx, by = ax + ln * np.cos(ang), ay + ln * np.sin(ang)
        dx, dy = bx - ax, by - ay
        tt = np.clip(((xx - ax) * dx + (yy - ay) * dy) / (dx * dx + dy * dy), 0, 1)
        d2 = (xx - (ax + tt * dx)) ** 2 + (yy - (ay + tt * dy)) ** 2
        img[d2 <= hw * hw] = rng.randint(10, 246)
    p = np.pad(img, 1, mode="edge")
    sm = sum(p[i:i + h, j:j + w] for i in range(3) for j in range(3)) / 9.0
    sm = sm + rng.normal(0.0, 3.0, sm.shape)
    return np.clip(np.rint(sm), 0, 255).astype(np.uint8)
