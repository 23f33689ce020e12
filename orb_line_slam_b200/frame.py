"""Host-side mirror of the reference's Frame for the front end: runs the per-frame pipeline through a FrontEndApi
(product `olf_*` or oracle `orc_*`) and builds the SoA query blocks the tracking matchers take.

Mirrors Frame::Frame(stereo+lines) (reference src/Frame.cc:136-221): ExtractORB x2, ExtractLine x2,
ComputeStereoMatches, ComputeStereoMatches_Lines; and the gathering a shim does before
ORBmatcher::SearchByProjection / LineMatcher::match (src/Tracking.cc:1296-1308)."""
from __future__ import annotations
import ctypes as C
from dataclasses import dataclass, field
import numpy as np
from .abi import (FrontEndApi, LineParams, LineMatchParams, Camera, SbpLastArgs, SbpMapArgs, KEYPOINT, KEYLINE, ptr,
                  FrontendParams, FrameHeader, FrameOffsets)


@dataclass
class StereoFrame:
    kps: np.ndarray
    desc: np.ndarray
    kps_r: np.ndarray
    desc_r: np.ndarray
    u_right: np.ndarray
    depth: np.ndarray
    kls: np.ndarray = None
    ldesc: np.ndarray = None
    kls_r: np.ndarray = None
    ldesc_r: np.ndarray = None
    line_matches: np.ndarray = None
    line_disp: np.ndarray = None
    line_le: np.ndarray = None
    Rcw: np.ndarray = None
    tcw: np.ndarray = None


class FrontEnd:
    """One stereo rig: two ORB extractors, two line extractors (like Tracking's mpORBextractorLeft/Right,
    mpLineextractorLeft/Right, src/Tracking.cc:131-138)."""

    def __init__(self, api: FrontEndApi, camera, nfeatures=2000, nlines=500, min_line_length=0.025, has_lines=True,
                 line_match: LineMatchParams | None = None):
        self.api = api
        self.w, self.h, self.fx, self.fy, self.cx, self.cy, self.bf = camera
        self.has_lines = has_lines
        self.orb_l, self.orb_r = api.orb_create(nfeatures), api.orb_create(nfeatures)
        self.lp = LineParams(lsd_nfeatures=nlines, min_line_length=min_line_length)
        self.line_l = api.line_create(self.lp) if has_lines else None
        self.line_r = api.line_create(self.lp) if has_lines else None
        self.lmp = line_match or LineMatchParams()
        self.scale_factors = api.orb_scale_factors(self.orb_l)[0]
        self.cam = Camera(self.fx, self.fy, self.cx, self.cy, self.bf, 0.0, float(self.w), 0.0, float(self.h))

    def close(self):
        self.api.orb_destroy(self.orb_l); self.api.orb_destroy(self.orb_r)
        if self.has_lines:
            self.api.line_destroy(self.line_l); self.api.line_destroy(self.line_r)

    def process(self, img_l, img_r, pose=None) -> StereoFrame:
        a = self.api
        kl, dl = a.orb_extract(self.orb_l, img_l)
        kr, dr = a.orb_extract(self.orb_r, img_r)
        f = StereoFrame(kl, dl, kr, dr, None, None)
        if self.has_lines:
            f.kls, f.ldesc = a.line_extract(self.line_l, img_l)
            f.kls_r, f.ldesc_r = a.line_extract(self.line_r, img_r)
        f.u_right, f.depth = a.stereo_points(self.orb_l, self.orb_r, kl, dl, kr, dr, self.bf, self.fx)
        if self.has_lines:
            f.line_matches, f.line_disp, f.line_le = a.stereo_lines(f.kls, f.ldesc, f.kls_r, f.ldesc_r, self.w, self.h, self.lmp)
        if pose is not None:
            f.Rcw, f.tcw = pose
        return f

    # ---- native whole-frame path (olf_frontend_*; product library only) ----
    def native(self, nfeatures=2000, nlines=500, cap_points=None, cap_lines=None, max_frames=1):
        """Create the native rig (3 host threads + 2 streams inside libolf.so) that fills one POD block per frame;
        max_frames > 1: olf_frontend_process_batch takes that many independent stereo frames per call."""
        cap_points = cap_points or (nfeatures + 256)
        cap_lines = cap_lines or max(nlines, 64) if nlines else 4096
        p = FrontendParams(nfeatures, 1.2, 8, 20, 7, int(self.has_lines), self.lp, self.lmp, self.cam, cap_points, cap_lines)
        return NativeFrontEnd(self.api, p, self.w, self.h, max_frames)

    # ---- tracking matchers (src/Tracking.cc:1296-1308) ----
    def sbp_last_args(self, cur: StereoFrame, last: StereoFrame, th=7.0, mono=False, check_orientation=True, observed=None):
        """Build olf_sbp_last_args: every last-frame keypoint with a stereo depth stands for a MapPoint
        (Frame::UnprojectStereo, src/Frame.cc:1050 ff.: x3Dw = Rwc * x3Dc + twc)."""
        n_last = len(last.kps)
        has = (last.depth > 0).astype(np.uint8)
        z = np.where(has > 0, last.depth, 1.0).astype(np.float32)
        x = ((last.kps["x"] - np.float32(self.cx)) * z / np.float32(self.fx)).astype(np.float32)
        y = ((last.kps["y"] - np.float32(self.cy)) * z / np.float32(self.fy)).astype(np.float32)
        xc = np.stack([x, y, z], 1).astype(np.float32)
        Rwc = last.Rcw.T.astype(np.float32)
        twc = (-Rwc @ last.tcw).astype(np.float32)
        world = np.ascontiguousarray((xc @ Rwc.T + twc).astype(np.float32))
        obs = np.ones(n_last, np.uint8) if observed is None else np.ascontiguousarray(observed, np.uint8)
        keep = dict(cur_kps=np.ascontiguousarray(cur.kps), cur_desc=np.ascontiguousarray(cur.desc), cur_u=np.ascontiguousarray(cur.u_right),
                    sf=np.ascontiguousarray(self.scale_factors), last_kps=np.ascontiguousarray(last.kps), has=has, obs=obs, world=world,
                    ldesc=np.ascontiguousarray(last.desc))
        a = SbpLastArgs()
        a.cur_kps, a.cur_desc, a.cur_u_right, a.n_cur = ptr(keep["cur_kps"]), ptr(keep["cur_desc"]), ptr(keep["cur_u"]), len(cur.kps)
        a.cam = self.cam
        a.scale_factors, a.nlevels = ptr(keep["sf"]), len(self.scale_factors)
        a.Rcw[:] = list(cur.Rcw.astype(np.float32).ravel()); a.tcw[:] = list(cur.tcw.astype(np.float32))
        a.Rlw[:] = list(last.Rcw.astype(np.float32).ravel()); a.tlw[:] = list(last.tcw.astype(np.float32))
        a.last_kps, a.n_last = ptr(keep["last_kps"]), n_last
        a.last_has_point, a.last_point_observed = ptr(keep["has"]), ptr(keep["obs"])
        a.last_world_pos, a.last_point_desc = ptr(keep["world"]), ptr(keep["ldesc"])
        a.th, a.mono, a.check_orientation = th, int(mono), int(check_orientation)
        return a, keep

    def sbp_map_args(self, cur: StereoFrame, last: StereoFrame, th=1.0, nn_ratio=0.8, occupied=None, observed=None):
        """Build olf_sbp_map_args from the same pseudo map points: what Frame::isInFrustum leaves in
        mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos (src/Frame.cc:388-444)."""
        sel = np.nonzero(last.depth > 0)[0]
        z = last.depth[sel].astype(np.float32)
        x = ((last.kps["x"][sel] - np.float32(self.cx)) * z / np.float32(self.fx)).astype(np.float32)
        y = ((last.kps["y"][sel] - np.float32(self.cy)) * z / np.float32(self.fy)).astype(np.float32)
        Rwc = last.Rcw.T.astype(np.float32); twc = (-Rwc @ last.tcw).astype(np.float32)
        world = (np.stack([x, y, z], 1) @ Rwc.T + twc).astype(np.float32)
        pc = (world @ cur.Rcw.astype(np.float32).T + cur.tcw.astype(np.float32)).astype(np.float32)
        ok = pc[:, 2] > 0
        invz = (np.float32(1.0) / np.where(ok, pc[:, 2], 1)).astype(np.float32)
        u = (np.float32(self.fx) * pc[:, 0] * invz + np.float32(self.cx)).astype(np.float32)
        v = (np.float32(self.fy) * pc[:, 1] * invz + np.float32(self.cy)).astype(np.float32)
        ok &= (u >= 0) & (u <= self.w) & (v >= 0) & (v <= self.h)
        idx = np.nonzero(ok)[0]
        n = len(idx)
        keep = dict(cur_kps=np.ascontiguousarray(cur.kps), cur_desc=np.ascontiguousarray(cur.desc), cur_u=np.ascontiguousarray(cur.u_right),
                    sf=np.ascontiguousarray(self.scale_factors),
                    px=np.ascontiguousarray(u[idx]), py=np.ascontiguousarray(v[idx]),
                    pxr=np.ascontiguousarray((u[idx] - np.float32(self.bf) * invz[idx]).astype(np.float32)),
                    lvl=np.ascontiguousarray(last.kps["octave"][sel][idx].astype(np.int32)),
                    vc=np.ascontiguousarray(np.where(np.arange(n) % 3 == 0, 0.9, 0.999).astype(np.float32)),
                    obs=np.ones(n, np.uint8) if observed is None else np.ascontiguousarray(observed[:n], np.uint8),
                    pdesc=np.ascontiguousarray(last.desc[sel][idx]),
                    occ=None if occupied is None else np.ascontiguousarray(occupied, np.uint8))
        a = SbpMapArgs()
        a.cur_kps, a.cur_desc, a.cur_u_right, a.n_cur = ptr(keep["cur_kps"]), ptr(keep["cur_desc"]), ptr(keep["cur_u"]), len(cur.kps)
        a.cur_occupied = ptr(keep["occ"])
        a.cam = self.cam
        a.scale_factors, a.nlevels = ptr(keep["sf"]), len(self.scale_factors)
        a.n_points = n
        a.proj_x, a.proj_y, a.proj_xr = ptr(keep["px"]), ptr(keep["py"]), ptr(keep["pxr"])
        a.pred_level, a.view_cos = ptr(keep["lvl"]), ptr(keep["vc"])
        a.point_observed, a.point_desc = ptr(keep["obs"]), ptr(keep["pdesc"])
        a.th, a.nn_ratio = th, nn_ratio
        return a, keep

    def track(self, cur: StereoFrame, last: StereoFrame, th=7.0, nnr_lines=0.9):
        """TrackWithMotionModelWithLine's matching calls (src/Tracking.cc:1296, 1308)."""
        args, keep = self.sbp_last_args(cur, last, th)
        assigned, cur_point, n = self.api.search_by_projection_last(args, keep)
        out = dict(assigned=assigned, cur_point=cur_point, nmatches=n)
        if self.has_lines:
            m, nl = self.api.match_lines(last.ldesc, cur.ldesc, nnr_lines, bool(self.lmp.best_lr_matches))
            out.update(line_matches=m, n_line_matches=nl)
        return out


class NativeFrontEnd:
    """olf_frontend_*: Frame::Frame(stereo+lines) in one native call; results land in a fixed-capacity POD block."""

    def __init__(self, api: FrontEndApi, params: FrontendParams, w, h, max_frames=1):
        self.api, self.params, self.w, self.h, self.max_frames = api, params, w, h, max_frames
        self.layout = BlockLayout(api, params.cap_points, params.cap_lines, bool(params.has_lines))
        self.off = self.layout.off
        self.handle = api.frontend_create(params, max_frames)

    def close(self):
        self.api.frontend_destroy(self.handle)

    def new_block(self) -> np.ndarray:
        return np.zeros(int(self.off.total), dtype=np.uint8)

    def process(self, img_l, img_r, block: np.ndarray, on_device=False, stride=None):
        self.api.frontend_process(self.handle, img_l, img_r, self.w, self.h, stride or self.w, on_device, block)
        return block

    def process_batch(self, imgs_l, imgs_r, blocks, on_device=False, stride=None):
        """olf_frontend_process_batch: len(imgs_l) <= max_frames independent stereo frames through one call."""
        self.api.frontend_process_batch(self.handle, imgs_l, imgs_r, self.w, self.h, stride or self.w, on_device, blocks)
        return blocks

    def view(self, block: np.ndarray, pose=None) -> StereoFrame:
        """Zero-copy numpy views into a result block."""
        return self.layout.view(block, pose)


class BlockLayout:
    """The fixed-capacity POD block of one stereo frame (olf_frame_header + arrays at olf_frame_layout's offsets, include/olf_abi.h): what
    olf_frontend_process fills, what the ranks of a sharded sequence exchange."""

    def __init__(self, api: FrontEndApi, cap_points: int, cap_lines: int, has_lines: bool = True):
        self.cap_points, self.cap_lines, self.has_lines = cap_points, cap_lines, has_lines
        self.off = api.frame_layout(cap_points, cap_lines)          # host arithmetic inside libolf.so, no device needed
        self.nbytes = int(self.off.total)

    def _arr(self, block, off, dtype, count, shape=None):
        a = np.frombuffer(block, dtype=dtype, count=count, offset=int(off))
        return a.reshape(shape) if shape else a

    def view(self, block: np.ndarray, pose=None) -> StereoFrame:
        """Zero-copy numpy views into a result block."""
        o, cp, cl = self.off, self.cap_points, self.cap_lines
        hd = FrameHeader.from_buffer(block)
        arr = lambda off, dtype, count, shape=None: self._arr(block, off, dtype, count, shape)      # noqa: E731
        f = StereoFrame(arr(o.kps_l, KEYPOINT, cp)[:hd.n_l], arr(o.desc_l, np.uint8, cp * 32, (cp, 32))[:hd.n_l],
                        arr(o.kps_r, KEYPOINT, cp)[:hd.n_r], arr(o.desc_r, np.uint8, cp * 32, (cp, 32))[:hd.n_r],
                        arr(o.u_right, np.float32, cp)[:hd.n_l], arr(o.depth, np.float32, cp)[:hd.n_l])
        if self.has_lines:
            f.kls = arr(o.kls_l, KEYLINE, cl)[:hd.m_l]; f.ldesc = arr(o.ldesc_l, np.uint8, cl * 32, (cl, 32))[:hd.m_l]
            f.kls_r = arr(o.kls_r, KEYLINE, cl)[:hd.m_r]; f.ldesc_r = arr(o.ldesc_r, np.uint8, cl * 32, (cl, 32))[:hd.m_r]
            f.line_matches = arr(o.lmatch, np.int32, cl)[:hd.m_l]
            f.line_disp = arr(o.ldisp, np.float32, cl * 2, (cl, 2))[:hd.m_l]
            f.line_le = arr(o.lle, np.float64, cl * 3, (cl, 3))[:hd.m_l]
        if pose is not None:
            f.Rcw, f.tcw = pose
        return f

    def pack(self, f: StereoFrame, block: np.ndarray | None = None) -> np.ndarray:
        """The inverse of view(): a frame produced call by call (FrontEnd.process) laid out as olf_frontend_process would.  Raises when the
        frame exceeds the capacities, as the native call reports OLF_ERR_CAPACITY."""
        block = np.zeros(self.nbytes, np.uint8) if block is None else block
        o, cp, cl = self.off, self.cap_points, self.cap_lines
        n_l, n_r = len(f.kps), len(f.kps_r)
        m_l, m_r = (len(f.kls), len(f.kls_r)) if self.has_lines and f.kls is not None else (0, 0)
        if max(n_l, n_r) > cp or max(m_l, m_r) > cl:
            raise RuntimeError("frame exceeds the block capacities")
        block[:] = 0
        hd = FrameHeader.from_buffer(block)
        hd.n_l, hd.n_r, hd.m_l, hd.m_r, hd.cap_points, hd.cap_lines, hd.status = n_l, n_r, m_l, m_r, cp, cl, 0
        self._arr(block, o.kps_l, KEYPOINT, cp)[:n_l] = f.kps; self._arr(block, o.desc_l, np.uint8, cp * 32, (cp, 32))[:n_l] = f.desc
        self._arr(block, o.kps_r, KEYPOINT, cp)[:n_r] = f.kps_r; self._arr(block, o.desc_r, np.uint8, cp * 32, (cp, 32))[:n_r] = f.desc_r
        self._arr(block, o.u_right, np.float32, cp)[:n_l] = f.u_right; self._arr(block, o.depth, np.float32, cp)[:n_l] = f.depth
        if m_l or m_r:
            self._arr(block, o.kls_l, KEYLINE, cl)[:m_l] = f.kls; self._arr(block, o.ldesc_l, np.uint8, cl * 32, (cl, 32))[:m_l] = f.ldesc
            self._arr(block, o.kls_r, KEYLINE, cl)[:m_r] = f.kls_r; self._arr(block, o.ldesc_r, np.uint8, cl * 32, (cl, 32))[:m_r] = f.ldesc_r
            self._arr(block, o.lmatch, np.int32, cl)[:m_l] = f.line_matches
            self._arr(block, o.ldisp, np.float32, cl * 2, (cl, 2))[:m_l] = f.line_disp
            self._arr(block, o.lle, np.float64, cl * 3, (cl, 3))[:m_l] = f.line_le
        return block
