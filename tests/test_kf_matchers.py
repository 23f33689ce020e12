"""the remaining ORBmatcher overloads -- Fuse (src/ORBmatcher.cc:827-1102), SearchByProjection(KF, Scw) (:292-405),
SearchBySim3 (:1104-1328), SearchByProjection(Frame, KF) (:1620-1747), SearchForTriangulation (:659-825), SearchByBoW(KF, KF) (:524-657).
CPU: the oracle against an independent numpy restatement.  GPU: the product (C-ABI) against the oracle, bit for bit.
(tests/test_oracle_vs_ref.py pins the oracle to the reference's own function text.)"""
import numpy as np
import pytest
from orc import oracle
import kf_search as K

f32 = np.float32


def popcount_rows(a, b):
    return np.unpackbits(np.bitwise_xor(a, b), axis=-1).sum(-1)


def window_search_numpy(kf, q, max_dist, blocked=None, sequential=False, chi2=None):
    """Independent restatement: explicit 64 x 48 grid, cell-major enumeration (src/KeyFrame.cc:747-786), first minimum."""
    kps, desc, cam = kf["kps"], kf["desc"], kf["cam"]
    n = len(kps)
    inv_w = f32(f32(64) / f32(cam.max_x - cam.min_x)); inv_h = f32(f32(48) / f32(cam.max_y - cam.min_y))
    px = np.round((kps["x"] - f32(cam.min_x)) * inv_w).astype(int); py = np.round((kps["y"] - f32(cam.min_y)) * inv_h).astype(int)
    # C roundf rounds half away from zero, numpy half to even: positions are non-negative here, fix the .5 cases
    fx_ = (kps["x"] - f32(cam.min_x)) * inv_w; fy_ = (kps["y"] - f32(cam.min_y)) * inv_h
    px = np.floor(fx_ + f32(0.5)).astype(int); py = np.floor(fy_ + f32(0.5)).astype(int)
    blocked = np.zeros(n, bool) if blocked is None else blocked.astype(bool).copy()
    bi = np.full(len(q["u"]), -1, np.int32); bd = np.full(len(q["u"]), 256, np.int32)
    for i in range(len(q["u"])):
        u, v, r = q["u"][i], q["v"][i], q["radius"][i]
        c0 = max(0, int(np.floor(f32(f32(f32(u - f32(cam.min_x)) - r) * inv_w)))); c1 = min(63, int(np.ceil(f32(f32(f32(u - f32(cam.min_x)) + r) * inv_w))))
        r0 = max(0, int(np.floor(f32(f32(f32(v - f32(cam.min_y)) - r) * inv_h)))); r1 = min(47, int(np.ceil(f32(f32(f32(v - f32(cam.min_y)) + r) * inv_h))))
        if c0 >= 64 or c1 < 0 or r0 >= 48 or r1 < 0:
            continue
        m = (px >= c0) & (px <= c1) & (py >= r0) & (py <= r1) & (px < 64) & (py < 48) & (px >= 0) & (py >= 0)
        m &= (np.abs(kps["x"] - u) < r) & (np.abs(kps["y"] - v) < r)
        m &= (kps["octave"] >= q["min_level"][i]) & (kps["octave"] <= q["max_level"][i]) & ~blocked
        if chi2 is not None:
            ur_kf, inv_sig, ur_q = chi2
            ex = (u - kps["x"]).astype(np.float32); ey = (v - kps["y"]).astype(np.float32); er = (ur_q[i] - ur_kf).astype(np.float32)
            e_m = (ex * ex + ey * ey).astype(np.float32); e_s = (e_m + er * er).astype(np.float32)
            w = inv_sig[kps["octave"]]
            m &= np.where(ur_kf >= 0, (e_s * w).astype(np.float32).astype(np.float64) <= 7.8, (e_m * w).astype(np.float32).astype(np.float64) <= 5.99)
        idx = np.nonzero(m)[0]
        if len(idx) == 0:
            continue
        d = popcount_rows(desc[idx], q["qdesc"][i][None, :])
        order = np.lexsort((idx, py[idx], px[idx], d))          # first minimum in (ix, iy, index) enumeration order
        j, dj = idx[order[0]], int(d[order[0]])
        if dj <= max_dist:
            bi[i] = j; bd[i] = dj
            if sequential:
                blocked[j] = True
    return bi, bd


CASES = [dict(max_dist=50, sequential=False, chi2=False, blocked=False),        # Fuse(KF, Scw, ..), SearchBySim3 uses 100
         dict(max_dist=100, sequential=False, chi2=False, blocked=False),
         dict(max_dist=50, sequential=False, chi2=True, blocked=False),         # Fuse(KF, MPs, th)
         dict(max_dist=50, sequential=True, chi2=False, blocked=True),          # SearchByProjection(KF, Scw, ..)
         dict(max_dist=100, sequential=True, chi2=False, blocked=True),         # SearchByProjection(Frame, KF, found, 10, 100)
         dict(max_dist=64, sequential=True, chi2=False, blocked=False)]


def run_window(api, kf, q, case, inv_sig):
    blocked = (np.arange(len(kf["kps"])) % 5 == 0).astype(np.uint8) if case["blocked"] else None
    chi2 = (kf["u_right"], inv_sig, q["ur"]) if case["chi2"] else None
    return api.window_search(kf["kps"], kf["desc"], kf["cam"], q["u"], q["v"], q["radius"], q["min_level"], q["max_level"], q["qdesc"],
                             case["max_dist"], blocked, case["sequential"], chi2), blocked, chi2


@pytest.mark.parametrize("case", CASES)
def test_oracle_window_search_matches_numpy(case):
    kf = K.make_keyframe(5, n=900)
    q = K.random_queries(kf, 260, 6)
    _, _, inv_sig, _ = K.scale_tables()
    (bi, bd), blocked, chi2 = run_window(oracle(), kf, q, case, inv_sig)
    bi2, bd2 = window_search_numpy(kf, q, case["max_dist"], blocked, case["sequential"], chi2)
    assert np.array_equal(bi, bi2) and np.array_equal(bd, bd2)
    assert (bi >= 0).sum() > (20 if case["chi2"] else 40)


def test_oracle_window_search_edge_cases():
    o = oracle()
    kf = K.make_keyframe(7, n=50)
    q = K.random_queries(kf, 0, 1)
    bi, bd = o.window_search(kf["kps"], kf["desc"], kf["cam"], q["u"], q["v"], q["radius"], q["min_level"], q["max_level"], q["qdesc"], 50)
    assert len(bi) == 0
    q = K.random_queries(kf, 20, 2)
    q["u"][:5] = [-500, 5000, 10, 10, 1279.5]; q["v"][:5] = [10, 10, -900, 9000, 719.5]          # windows outside / at the border of the grid
    bi, bd = o.window_search(kf["kps"], kf["desc"], kf["cam"], q["u"], q["v"], q["radius"], q["min_level"], q["max_level"], q["qdesc"], 256)
    bi2, bd2 = window_search_numpy(kf, q, 256)
    assert np.array_equal(bi, bi2) and np.array_equal(bd, bd2) and (bi[:4] == -1).all()
    empty = dict(kps=kf["kps"][:0], desc=kf["desc"][:0], cam=kf["cam"])
    bi, bd = o.window_search(empty["kps"], empty["desc"], empty["cam"], q["u"], q["v"], q["radius"], q["min_level"], q["max_level"], q["qdesc"], 256)
    assert (bi == -1).all() and (bd == 256).all()


def triangulation_numpy(kf1, kf2, F12, ex, ey, only_stereo, check_orientation):
    m12 = np.full(len(kf1["kps"]), -1, np.int32)
    n1n, b1, i1 = kf1["fv"]; n2n, b2, i2 = kf2["fv"]
    F = np.asarray(F12, np.float32).reshape(3, 3)
    sf, sig = kf2["scale_factors"], kf2["level_sigma2"]
    for a, node in enumerate(n1n):
        pos = np.nonzero(n2n == node)[0]
        if len(pos) == 0:
            continue
        c2 = i2[b2[pos[0]]:b2[pos[0] + 1]]
        for idx1 in i1[b1[a]:b1[a + 1]]:
            s1 = kf1["u_right"][idx1] >= 0
            if kf1["skip"][idx1] or (only_stereo and not s1):
                continue
            k1 = kf1["kps"][idx1]
            la = f32(f32(f32(k1["x"] * F[0, 0]) + f32(k1["y"] * F[1, 0])) + F[2, 0])
            lb = f32(f32(f32(k1["x"] * F[0, 1]) + f32(k1["y"] * F[1, 1])) + F[2, 1])
            lc = f32(f32(f32(k1["x"] * F[0, 2]) + f32(k1["y"] * F[1, 2])) + F[2, 2])
            den = f32(f32(la * la) + f32(lb * lb))
            best, bj = 51, -1
            for idx2 in c2:
                s2 = kf2["u_right"][idx2] >= 0
                if kf2["skip"][idx2] or (only_stereo and not s2):
                    continue
                d = int(popcount_rows(kf1["desc"][idx1], kf2["desc"][idx2]))
                if d > 50 or d > best:
                    continue
                k2 = kf2["kps"][idx2]
                if not s1 and not s2:
                    dx = f32(ex - k2["x"]); dy = f32(ey - k2["y"])
                    if f32(f32(dx * dx) + f32(dy * dy)) < f32(f32(100) * sf[k2["octave"]]):
                        continue
                if den == 0:
                    continue
                num = f32(f32(f32(la * k2["x"]) + f32(lb * k2["y"])) + lc)
                dsqr = f32(f32(num * num) / den)
                if float(dsqr) < 3.84 * float(sig[k2["octave"]]):
                    best, bj = d, idx2
            m12[idx1] = bj
    n = int((m12 >= 0).sum())
    if check_orientation:
        hist = [[] for _ in range(30)]
        for i in np.nonzero(m12 >= 0)[0]:
            rot = f32(kf1["kps"]["angle"][i] - kf2["kps"]["angle"][m12[i]])
            if rot < 0:
                rot = f32(rot + f32(360))
            b = int(np.floor(float(f32(rot * f32(f32(1.0) / f32(30)))) + 0.5))
            hist[0 if b == 30 else b].append(i)
        sizes = [len(h) for h in hist]
        mx = [0, 0, 0]; ind = [-1, -1, -1]
        for i, s in enumerate(sizes):
            if s > mx[0]:
                mx = [s, mx[0], mx[1]]; ind = [i, ind[0], ind[1]]
            elif s > mx[1]:
                mx = [mx[0], s, mx[1]]; ind = [ind[0], i, ind[1]]
            elif s > mx[2]:
                mx[2] = s; ind[2] = i
        if mx[1] < f32(0.1) * f32(mx[0]):
            ind[1] = ind[2] = -1
        elif mx[2] < f32(0.1) * f32(mx[0]):
            ind[2] = -1
        for i, h in enumerate(hist):
            if i not in ind:
                for e in h:
                    m12[e] = -1; n -= 1
    return m12, n


@pytest.mark.parametrize("only_stereo,ori", [(False, False), (True, False), (False, True)])
def test_oracle_triangulation_matches_numpy(only_stereo, ori):
    kf1, kf2, F12, ex, ey = K.make_stereo_pair_keyframes(3, n=500)
    m, n = oracle().search_for_triangulation(kf1, kf2, F12, ex, ey, only_stereo, ori)
    m2, n2 = triangulation_numpy(kf1, kf2, F12, ex, ey, only_stereo, ori)
    assert np.array_equal(m, m2) and n == n2
    assert n > 60


def bow_kf_numpy(kf1, kf2, has1, has2, nn_ratio):
    m12 = np.full(len(kf1["kps"]), -1, np.int32)
    matched2 = np.zeros(len(kf2["kps"]), bool)
    n1n, b1, i1 = kf1["fv"]; n2n, b2, i2 = kf2["fv"]
    for a, node in enumerate(n1n):
        pos = np.nonzero(n2n == node)[0]
        if len(pos) == 0:
            continue
        c2 = i2[b2[pos[0]]:b2[pos[0] + 1]]
        for idx1 in i1[b1[a]:b1[a + 1]]:
            if not has1[idx1]:
                continue
            cand = [j for j in c2 if has2[j] and not matched2[j]]
            if not cand:
                continue
            d = popcount_rows(kf2["desc"][cand], kf1["desc"][idx1][None, :])
            o = np.argsort(d, kind="stable")
            d1 = int(d[o[0]]); d2 = int(d[o[1]]) if len(o) > 1 else 256
            if d1 < 50 and f32(d1) < f32(f32(nn_ratio) * f32(d2)):
                m12[idx1] = cand[o[0]]; matched2[cand[o[0]]] = True
    return m12, int((m12 >= 0).sum())


def test_oracle_bow_kf_matches_numpy():
    kf1, kf2, *_ = K.make_stereo_pair_keyframes(9, n=500, pool_noise=30)
    has1 = 1 - kf1["skip"]; has2 = 1 - kf2["skip"]
    m, n = oracle().search_by_bow_kf(kf1["desc"], kf1["kps"], has1, kf1["fv"], kf2["desc"], kf2["kps"], has2, kf2["fv"], 0.75, False)
    m2, n2 = bow_kf_numpy(kf1, kf2, has1, has2, 0.75)
    assert np.array_equal(m, m2) and n == n2 and n > 50
    assert len(set(m[m >= 0])) == n                                   # a feature of KF2 is matched at most once


# ---- GPU: product vs oracle ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("seed,n,nq", [(1, 2000, 1500), (2, 4000, 3000), (3, 300, 100)])
def test_gpu_window_search(case, seed, n, nq):
    import orb_line_slam_b200 as olf
    kf = K.make_keyframe(seed, n=n)
    q = K.random_queries(kf, nq, seed + 100, max_radius=45.0 if n < 3000 else 25.0)
    _, _, inv_sig, _ = K.scale_tables()
    (gi, gd), _, _ = run_window(olf.api(0), kf, q, case, inv_sig)
    (oi, od), _, _ = run_window(oracle(), kf, q, case, inv_sig)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    assert (gi >= 0).sum() > nq // 10


def dense_keyframe(seed, n):
    """Every keypoint inside a 260 x 200 patch, two octaves, two descriptor families: a 100-pixel window holds hundreds of candidates within
    max_dist -- far beyond the 128 slots of the first device pass (the reference's vectors have no such limit)."""
    kf = K.make_keyframe(seed, n=n, pool=2)
    rng = np.random.RandomState(seed + 7)
    kf["kps"]["x"] = (500 + rng.rand(n) * 260).astype(np.float32); kf["kps"]["y"] = (300 + rng.rand(n) * 200).astype(np.float32)
    kf["kps"]["octave"] = rng.randint(0, 2, n)
    kf["u_right"] = np.where(kf["u_right"] >= 0, kf["kps"]["x"] - (5 + 40 * rng.rand(n)).astype(np.float32), np.float32(-1)).astype(np.float32)
    return kf


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_window_search_dense_windows(case):
    import orb_line_slam_b200 as olf
    kf = dense_keyframe(11, 1500)
    q = K.random_queries(kf, 700, 111, max_radius=120.0, pool=2)
    q["min_level"][:] = 0; q["max_level"][:] = 1
    q["radius"][::7] = 4.0                                                        # and a few ordinary windows among them
    inside = (np.abs(kf["kps"]["x"][None, :] - q["u"][:40, None]) < q["radius"][:40, None]) & (np.abs(kf["kps"]["y"][None, :] - q["v"][:40, None]) < q["radius"][:40, None])
    near = np.stack([popcount_rows(kf["desc"], q["qdesc"][i][None, :]) <= case["max_dist"] for i in range(40)])
    assert (inside & near).sum(1).max() > 300                                     # the case does overflow the first pass
    _, _, inv_sig, _ = K.scale_tables()
    (gi, gd), _, _ = run_window(olf.api(0), kf, q, case, inv_sig)
    (oi, od), _, _ = run_window(oracle(), kf, q, case, inv_sig)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od) and (gi >= 0).sum() > 100


@pytest.mark.gpu
@pytest.mark.parametrize("mode,th,max_dist", [("fuse", 3.0, 50), ("fuse_sim3", 4.0, 50), ("sbp_kf", 10.0, 50), ("reloc", 10.0, 100), ("reloc", 3.0, 64)])
def test_gpu_projection_overloads(mode, th, max_dist):
    """The real call pattern: prologue of the overload (tests/kf_search.py) -> window search with the overload's gate / blocking rule."""
    import orb_line_slam_b200 as olf
    kf = K.make_keyframe(11, n=2000)
    P, Rcw, tcw, Ow = K.make_points(kf, 12, n=1200)
    s, sig, inv_sig, log_sf = K.scale_tables()
    ok, u, v, ur, lvl, radius = K.project_points(mode, P, Rcw, tcw, Ow, kf["cam"], s, log_sf, th)
    assert ok.sum() > 500
    sel = np.nonzero(ok)[0]
    maxl = lvl[sel] + (1 if mode == "reloc" else 0)
    args = (kf["kps"], kf["desc"], kf["cam"], u[sel], v[sel], radius[sel], lvl[sel] - 1, maxl, P["desc"][sel], max_dist)
    blocked = (np.arange(len(kf["kps"])) % 7 == 0).astype(np.uint8) if mode in ("sbp_kf", "reloc") else None
    kw = dict(blocked=blocked, sequential=mode in ("sbp_kf", "reloc"), chi2=(kf["u_right"], inv_sig, ur[sel]) if mode == "fuse" else None)
    gi, gd = olf.api(0).window_search(*args, **kw)
    oi, od = oracle().window_search(*args, **kw)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od) and (gi >= 0).sum() > 50


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,only_stereo,ori", [(1, 2000, False, False), (2, 2000, True, False), (3, 3000, False, True), (4, 200, False, True)])
def test_gpu_search_for_triangulation(seed, n, only_stereo, ori):
    import orb_line_slam_b200 as olf
    kf1, kf2, F12, ex, ey = K.make_stereo_pair_keyframes(seed, n=n)
    gm, gn = olf.api(0).search_for_triangulation(kf1, kf2, F12, ex, ey, only_stereo, ori)
    om, on = oracle().search_for_triangulation(kf1, kf2, F12, ex, ey, only_stereo, ori)
    assert np.array_equal(gm, om) and gn == on and gn > n // 20


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,ratio,ori", [(1, 2000, 0.75, True), (2, 2000, 0.6, False), (3, 150, 0.9, True)])
def test_gpu_search_by_bow_kf(seed, n, ratio, ori):
    import orb_line_slam_b200 as olf
    kf1, kf2, *_ = K.make_stereo_pair_keyframes(seed, n=n, pool_noise=30)
    has1 = 1 - kf1["skip"]; has2 = 1 - kf2["skip"]
    args = (kf1["desc"], kf1["kps"], has1, kf1["fv"], kf2["desc"], kf2["kps"], has2, kf2["fv"], ratio, ori)
    gm, gn = olf.api(0).search_by_bow_kf(*args)
    om, on = oracle().search_by_bow_kf(*args)
    assert np.array_equal(gm, om) and gn == on and gn > n // 20


def _empty_like(kf):
    e = dict(kf)
    e["kps"] = kf["kps"][:0]; e["desc"] = kf["desc"][:0]; e["skip"] = kf["skip"][:0]; e["u_right"] = kf["u_right"][:0]
    e["fv"] = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    return e


def _edge_cases(api):
    """Empty key frames, disjoint vocabulary nodes, every feature skipped: what the reference's loops simply fall through."""
    kf1, kf2, F12, ex, ey = K.make_stereo_pair_keyframes(5, n=300)
    out = []
    out.append(api.search_for_triangulation(_empty_like(kf1), kf2, F12, ex, ey))
    out.append(api.search_for_triangulation(kf1, _empty_like(kf2), F12, ex, ey))
    disjoint = dict(kf2); disjoint["fv"] = (kf2["fv"][0] + 1000, kf2["fv"][1], kf2["fv"][2])     # no node in common
    out.append(api.search_for_triangulation(kf1, disjoint, F12, ex, ey))
    allskip = dict(kf1); allskip["skip"] = np.ones_like(kf1["skip"])
    out.append(api.search_for_triangulation(allskip, kf2, F12, ex, ey, False, True))
    zeroF = np.zeros(9, np.float32)                                                               # den == 0: CheckDistEpipolarLine returns false
    out.append(api.search_for_triangulation(kf1, kf2, zeroF, ex, ey))
    has1, has2 = 1 - kf1["skip"], 1 - kf2["skip"]
    out.append(api.search_by_bow_kf(kf1["desc"], kf1["kps"], has1, kf1["fv"], kf2["desc"], kf2["kps"], np.zeros_like(has2), kf2["fv"]))
    out.append(api.search_by_bow_kf(kf1["desc"], kf1["kps"], has1, kf1["fv"], kf2["desc"], kf2["kps"], has2, disjoint["fv"]))
    e1 = _empty_like(kf1)
    out.append(api.search_by_bow_kf(e1["desc"], e1["kps"], e1["skip"], e1["fv"], kf2["desc"], kf2["kps"], has2, kf2["fv"]))
    return out


def test_oracle_edge_cases():
    for m, n in _edge_cases(oracle()):
        assert n == 0 and (m == -1).all()


@pytest.mark.gpu
def test_gpu_edge_cases():
    import orb_line_slam_b200 as olf
    api = olf.api(0)
    for (gm, gn), (om, on) in zip(_edge_cases(api), _edge_cases(oracle())):
        assert gn == on == 0 and np.array_equal(gm, om)
    kf = K.make_keyframe(7, n=50)
    q = K.random_queries(kf, 20, 2)
    q["u"][:5] = [-500, 5000, 10, 10, 1279.5]; q["v"][:5] = [10, 10, -900, 9000, 719.5]
    args = (kf["kps"], kf["desc"], kf["cam"], q["u"], q["v"], q["radius"], q["min_level"], q["max_level"], q["qdesc"], 256)
    gi, gd = api.window_search(*args); oi, od = oracle().window_search(*args)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    q0 = K.random_queries(kf, 0, 1)
    gi, gd = api.window_search(kf["kps"], kf["desc"], kf["cam"], q0["u"], q0["v"], q0["radius"], q0["min_level"], q0["max_level"], q0["qdesc"], 50)
    assert len(gi) == 0
    gi, gd = api.window_search(kf["kps"][:0], kf["desc"][:0], kf["cam"], q["u"], q["v"], q["radius"], q["min_level"], q["max_level"], q["qdesc"], 256)
    assert (gi == -1).all() and (gd == 256).all()
    # a window with more than 128 admissible candidates (here: thousands) is neither truncated nor refused
    dense = K.make_keyframe(8, n=4000, w=320, h=240)
    dense["cam"] = K.Camera(300.0, 300.0, 160.0, 120.0, 40.0, 0.0, 320.0, 0.0, 240.0)
    args = (dense["kps"], dense["desc"], dense["cam"], np.array([160.0], np.float32), np.array([120.0], np.float32), np.array([100.0], np.float32),
            np.array([0], np.int32), np.array([7], np.int32), dense["desc"][:1], 256)
    gi, gd = api.window_search(*args); oi, od = oracle().window_search(*args)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od) and gd[0] == 0


# ---- ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:407-522) ---------------------------------------------------------
def init_case(seed, n=1500, shift=(6.0, -3.0)):
    """Two monocular frames: F2 = F1's keypoints moved by a small flow + noise, descriptors of a small pool (take-overs and ties are frequent)."""
    kf = K.make_keyframe(seed, n=n, pool=200)
    rng = np.random.RandomState(seed + 1)
    k1 = kf["kps"].copy(); k1["octave"] = np.where(rng.rand(n) < 0.6, 0, k1["octave"])
    k2 = k1.copy()
    k2["x"] = np.clip(k1["x"] + shift[0] + rng.randn(n) * 3, 0, 1279).astype(np.float32); k2["y"] = np.clip(k1["y"] + shift[1] + rng.randn(n) * 3, 0, 719).astype(np.float32)
    k2["angle"] = ((k1["angle"] + rng.randn(n) * 4) % 360).astype(np.float32)
    d2 = kf["desc"].copy()
    for i in range(n):
        for b in rng.randint(0, 256, rng.randint(0, 25)):
            d2[i, b >> 3] ^= np.uint8(1 << (b & 7))
    perm = rng.permutation(n)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)                 # vbPrevMatched starts as F1's keypoints (src/Tracking.cc:731-733)
    return k1, kf["desc"], k2[perm], d2[perm], kf["cam"], prev


def init_numpy(k1, d1, k2, d2, cam, prev, window, ratio, ori):
    n1, n2 = len(k1), len(k2)
    m12 = np.full(n1, -1, np.int32); m21 = np.full(n2, -1, np.int64); md = np.full(n2, 2 ** 31 - 1, np.int64)
    kf2 = dict(kps=k2, desc=d2, cam=cam)
    hist = [[] for _ in range(30)]; nm = 0
    inv_w = f32(f32(64) / f32(cam.max_x - cam.min_x)); inv_h = f32(f32(48) / f32(cam.max_y - cam.min_y))
    px = np.floor((k2["x"] - f32(cam.min_x)) * inv_w + f32(0.5)).astype(int); py = np.floor((k2["y"] - f32(cam.min_y)) * inv_h + f32(0.5)).astype(int)
    for i1 in range(n1):
        if k1["octave"][i1] > 0:
            continue
        u, v, r = prev[i1, 0], prev[i1, 1], f32(window)
        c0 = max(0, int(np.floor(f32(f32(f32(u - f32(cam.min_x)) - r) * inv_w)))); c1 = min(63, int(np.ceil(f32(f32(f32(u - f32(cam.min_x)) + r) * inv_w))))
        r0 = max(0, int(np.floor(f32(f32(f32(v - f32(cam.min_y)) - r) * inv_h)))); r1 = min(47, int(np.ceil(f32(f32(f32(v - f32(cam.min_y)) + r) * inv_h))))
        if c0 >= 64 or c1 < 0 or r0 >= 48 or r1 < 0:
            continue
        m = (px >= c0) & (px <= c1) & (py >= r0) & (py <= r1) & (px < 64) & (py < 48) & (k2["octave"] == 0)
        m &= (np.abs(k2["x"] - u) < r) & (np.abs(k2["y"] - v) < r)
        idx = np.nonzero(m)[0]
        if len(idx) == 0:
            continue
        d = popcount_rows(d2[idx], d1[i1][None, :])
        ok = md[idx] > d
        idx, d = idx[ok], d[ok]
        if len(idx) == 0:
            continue
        order = np.lexsort((idx, py[idx], px[idx], d))
        best, bd = idx[order[0]], int(d[order[0]]); bd2 = int(d[order[1]]) if len(order) > 1 else 2 ** 31 - 1
        if bd <= 50 and f32(bd) < f32(f32(bd2) * f32(ratio)):
            if m21[best] >= 0:
                m12[m21[best]] = -1; nm -= 1
            m12[i1] = best; m21[best] = i1; md[best] = bd; nm += 1
            rot = f32(k1["angle"][i1] - k2["angle"][best])
            if rot < 0:
                rot = f32(rot + f32(360))
            b = int(np.floor(float(f32(rot * f32(f32(1.0) / f32(30)))) + 0.5))
            hist[0 if b == 30 else b].append(i1)
    if ori:
        sizes = [len(h) for h in hist]; mx = [0, 0, 0]; ind = [-1, -1, -1]
        for i, s_ in enumerate(sizes):
            if s_ > mx[0]:
                mx = [s_, mx[0], mx[1]]; ind = [i, ind[0], ind[1]]
            elif s_ > mx[1]:
                mx = [mx[0], s_, mx[1]]; ind = [ind[0], i, ind[1]]
            elif s_ > mx[2]:
                mx[2] = s_; ind[2] = i
        if mx[1] < f32(0.1) * f32(mx[0]):
            ind[1] = ind[2] = -1
        elif mx[2] < f32(0.1) * f32(mx[0]):
            ind[2] = -1
        for i, h in enumerate(hist):
            if i not in ind:
                for e in h:
                    if m12[e] >= 0:
                        m12[e] = -1; nm -= 1
    out = prev.copy()
    for i1 in np.nonzero(m12 >= 0)[0]:
        out[i1] = (k2["x"][m12[i1]], k2["y"][m12[i1]])
    return m12, nm, out


@pytest.mark.parametrize("seed,window,ratio,ori", [(1, 100, 0.9, True), (2, 40, 0.9, False), (3, 100, 0.7, True)])
def test_oracle_search_for_initialization_matches_numpy(seed, window, ratio, ori):
    k1, d1, k2, d2, cam, prev = init_case(seed, n=500)
    m, n, pm = oracle().search_for_initialization(k1, d1, k2, d2, cam, prev, window, ratio, ori)
    m2, n2, pm2 = init_numpy(k1, d1, k2, d2, cam, prev, window, ratio, ori)
    assert np.array_equal(m, m2) and n == n2 and np.array_equal(pm, pm2) and n > 30


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,window,ratio,ori", [(1, 2000, 100, 0.9, True), (2, 2000, 30, 0.9, False), (3, 1000, 100, 0.7, True), (4, 60, 100, 0.9, True)])
def test_gpu_search_for_initialization(seed, n, window, ratio, ori):
    import orb_line_slam_b200 as olf
    k1, d1, k2, d2, cam, prev = init_case(seed, n=n)
    gm, gn, gp = olf.api(0).search_for_initialization(k1, d1, k2, d2, cam, prev, window, ratio, ori)
    om, on, op = oracle().search_for_initialization(k1, d1, k2, d2, cam, prev, window, ratio, ori)
    assert np.array_equal(gm, om) and gn == on and np.array_equal(gp, op) and gn > n // 30
    e = olf.api(0).search_for_initialization(k1[:0], d1[:0], k2, d2, cam, prev[:0], window, ratio, ori)
    assert len(e[0]) == 0 and e[1] == 0
    gm, gn, gp = olf.api(0).search_for_initialization(k1, d1, k2[:0], d2[:0], cam, prev, window, ratio, ori)
    assert gn == 0 and (gm == -1).all() and np.array_equal(gp, prev)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,ratio,ori", [(5, 0.9, True), (6, 0.8, False)])
def test_gpu_search_for_initialization_dense_windows(seed, ratio, ori):
    """4000 keypoints squeezed into a quarter of the image: the 100-pixel window of SearchForInitialization (src/ORBmatcher.cc:407) holds many
    hundreds of level-0 keypoints, the lists go through the exact-capacity second pass."""
    import orb_line_slam_b200 as olf
    k1, d1, k2, d2, cam, prev = init_case(seed, n=4000)
    for k in (k1, k2):
        k["x"] = (400 + k["x"] * 0.25).astype(np.float32); k["y"] = (250 + k["y"] * 0.3).astype(np.float32)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    lvl0 = k2[k2["octave"] == 0]
    assert ((np.abs(lvl0["x"] - prev[0, 0]) < 100) & (np.abs(lvl0["y"] - prev[0, 1]) < 100)).sum() > 300
    gm, gn, gp = olf.api(0).search_for_initialization(k1, d1, k2, d2, cam, prev, 100, ratio, ori)
    om, on, op = oracle().search_for_initialization(k1, d1, k2, d2, cam, prev, 100, ratio, ori)
    assert np.array_equal(gm, om) and gn == on and np.array_equal(gp, op) and gn > 20
