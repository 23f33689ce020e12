"""Bag of words (Frame::ComputeBoW -> DBoW2 transform, ORBmatcher::SearchByBoW).
CPU: the oracle against an independent numpy restatement of the tree descent and of the std::map bookkeeping.
GPU: the product against the oracle, bit for bit (word ids, weights, node ids, BowVector doubles, match arrays)."""
import numpy as np
import pytest
from orc import oracle
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS, random_image
from bow_util import make_vocabulary, transform_numpy


def orb_frames(api, n=2, nfeatures=1000):
    sc = Scene("euroc", 3)
    h = api.orb_create(nfeatures)
    out = [api.orb_extract(h, sc.stereo(f)[0]) for f in range(n)]
    api.orb_destroy(h)
    return out


def test_oracle_transform_matches_numpy_restatement():
    o = oracle()
    (k0, d0), = orb_frames(o, 1, 300)
    tree = make_vocabulary(10, 3, seed=1, seed_desc=d0)
    v = o.vocab_create(tree)
    for levelsup in (0, 1, 2, 3, 4):
        w, val, nd = o.bow_transform(v, d0, levelsup)
        w2, val2, nd2 = transform_numpy(tree, d0, levelsup)
        assert np.array_equal(w, w2) and np.array_equal(val, val2) and np.array_equal(nd, nd2), levelsup
    assert len(np.unique(w)) > 50                                  # the descriptors spread over the tree
    # std::map bookkeeping: words ascending, sums in feature order, L1 norm; nodes ascending, indices in feature order
    w, val, nd = o.bow_transform(v, d0, 1)
    bw, bv, fn, fb, fi = o.bow_assemble(w, val, nd)
    acc = {}
    for i in range(len(w)):
        if val[i] > 0:
            acc[int(w[i])] = acc.get(int(w[i]), 0.0) + val[i]
    words = sorted(acc)
    norm = 0.0
    for x in words:
        norm += abs(acc[x])
    assert list(bw) == words and np.array_equal(bv, np.array([acc[x] / norm for x in words]))
    assert abs(bv.sum() - 1.0) < 1e-12
    nodes = sorted({int(x) for x, vv in zip(nd, val) if vv > 0})
    assert list(fn) == nodes and fb[-1] == (val > 0).sum()
    for j, node in enumerate(nodes):
        assert list(fi[fb[j]:fb[j + 1]]) == [i for i in range(len(nd)) if nd[i] == node and val[i] > 0]
    o.vocab_destroy(v)


def bow_case(api, tree, frames, levelsup, nn_ratio, check_orientation, has_mode):
    v = api.vocab_create(tree)
    (kk, dk), (kf_, df) = frames
    tk, tf = api.bow_transform(v, dk, levelsup), api.bow_transform(v, df, levelsup)
    ak, af = api.bow_assemble(*tk), api.bow_assemble(*tf)
    has = np.ones(len(dk), np.uint8) if has_mode == "all" else (np.arange(len(dk)) % 3 != 0).astype(np.uint8)
    m, n = api.search_by_bow(dk, kk, has, ak[2:], df, kf_, af[2:], nn_ratio, check_orientation)
    api.vocab_destroy(v)
    return tk, tf, ak, af, m, n


@pytest.mark.gpu
@pytest.mark.parametrize("L,levelsup,nn_ratio,ori,has_mode", [(3, 2, 0.7, True, "all"), (3, 1, 0.9, False, "alt"), (4, 4, 0.7, True, "alt"),
                                                              (4, 2, 0.6, True, "all"), (2, 4, 0.75, True, "all")])
def test_bow_parity(L, levelsup, nn_ratio, ori, has_mode):
    import orb_line_slam_b200 as olf
    o, g = oracle(), olf.api(0)
    frames = orb_frames(o, 2, 1000)
    tree = make_vocabulary(10, L, seed=L, seed_desc=frames[0][1])
    ro = bow_case(o, tree, frames, levelsup, nn_ratio, ori, has_mode)
    rg = bow_case(g, tree, frames, levelsup, nn_ratio, ori, has_mode)
    for a, b in zip(ro[0] + ro[1], rg[0] + rg[1]):
        assert np.array_equal(a, b)                                # word ids, weights (bits), node ids
    for a, b in zip(ro[2] + ro[3], rg[2] + rg[3]):
        assert np.array_equal(a, b)                                # BowVector / FeatureVector
    assert ro[5] == rg[5] and np.array_equal(ro[4], rg[4])
    if levelsup < L and nn_ratio >= 0.7:
        assert ro[5] > 20                                          # the case is not vacuous


@pytest.mark.gpu
def test_bow_edge_cases():
    import orb_line_slam_b200 as olf
    o, g = oracle(), olf.api(0)
    tree = make_vocabulary(4, 2, seed=9)
    rng = np.random.RandomState(3)
    for api in (o, g):
        v = api.vocab_create(tree)
        w, val, nd = api.bow_transform(v, np.zeros((0, 32), np.uint8), 1)
        assert len(w) == 0
        d = rng.randint(0, 256, (5, 32)).astype(np.uint8)
        t = api.bow_transform(v, d, 1)
        a = api.bow_assemble(*t)
        kp = np.zeros(5, dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
        m, n = api.search_by_bow(d, kp, np.ones(5, np.uint8), a[2:], d, kp, a[2:], 0.9, True)     # a frame against itself
        assert n == int((m >= 0).sum())
        api.vocab_destroy(v)
    # identical descriptors at distance 0 with a unique nearest neighbour match themselves


def test_product_bow_assemble_host_logic_matches_oracle():
    """olf_bow_assemble is host code inside libolf.so (the std::map bookkeeping of DBoW2's transform): no GPU needed."""
    import orb_line_slam_b200 as olf
    o, g = oracle(), olf.api(0)
    rng = np.random.RandomState(11)
    for n in (0, 1, 7, 500, 3000):
        w = rng.randint(0, max(1, n // 3 + 1), n).astype(np.int32)
        v = np.where(rng.rand(n) < 0.1, 0.0, rng.rand(n) * 5).astype(np.float64)          # some stopped words
        nd = rng.randint(1, 40, n).astype(np.int32)
        ro, rg = o.bow_assemble(w, v, nd), g.bow_assemble(w, v, nd)
        for a, b in zip(ro, rg):
            assert np.array_equal(a, b), n
