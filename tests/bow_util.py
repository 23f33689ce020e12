"""Synthetic DBoW2-shaped vocabulary for the bag-of-words tests (the reference's ORBvoc is a 145 MB text file that cannot
travel): a complete k-ary tree of depth L, node descriptors = noisy copies of their parent's (so that real ORB
descriptors spread over the tree instead of all falling into one branch), leaf weights = idf-like positive values with a
few stopped words (weight 0), children listed in a shuffled order (m_nodes[i].children need not be ascending)."""
import numpy as np


def make_vocabulary(k=10, L=3, seed=0, seed_desc=None, stop_fraction=0.02):
    rng = np.random.RandomState(seed)
    n = (k ** (L + 1) - 1) // (k - 1)
    desc = np.zeros((n, 32), np.uint8)
    child_begin = np.zeros(n, np.int32); child_count = np.zeros(n, np.int32)
    children, word_id, weight = [], np.full(n, -1, np.int32), np.zeros(n, np.float64)
    level = {0: 0}
    next_id, n_words = 1, 0
    order = [0]
    while order:
        i = order.pop(0)
        if level[i] == L:
            word_id[i] = n_words; n_words += 1
            weight[i] = 0.0 if rng.rand() < stop_fraction else float(np.float64(np.log(1.0 + 50.0 * rng.rand() + 1.0)))
            continue
        ids = list(range(next_id, next_id + k)); next_id += k
        child_begin[i] = len(children); child_count[i] = k
        perm = list(rng.permutation(ids))
        children.extend(perm)
        for c in ids:
            level[c] = level[i] + 1
            if level[i] == 0:
                base = rng.randint(0, 256, 32).astype(np.uint8) if seed_desc is None else seed_desc[rng.randint(len(seed_desc))].copy()
            else:
                base = desc[i].copy()
                flips = rng.randint(0, 256, max(2, 24 >> level[i]))          # flip a few bits of the parent's descriptor
                for b in flips:
                    base[b >> 3] ^= np.uint8(1 << (b & 7))
            desc[c] = base
            order.append(c)
    return dict(k=k, L=L, node_desc=desc, child_begin=child_begin, child_count=child_count,
                children=np.array(children, np.int32), word_id=word_id, weight=weight)


def transform_numpy(tree, desc, levelsup):
    """Independent restatement of TemplatedVocabulary::transform (feature) in numpy, for pinning the oracle."""
    bits = np.unpackbits(tree["node_desc"], axis=1)
    out_w, out_v, out_n = [], [], []
    nid_level = tree["L"] - levelsup
    for d in desc:
        fb = np.unpackbits(d)
        node, lvl, nid = 0, 0, 0
        while True:
            lvl += 1
            ch = tree["children"][tree["child_begin"][node]:tree["child_begin"][node] + tree["child_count"][node]]
            dist = (bits[ch] != fb).sum(1)
            node = int(ch[int(np.argmin(dist))])                 # first minimum
            if lvl == nid_level:
                nid = node
            if tree["child_count"][node] == 0:
                break
        out_w.append(tree["word_id"][node]); out_v.append(tree["weight"][node]); out_n.append(nid)
    return np.array(out_w, np.int32), np.array(out_v, np.float64), np.array(out_n, np.int32)


def as_text_file(tree):
    """The vocabulary in the reference's ORBvoc.txt format (TemplatedVocabulary::loadFromTextFile,
    Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1337-1407): header `k L scoring weighting` (0 0 = L1_NORM, TF_IDF), then one line per
    non-root node in id order: `parent is_leaf d0 .. d31 weight`.  The loader numbers nodes by line and words by leaf order and lists
    children in line order, so the tree handed to the oracle is `loader_view(tree)`.
    No trailing newline: the loader's `while(!f.eof())` would parse the empty last line into one more node whose parent, leaf flag
    and descriptor are never assigned (indeterminate in the reference, FORB.cpp:120-135) -- nothing a parity test can hold on to."""
    n = len(tree["child_begin"])
    parent = np.zeros(n, np.int64)
    for i in range(n):
        for c in tree["children"][tree["child_begin"][i]:tree["child_begin"][i] + tree["child_count"][i]]:
            parent[c] = i
    lines = [f"{tree['k']} {tree['L']} 0 0"]
    for i in range(1, n):
        leaf = int(tree["child_count"][i] == 0)
        lines.append(f"{parent[i]} {leaf} " + " ".join(str(int(x)) for x in tree["node_desc"][i]) + f" {float(tree['weight'][i])!r}")
    return np.frombuffer("\n".join(lines).encode(), np.uint8).copy()


def loader_view(tree):
    """What loadFromTextFile builds from as_text_file(tree): the same nodes with every child list in ascending id order."""
    t = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in tree.items()}
    for i in range(len(t["child_begin"])):
        b, c = t["child_begin"][i], t["child_count"][i]
        t["children"][b:b + c] = np.sort(t["children"][b:b + c])
    return t
