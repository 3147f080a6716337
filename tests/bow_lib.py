"""Vocabulary-guided matching (ORBmatcher::SearchByBoW, src/ORBmatcher.cc:552-832): synthetic feature vectors in CSR
form + one Python signature for the reference (oracle/_ref) and the plain-C oracle.  Test infrastructure."""
import ctypes as C

import numpy as np

from matcher_lib import flip_bits
from oracle_lib import oracle, ref

vp, ci, cf = C.c_void_p, C.c_int, C.c_float


def _p(a):
    return None if a is None else a.ctypes.data


def feature_vector(desc, bits=(4, 4)):
    """A stand-in for DBoW2's FeatureVector (NodeId -> feature indices in index order): the node of a feature is
    read off the top bits of its first two descriptor bytes, so near-identical descriptors mostly share a node --
    what a vocabulary tree does.  Returns (node_id [nn] ascending, node_off [nn+1], feat [n])."""
    node = (desc[:, 0].astype(np.int32) >> (8 - bits[0])) * (1 << bits[1]) + (desc[:, 1].astype(np.int32) >> (8 - bits[1]))
    order = np.argsort(node, kind="stable").astype(np.int32)
    ids, counts = np.unique(node, return_counts=True)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return ids.astype(np.int32) * 7 + 3, off, order          # arbitrary ascending NodeIds


def bow_scene(kps, desc, seed, kmax=40, n2=None):
    rng = np.random.default_rng(seed)
    n = len(kps)
    perm = rng.permutation(n)[:n2 or n]
    k2 = kps[perm].copy()
    k2["angle"] = np.where(rng.random(len(k2)) < 0.8, (k2["angle"] + 25 + rng.normal(0, 5, len(k2))) % 360, rng.random(len(k2)) * 360).astype(np.float32)
    d2 = flip_bits(desc[perm], rng, kmax)
    s = dict(k1=kps.copy(), d1=desc.copy(), k2=k2, d2=d2,
             valid1=(rng.random(n) < 0.85).astype(np.uint8), bad1=(rng.random(n) < 0.05).astype(np.uint8),
             valid2=(rng.random(len(k2)) < 0.9).astype(np.uint8), bad2=(rng.random(len(k2)) < 0.05).astype(np.uint8))
    s["fv1"] = feature_vector(s["d1"])
    s["fv2"] = feature_vector(s["d2"])
    return s


def search_by_bow(impl, s, nnratio, check_ori, kf_kf):
    """(nmatches, match [n1] -> feature of side 2 or -1).  kf_kf: the KeyFrame/KeyFrame overload (:700-832), else
    KeyFrame/Frame (:552-697; the reference reports per FRAME feature, converted here)."""
    n1, n2 = len(s["k1"]), len(s["k2"])
    (id1, off1, f1), (id2, off2, f2) = s["fv1"], s["fv2"]
    if impl == "ref":
        L = ref()
        if kf_kf:
            out = np.zeros(n1, np.int32)
            f = L.orbref_search_by_bow_kf_kf
            f.argtypes = [ci, vp, vp, vp, vp, ci, vp, vp, vp] * 2 + [cf, ci, vp]
            nm = f(n1, _p(s["k1"]), _p(s["d1"]), _p(s["valid1"]), _p(s["bad1"]), len(id1), _p(id1), _p(off1), _p(f1),
                   n2, _p(s["k2"]), _p(s["d2"]), _p(s["valid2"]), _p(s["bad2"]), len(id2), _p(id2), _p(off2), _p(f2), nnratio, int(check_ori), _p(out))
            return nm, out
        mf = np.zeros(n2, np.int32)
        f = L.orbref_search_by_bow_kf_frame
        f.argtypes = [ci, vp, vp, vp, vp, ci, vp, vp, vp] + [ci, vp, vp, ci, vp, vp, vp] + [cf, ci, vp]
        nm = f(n1, _p(s["k1"]), _p(s["d1"]), _p(s["valid1"]), _p(s["bad1"]), len(id1), _p(id1), _p(off1), _p(f1),
               n2, _p(s["k2"]), _p(s["d2"]), len(id2), _p(id2), _p(off2), _p(f2), nnratio, int(check_ori), _p(mf))
        out = np.full(n1, -1, np.int32)
        out[mf[mf >= 0]] = np.nonzero(mf >= 0)[0]
        return nm, out
    O = oracle()
    O.orbo_search_by_bow.argtypes = [ci, vp, vp, vp, ci, vp, vp, vp] * 2 + [cf, ci, ci, vp, vp]
    v1 = np.ascontiguousarray(s["valid1"] & (1 - s["bad1"]))
    v2 = np.ascontiguousarray(s["valid2"] & (1 - s["bad2"])) if kf_kf else None
    out = np.zeros(n1, np.int32)
    nm = O.orbo_search_by_bow(n1, _p(s["k1"]), _p(s["d1"]), _p(v1), len(id1), _p(id1), _p(off1), _p(f1),
                              n2, _p(s["k2"]), _p(s["d2"]), _p(v2), len(id2), _p(id2), _p(off2), _p(f2), nnratio, int(check_ori), int(kf_kf), _p(out), None)
    return nm, out
