"""BoW re-transform of compacted keyframes + keyframe-database inverted file on the device (include/mss.h mss_voc_* /
mss_bow_transform / mss_kfdb_*) against oracle/bow.py (DBoW2's transform restated: TemplatedVocabulary.h:1126-1259).
Words, nodes and the FeatureVector are exact; the BowVector doubles are bit-identical (same additions in the same order)."""
import numpy as np
import pytest

from conftest import LAM, GLAM
from ms_slam_b200 import bow as B
from oracle import bow as ob


def test_oracle_on_a_hand_made_tree():
    """root -> a (inner), b (leaf); a -> c, d (leaves).  Words in file order: b = 0, c = 1, d = 2."""
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    half = np.concatenate([np.full(16, 255, np.uint8), np.zeros(16, np.uint8)])
    voc = dict(parent=[0, 0, 0, 1, 1], is_leaf=[0, 0, 1, 1, 1], desc=np.stack([z, z, o, z, half]), weight=[0, 0, 2.0, 3.0, 0.5], L=2)
    V = ob.Vocabulary(voc)
    assert V.n_words == 3 and V.word_id.tolist() == [-1, -1, 0, 1, 2]
    assert V.transform_one(o, 1) == (0, 2.0, 2)              # nearest child of the root is b: a leaf above level 1 -> itself
    assert V.transform_one(z, 1)[:2] == (1, 3.0) and V.transform_one(z, 1)[2] == 1
    assert V.transform_one(half, 0)[0] == 2                  # tie at the root (128 / 128): the first child (a) wins, then d
    t = V.transform(np.stack([z, z, half, o]), levelsup=1)
    assert t["bow_word"].tolist() == [0, 1, 2]
    assert np.allclose(t["bow_value"], np.array([2.0, 6.0, 0.5]) / 8.5) and t["bow_value"].sum() == pytest.approx(1.0)
    assert t["fv_node"].tolist() == [1, 1, 1, 2] and t["fv_feature"].tolist() == [0, 1, 2, 3]


def test_synthetic_vocabulary_is_a_valid_dbow2_file_order():
    voc = B.synthetic_vocabulary(k=10, L=3, seed=1)
    p = voc["parent"]
    assert (p[1:] < np.arange(1, p.size)).all() and voc["is_leaf"][0] == 0
    V = ob.Vocabulary(voc)
    assert V.n_words == int(voc["is_leaf"].sum()) and max(len(c) for c in V.children) <= 10


@pytest.fixture(scope="module")
def eng(build_native):
    from ms_slam_b200.engine import Engine
    e = Engine(N=100, lam=LAM, grid_lam=GLAM, device=0)
    yield e
    e.close()


def upload(eng, a):
    p = eng.lib.mss_device_alloc(eng.handle, max(a.nbytes, 16))
    if a.nbytes:
        eng._check(eng.lib.mss_memcpy_h2d(eng.handle, p, a.ctypes.data, a.nbytes))
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("k,L,levelsup,seed", [(10, 3, 2, 0), (10, 4, 4, 1), (7, 3, 1, 2), (16, 2, 0, 3), (20, 2, 3, 4)])
def test_transform_matches_dbow2(eng, k, L, levelsup, seed):
    voc = B.synthetic_vocabulary(k=k, L=L, seed=seed)
    V, O = B.Vocabulary(eng, voc), ob.Vocabulary(voc)
    assert V.n_words == O.n_words
    rng = np.random.default_rng(seed)
    counts = [0, 1, 37, 500, 1200, 2000]
    # half of the descriptors are noisy copies of node descriptors (realistic: small distances, ties), half random
    descs = []
    for n in counts:
        d = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        src = rng.integers(0, voc["desc"].shape[0], size=n)
        noisy = voc["desc"][src] ^ (rng.random((n, 32)) < 0.02).astype(np.uint8)
        pick = rng.random(n) < 0.5
        d[pick] = noisy[pick]
        descs.append(np.ascontiguousarray(d))
    ptrs = [upload(eng, d) for d in descs]
    res = V.transform(ptrs, counts, kf_ids=list(range(len(counts))), levelsup=levelsup, add_to_database=True)
    db = ob.KeyFrameDatabase(O.n_words)
    for q, (d, r) in enumerate(zip(descs, res)):
        t = O.transform(d, levelsup)
        for key in ("word", "node", "bow_word", "fv_node", "fv_feature"):
            assert np.array_equal(r[key], t[key]), (q, key)
        assert np.array_equal(r["bow_value"].view(np.uint64), t["bow_value"].view(np.uint64)), q      # bit-identical doubles
        if r["bow_value"].size:
            assert r["bow_value"].sum() == pytest.approx(1.0)
        db.add(q, t["bow_word"])
    assert V.postings() == sum(len(r["bow_word"]) for r in res)
    # the counting loop of the detection queries: words in common with every database keyframe
    for q in (3, 5):
        assert np.array_equal(V.common_words(res[q]["bow_word"], len(counts)), db.common_words(res[q]["bow_word"], len(counts)))
    assert V.common_words([], len(counts)).sum() == 0
    for p in ptrs:
        eng.lib.mss_device_free(eng.handle, p)
    V.close()


@pytest.mark.gpu
def test_compaction_then_retransform(eng):
    """the f2 -> f4 chain of KeyFrame::EraseBadDescriptor on the device: rows compacted in place, then transformed again"""
    from ms_slam_b200.mirror import KeyframePayload, compact_keyframes
    from oracle import mirror_model as mm
    voc = B.synthetic_vocabulary(k=10, L=3, seed=9)
    V, O = B.Vocabulary(eng, voc), ob.Vocabulary(voc)
    rng = np.random.default_rng(5)
    n = 1500
    desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    keep = rng.random(n) < 0.2
    pay = KeyframePayload(eng, n, keep, descriptors=desc)
    left = int(compact_keyframes(eng, [pay])[0])
    r = V.transform([pay.ptr["descriptors"]], [left], levelsup=4)[0]
    t = O.transform(mm.erase_bad_descriptor_rows(keep, desc)[0], 4)
    assert left == int(keep.sum()) and np.array_equal(r["word"], t["word"]) and np.array_equal(r["bow_word"], t["bow_word"])
    assert np.array_equal(r["bow_value"].view(np.uint64), t["bow_value"].view(np.uint64))
    assert np.array_equal(r["fv_node"], t["fv_node"]) and np.array_equal(r["fv_feature"], t["fv_feature"])
    pay.free()
    V.close()
