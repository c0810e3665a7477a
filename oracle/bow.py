"""ORACLE (test infrastructure, not product code) -- DBoW2's bag-of-words transform restated on the CPU.

The reference calls mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) on the compacted keyframe
(/root/reference/src/KeyFrame.cc:352-354) and KeyFrameDatabase::add afterwards (src/LoopClosing.cc:318-329).  The algorithm
lives in /root/reference/Thirdparty/DBoW2/DBoW2 (vendored upstream, but it needs OpenCV's cv::Mat and cannot be compiled
here): TemplatedVocabulary.h:1126-1205 (transform of a feature set), :1217-1259 (descent of one feature), :1338-1418 (tree
from the text file: children in order of appearance, word ids to the leaves in file order), FORB.cpp:81-101 (Hamming
distance), BowVector.cpp:34-46 (addWeight), :62-84 (L1 normalisation), FeatureVector.cpp:31-45 (addFeature).
PARITY UNPINNED upstream (the reference holds no vectors for this step); the vocabulary in the tests is synthetic (the real
ORBvoc.txt is 145 MB and not shipped to the GPU box).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


class Vocabulary:
    def __init__(self, voc):
        self.parent = np.asarray(voc["parent"], np.int64)
        self.is_leaf = np.asarray(voc["is_leaf"], bool)
        self.desc = np.asarray(voc["desc"], np.uint8)
        self.weight = np.asarray(voc["weight"], np.float64)
        self.L = int(voc["L"])
        n = self.parent.size
        self.children = [[] for _ in range(n)]
        for i in range(1, n):
            self.children[int(self.parent[i])].append(i)                 # m_nodes[pid].children.push_back(nid)   (:1391)
        self.word_id = np.full(n, -1, np.int64)
        w = 0
        for i in range(1, n):
            if self.is_leaf[i]:                                          # (:1408-1414)
                self.word_id[i] = w
                w += 1
        self.n_words = w

    def transform_one(self, f, levelsup):
        """(:1217-1259) -> (word id, weight, node id at level L - levelsup)"""
        nid_level = self.L - levelsup
        nid = 0
        final, level = 0, 0
        while self.children[final]:
            level += 1
            ch = self.children[final]
            d = _POP[np.bitwise_xor(self.desc[ch], f[None, :])].sum(axis=1)
            final = ch[int(np.argmin(d))]                                # strict '<': the first minimum wins (:1244)
            if level == nid_level:
                nid = final
        if nid_level > level:
            nid = final          # a leaf above the requested level: upstream leaves *nid unset; the device reports the leaf itself
        if nid_level <= 0:
            nid = 0
        return int(self.word_id[final]), float(self.weight[final]), int(nid)

    def transform(self, descriptors, levelsup=4):
        """(:1126-1205, TF-IDF weighting + L1 scoring) -> dict(word, node, bow_word, bow_value, fv_node, fv_feature)"""
        descriptors = np.asarray(descriptors, np.uint8).reshape(-1, 32)
        n = descriptors.shape[0]
        word, node = np.zeros(n, np.int32), np.zeros(n, np.int32)
        bow, fv = {}, {}
        for i in range(n):
            wid, w, nid = self.transform_one(descriptors[i], levelsup)
            word[i], node[i] = wid, nid
            if w > 0:                                                    # not stopped (:1157)
                bow[wid] = bow.get(wid, 0.0) + w                         # addWeight
                fv.setdefault(nid, []).append(i)                         # addFeature
        keys = sorted(bow)
        norm = 0.0
        for k in keys:                                                   # BowVector::normalize(L1): std::map order
            norm += abs(bow[k])
        vals = [bow[k] / norm if norm > 0.0 else bow[k] for k in keys]
        fn, ff = [], []
        for k in sorted(fv):
            fn += [k] * len(fv[k])
            ff += fv[k]
        return dict(word=word, node=node, bow_word=np.array(keys, np.int32), bow_value=np.array(vals, np.float64),
                    fv_node=np.array(fn, np.int32), fv_feature=np.array(ff, np.int32))


class KeyFrameDatabase:
    """inverted file: word -> keyframes (KeyFrameDatabase::add) and the words-in-common count of the detection queries
    (/root/reference/src/KeyFrameDatabase.cc:610-640)"""
    def __init__(self, n_words):
        self.inv = [[] for _ in range(n_words)]

    def add(self, kf_id, bow_words):
        for w in bow_words:
            self.inv[int(w)].append(int(kf_id))

    def common_words(self, query_words, kf_cap):
        out = np.zeros(kf_cap, np.int32)
        for w in query_words:
            for kf in self.inv[int(w)]:
                if 0 <= kf < kf_cap:
                    out[kf] += 1
        return out
