// mss_bow.cu -- bag-of-words re-transform of compacted keyframes and the keyframe database's inverted file, on the device
// (C-ABI in include/mss.h "BoW re-transform").
//
// What it replaces in the reference: after KeyFrame::EraseBadDescriptor has compacted a sparsified keyframe it transforms
// the surviving ORB descriptors again (mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4),
// /root/reference/src/KeyFrame.cc:352-354; DBoW2 TemplatedVocabulary.h:1126-1259: per descriptor a descent of the k-ary
// vocabulary tree by Hamming distance, FORB.cpp:81-101) and LoopClosing::DeleteOutdatedInfo puts the keyframe into the
// KeyFrameDatabase (src/LoopClosing.cc:318-329; KeyFrameDatabase::add appends it to the inverted-file list of every word
// of its BowVector; the detection queries, src/KeyFrameDatabase.cc:610-640, count per database keyframe the words it
// shares with the query).  Here:
//   bow_descend   16 lanes per descriptor: the children of a node lie contiguously (the tree is renumbered breadth first at
//                 load time), every lane takes one child (popc over 8 words), the best child is the minimum of
//                 (distance << 8 | child order) -- the first minimum wins, like the strict '<' of :1244
//   bow_vectors   one CTA per keyframe: (word, feature) and (node, feature) keys sorted in shared memory (bitonic), runs ->
//                 BowVector (weights added once per feature in feature order, L1-normalised in word order: the additions
//                 happen in the order std::map iteration gives them upstream, so the doubles are bit-identical) and
//                 FeatureVector (node -> ascending feature indices)
//   db_add / db_common   postings (word, keyframe) appended to the inverted file; words in common per database keyframe
// The vocabulary (ORBvoc: k = 10, L = 6, ~1.1 M nodes x 32 B = 35 MB) stays resident in L2; the work is popc + L2 traffic.
#include "../../include/mss.h"
#define MSS_KERNELS_TYPES_ONLY
#include "mss_internal.h"

#include <algorithm>
#include <cstring>
#include <vector>

using namespace mssi;

namespace mssb {

constexpr int kT = 256;
constexpr int kGroup = 16;           // lanes per descriptor
constexpr int kMaxFeat = 4096;       // features per keyframe handled by the shared-memory sort

struct VocDev {
    const uint4* desc;       // [n][2] descriptors, breadth-first order
    const int* child0;       // [n] first child (breadth-first id) or -1
    const int* nchild;       // [n]
    const int* orig;         // [n] node id of the text file (DBoW2 NodeId)
    const int* word;         // [n] word id or -1
    const double* weight;    // [n]
    int n, L;
};

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// word / node of every descriptor of a batch of keyframes; desc_ptr[q] = descriptors of keyframe q, feat_off[q] = its first
// row in the flat outputs
__global__ void bow_descend(VocDev V, int nkf, const uint4* const* desc_ptr, const int* feat_off, int levelsup, int* out_word,
                            int* out_node, double* out_w) {
    const int total = feat_off[nkf];
    const int lane = threadIdx.x & (kGroup - 1);
    const unsigned gmask = 0xFFFFu << ((threadIdx.x & 16));
    const int nid_level = V.L - levelsup;
    for (int f = (blockIdx.x * blockDim.x + threadIdx.x) / kGroup; f < total; f += gridDim.x * blockDim.x / kGroup) {
        // keyframe of feature f: binary search in the offsets
        int lo = 0, hi = nkf;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (feat_off[mid] <= f) lo = mid; else hi = mid; }
        const uint4* d = desc_ptr[lo] + (size_t)(f - feat_off[lo]) * 2;
        const uint4 a0 = d[0], a1 = d[1];
        int cur = 0, level = 0, nid = 0;
        while (true) {
            const int c0 = V.child0[cur], cn = V.nchild[cur];
            if (cn <= 0) break;                                              // leaf (TemplatedVocabulary.h:1254)
            ++level;
            unsigned best = 0xFFFFFFFFu;
            for (int c = lane; c < cn; c += kGroup) {
                const uint4 b0 = V.desc[(size_t)(c0 + c) * 2], b1 = V.desc[(size_t)(c0 + c) * 2 + 1];
                best = min(best, ((unsigned)hamming256(a0, a1, b0, b1) << 8) | (unsigned)min(c, 255));
            }
#pragma unroll
            for (int o = kGroup / 2; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(gmask, best, o, kGroup));
            cur = c0 + (int)(best & 0xFFu);
            if (level == nid_level) nid = cur;
        }
        if (nid_level > level) nid = cur;         // leaf above the requested level (upstream leaves *nid unset): the leaf itself
        if (lane == 0) {
            out_word[f] = V.word[cur];
            out_node[f] = nid_level <= 0 ? 0 : V.orig[nid];
            out_w[f] = V.weight[cur];
        }
    }
}

__device__ __forceinline__ void bitonic_sort(unsigned long long* key, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < n_pow2; i += kT) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = key[i], b = key[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { key[i] = b; key[p] = a; }
                }
            }
        }
    __syncthreads();
}

// one CTA per keyframe: BowVector and FeatureVector from the per-feature words / nodes
__global__ void bow_vectors(int nkf, const int* feat_off, const int* word, const int* node, const double* w, int* bow_word,
                            double* bow_val, int* n_bow, int* fv_node, int* fv_feat, int* n_fv, int* err) {
    extern __shared__ unsigned long long key[];
    __shared__ int s_cnt;
    __shared__ double s_norm;
    for (int q = blockIdx.x; q < nkf; q += gridDim.x) {
        const int f0 = feat_off[q], n = feat_off[q + 1] - f0;
        if (n > kMaxFeat) { if (threadIdx.x == 0) { atomicOr(err, 1); n_bow[q] = 0; n_fv[q] = 0; } continue; }
        int np2 = 1;
        while (np2 < n) np2 <<= 1;
        // ---- BowVector: keys (word, feature); stopped words (weight <= 0) drop out (TemplatedVocabulary.h:1157) -------------
        for (int i = threadIdx.x; i < np2; i += kT)
            key[i] = (i < n && w[f0 + i] > 0.0) ? ((unsigned long long)(unsigned)word[f0 + i] << 32) | (unsigned)i : ~0ull;
        bitonic_sort(key, np2);
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        // run starts -> slots (ordered: a run's slot = number of run starts before it); done by a serial scan of thread 0
        // over at most 4096 keys -- the additions below must happen in map order anyway
        if (threadIdx.x == 0) {
            int m = 0;
            double norm = 0.0;
            int i = 0;
            while (i < n && key[i] != ~0ull) {
                const unsigned wd = (unsigned)(key[i] >> 32);
                double v = 0.0;
                while (i < n && key[i] != ~0ull && (unsigned)(key[i] >> 32) == wd) { v += w[f0 + (int)(unsigned)key[i]]; ++i; }   // addWeight, feature order
                bow_word[f0 + m] = (int)wd;
                bow_val[f0 + m] = v;
                norm += fabs(v);                                             // BowVector::normalize(L1), word order
                ++m;
            }
            s_cnt = m;
            n_bow[q] = m;
            s_norm = norm;                   // (the division runs in parallel below)
        }
        __syncthreads();
        {
            const int m = s_cnt;
            const double norm = s_norm;
            if (norm > 0.0)
                for (int i = threadIdx.x; i < m; i += kT) bow_val[f0 + i] = bow_val[f0 + i] / norm;
        }
        __syncthreads();
        // ---- FeatureVector: keys (node, feature) -> pairs in map order ------------------------------------------------------
        for (int i = threadIdx.x; i < np2; i += kT)
            key[i] = (i < n && w[f0 + i] > 0.0) ? ((unsigned long long)(unsigned)node[f0 + i] << 32) | (unsigned)i : ~0ull;
        bitonic_sort(key, np2);
        int cnt = 0;
        for (int i = threadIdx.x; i < n; i += kT)
            if (key[i] != ~0ull) { fv_node[f0 + i] = (int)(key[i] >> 32); fv_feat[f0 + i] = (int)(unsigned)key[i]; ++cnt; }
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (threadIdx.x == 0) n_fv[q] = s_cnt;
        __syncthreads();
    }
}

// KeyFrameDatabase::add for a batch: one posting per (keyframe, word of its BowVector)
__global__ void db_add(int nkf, const int* feat_off, const int* kf_id, const int* bow_word, const int* n_bow, int2* postings,
                       int* cursor, int cap, int* word_count, int* err) {
    for (int q = blockIdx.x; q < nkf; q += gridDim.x) {
        const int f0 = feat_off[q], m = n_bow[q];
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            const int wd = bow_word[f0 + i];
            const int p = atomicAdd(cursor, 1);
            if (p < cap) { postings[p] = make_int2(wd, kf_id[q]); atomicAdd(&word_count[wd], 1); }
            else atomicOr(err, 2);
        }
    }
}

// words in common with a query BowVector per database keyframe (src/KeyFrameDatabase.cc:610-640: mnPlaceRecognitionWords)
__global__ void db_mark(const int* qwords, int nq, unsigned* bitmap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) atomicOr(&bitmap[qwords[i] >> 5], 1u << (qwords[i] & 31));
}
__global__ void db_common(const int2* postings, int n, const unsigned* bitmap, int* common, int kf_cap) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int2 p = postings[i];
        if (((bitmap[p.x >> 5] >> (p.x & 31)) & 1u) && p.y >= 0 && p.y < kf_cap) atomicAdd(&common[p.y], 1);
    }
}

}  // namespace mssb

struct mss_vocabulary {
    mss_handle* h = nullptr;
    int n = 0, L = 0, n_words = 0;
    DevBuf<uint4> desc;
    DevBuf<int> ints;            // child0 | nchild | orig | word
    DevBuf<double> weight;
    // scratch of mss_bow_transform
    DevBuf<uint8_t> scratch;
    // inverted file
    DevBuf<int2> postings;
    DevBuf<int> word_count;      // [n_words] + cursor + err
    DevBuf<unsigned> bitmap;
    int n_postings = 0;
    uint8_t* h_pin = nullptr; size_t h_pin_cap = 0;      // pinned hand-back buffer
};

extern "C" {

int mss_voc_create(mss_handle* h, int32_t n_nodes, int32_t levels, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* descriptors,
                   const double* weight, mss_vocabulary** out) {
    if (!h || !out) return MSS_E_BADARG;
    *out = nullptr;
    h->err.clear();
    if (n_nodes < 1 || levels < 1 || !parent || !is_leaf || !descriptors || !weight) { h->err = "voc_create: bad arguments"; return MSS_E_BADARG; }
    // children in order of appearance (TemplatedVocabulary.h:1388-1391), then breadth-first renumbering: the children of a
    // node become contiguous
    std::vector<std::vector<int>> ch(n_nodes);
    for (int i = 1; i < n_nodes; ++i) {
        if (parent[i] < 0 || parent[i] >= i) { h->err = "voc_create: node " + std::to_string(i) + ": parent must precede it"; return MSS_E_BADARG; }
        ch[parent[i]].push_back(i);
    }
    std::vector<int> order;      // breadth-first id -> text-file id
    order.reserve(n_nodes);
    order.push_back(0);
    std::vector<int> child0(n_nodes, -1), nchild(n_nodes, 0);
    for (size_t b = 0; b < order.size(); ++b) {
        const int o = order[b];
        if (!ch[o].empty() && ch[o].size() > 255) { h->err = "voc_create: more than 255 children"; return MSS_E_BADARG; }
        child0[b] = ch[o].empty() ? -1 : (int)order.size();
        nchild[b] = (int)ch[o].size();
        for (int c : ch[o]) order.push_back(c);
    }
    if ((int)order.size() != n_nodes) { h->err = "voc_create: the nodes do not form one tree"; return MSS_E_BADARG; }
    std::vector<int> word_of(n_nodes, -1);          // text-file id -> word id: leaves in file order (:1408-1414)
    int nw = 0;
    for (int i = 1; i < n_nodes; ++i) if (is_leaf[i]) word_of[i] = nw++;
    std::vector<int> ints((size_t)4 * n_nodes);
    std::vector<uint8_t> d((size_t)n_nodes * 32);
    std::vector<double> wt(n_nodes);
    for (int b = 0; b < n_nodes; ++b) {
        const int o = order[b];
        ints[b] = child0[b]; ints[(size_t)n_nodes + b] = nchild[b]; ints[(size_t)2 * n_nodes + b] = o; ints[(size_t)3 * n_nodes + b] = word_of[o];
        memcpy(&d[(size_t)b * 32], descriptors + (size_t)o * 32, 32);
        wt[b] = weight[o];
    }
    mss_vocabulary* v = new (std::nothrow) mss_vocabulary();
    if (!v) return MSS_E_NOMEM;
    v->h = h; v->n = n_nodes; v->L = levels; v->n_words = nw;
    auto upload = [&]() -> int {
        MSS_CUDA(h, cudaSetDevice(h->device));
        int rc;
        if ((rc = ensure(h, v->desc, (size_t)n_nodes * 2)) || (rc = ensure(h, v->ints, (size_t)4 * n_nodes)) || (rc = ensure(h, v->weight, (size_t)n_nodes)) ||
            (rc = ensure(h, v->word_count, (size_t)std::max(nw, 1) + 8)) || (rc = ensure(h, v->bitmap, (size_t)(nw + 31) / 32 + 1))) return rc;
        MSS_CUDA(h, cudaMemcpyAsync(v->desc.p, d.data(), d.size(), cudaMemcpyHostToDevice, h->stream));
        MSS_CUDA(h, cudaMemcpyAsync(v->ints.p, ints.data(), ints.size() * 4, cudaMemcpyHostToDevice, h->stream));
        MSS_CUDA(h, cudaMemcpyAsync(v->weight.p, wt.data(), wt.size() * 8, cudaMemcpyHostToDevice, h->stream));
        MSS_CUDA(h, cudaMemsetAsync(v->word_count.p, 0, ((size_t)std::max(nw, 1) + 8) * 4, h->stream));
        MSS_CUDA(h, cudaStreamSynchronize(h->stream));
        return MSS_OK;
    };
    if (const int rc = upload()) { mss_voc_destroy(v); return rc; }      // (nothing is left behind on a failed upload)
    *out = v;
    return MSS_OK;
}

void mss_voc_destroy(mss_vocabulary* v) {
    if (!v) return;
    cudaSetDevice(v->h->device);
    release(v->desc); release(v->ints); release(v->weight); release(v->scratch); release(v->postings); release(v->word_count); release(v->bitmap);
    if (v->h_pin) cudaFreeHost(v->h_pin);
    delete v;
}

int32_t mss_voc_words(const mss_vocabulary* v) { return v ? v->n_words : 0; }

static mssb::VocDev dev_of(const mss_vocabulary* v) {
    return mssb::VocDev{v->desc.p, v->ints.p, v->ints.p + v->n, v->ints.p + 2 * (size_t)v->n, v->ints.p + 3 * (size_t)v->n, v->weight.p, v->n, v->L};
}

int mss_bow_transform(mss_vocabulary* v, int32_t nkf, const mss_bow_keyframe* kfs, int32_t levelsup, int32_t add_to_database) {
    if (!v) return MSS_E_BADARG;
    mss_handle* h = v->h;
    h->err.clear();
    if (nkf < 0 || (nkf > 0 && !kfs)) { h->err = "bow_transform: bad arguments"; return MSS_E_BADARG; }
    if (nkf == 0) return MSS_OK;
    MSS_CUDA(h, cudaSetDevice(h->device));
    std::vector<int> off(nkf + 1, 0), ids(nkf);
    std::vector<const void*> dptr(nkf);
    for (int q = 0; q < nkf; ++q) {
        if (kfs[q].n < 0 || kfs[q].n > mssb::kMaxFeat || (kfs[q].n > 0 && !kfs[q].descriptors)) {
            h->err = "bow_transform: keyframe " + std::to_string(q) + ": 0..4096 descriptors expected";
            return MSS_E_BADARG;
        }
        off[q + 1] = off[q] + kfs[q].n; ids[q] = kfs[q].kf_id; dptr[q] = kfs[q].descriptors;
    }
    const int total = off[nkf];
    // scratch: desc_ptr[nkf] | feat_off[nkf+1] | kf_id[nkf] | n_bow[nkf] | n_fv[nkf] | err | word[T] | node[T] | bow_word[T] | fv_node[T] |
    //          fv_feat[T] | w[T] | bow_val[T]
    size_t o_ptr = 0, o_off = align_up((size_t)nkf * 8, 16), o_id = o_off + align_up((size_t)(nkf + 1) * 4, 16), o_nb = o_id + align_up((size_t)nkf * 4, 16),
           o_nf = o_nb + align_up((size_t)nkf * 4, 16), o_err = o_nf + align_up((size_t)nkf * 4, 16), o_word = o_err + 16,
           o_node = o_word + align_up((size_t)total * 4, 16), o_bw = o_node + align_up((size_t)total * 4, 16), o_fn = o_bw + align_up((size_t)total * 4, 16),
           o_ff = o_fn + align_up((size_t)total * 4, 16), o_w = o_ff + align_up((size_t)total * 4, 16), o_bv = o_w + align_up((size_t)total * 8, 16),
           bytes = o_bv + align_up((size_t)total * 8, 16);
    int rc;
    if ((rc = ensure(h, v->scratch, bytes + 16))) return rc;
    uint8_t* S = v->scratch.p;
    std::vector<uint8_t> up(o_nb);
    memcpy(up.data() + o_ptr, dptr.data(), (size_t)nkf * 8);
    memcpy(up.data() + o_off, off.data(), (size_t)(nkf + 1) * 4);
    memcpy(up.data() + o_id, ids.data(), (size_t)nkf * 4);
    MSS_CUDA(h, cudaMemcpyAsync(S, up.data(), up.size(), cudaMemcpyHostToDevice, h->stream));
    MSS_CUDA(h, cudaMemsetAsync(S + o_err, 0, 16, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));                 // `up` is pageable
    int *d_off = (int*)(S + o_off), *d_id = (int*)(S + o_id), *d_nb = (int*)(S + o_nb), *d_nf = (int*)(S + o_nf), *d_err = (int*)(S + o_err);
    int *d_word = (int*)(S + o_word), *d_node = (int*)(S + o_node), *d_bw = (int*)(S + o_bw), *d_fn = (int*)(S + o_fn), *d_ff = (int*)(S + o_ff);
    double *d_w = (double*)(S + o_w), *d_bv = (double*)(S + o_bv);
    MSS_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (total > 0) {
        const int groups_per_cta = mssb::kT / mssb::kGroup;
        const int grid = std::max(1, std::min((total + groups_per_cta - 1) / groups_per_cta, h->sm_count * 16));
        mssb::bow_descend<<<grid, mssb::kT, 0, h->stream>>>(dev_of(v), nkf, (const uint4* const*)(S + o_ptr), d_off, levelsup, d_word, d_node, d_w);
    }
    mssb::bow_vectors<<<std::min(nkf, h->sm_count * 4), mssb::kT, (size_t)mssb::kMaxFeat * 8, h->stream>>>(nkf, d_off, d_word, d_node, d_w, d_bw, d_bv, d_nb,
                                                                                                         d_fn, d_ff, d_nf, d_err);
    h->stats.kernel_launches += total > 0 ? 2 : 1;
    if (add_to_database) {
        // every feature can contribute at most one posting
        const size_t need = (size_t)v->n_postings + (size_t)total;
        if ((rc = ensure(h, v->postings, need + 16, true))) return rc;
        int* d_cursor = v->word_count.p + std::max(v->n_words, 1);
        mssb::db_add<<<std::min(nkf, h->sm_count * 4), mssb::kT, 0, h->stream>>>(nkf, d_off, d_id, d_bw, d_nb, v->postings.p, d_cursor, (int)v->postings.cap,
                                                                                v->word_count.p, d_err);
        h->stats.kernel_launches += 1;
        MSS_CUDA(h, cudaMemcpyAsync(&v->n_postings, d_cursor, 4, cudaMemcpyDeviceToHost, h->stream));
    }
    MSS_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    MSS_CUDA(h, cudaGetLastError());
    // hand-back: everything the kernels produced comes back with ONE copy into pinned memory (per-keyframe copies into the
    // caller's pageable arrays would cost ~10 us each), then plain memcpy on the host
    if ((rc = ensure_pinned(h, (void**)&v->h_pin, &v->h_pin_cap, bytes - o_nb))) return rc;
    MSS_CUDA(h, cudaMemcpyAsync(v->h_pin, S + o_nb, bytes - o_nb, cudaMemcpyDeviceToHost, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->stats.last_device_ms = ms;  // descent + vectors (+ inverted file)
    }
    const uint8_t* H = v->h_pin - o_nb;                              // same offsets as on the device
    const int *nb = (const int*)(H + o_nb), *nf = (const int*)(H + o_nf);
    if (*(const int*)(H + o_err)) { h->err = "bow_transform: device error " + std::to_string(*(const int*)(H + o_err)); return MSS_E_INTERNAL; }
    for (int q = 0; q < nkf; ++q) {
        const mss_bow_keyframe& k = kfs[q];
        const size_t f0 = (size_t)off[q];
        if (k.n_bow) *k.n_bow = nb[q];
        if (k.n_fv) *k.n_fv = nf[q];
        if (k.word && k.n) memcpy(k.word, (const int*)(H + o_word) + f0, (size_t)k.n * 4);
        if (k.node && k.n) memcpy(k.node, (const int*)(H + o_node) + f0, (size_t)k.n * 4);
        if (k.bow_word && nb[q]) memcpy(k.bow_word, (const int*)(H + o_bw) + f0, (size_t)nb[q] * 4);
        if (k.bow_value && nb[q]) memcpy(k.bow_value, (const double*)(H + o_bv) + f0, (size_t)nb[q] * 8);
        if (k.fv_node && nf[q]) memcpy(k.fv_node, (const int*)(H + o_fn) + f0, (size_t)nf[q] * 4);
        if (k.fv_feature && nf[q]) memcpy(k.fv_feature, (const int*)(H + o_ff) + f0, (size_t)nf[q] * 4);
    }
    return MSS_OK;
}

int mss_kfdb_common_words(mss_vocabulary* v, int32_t n_query_words, const int32_t* query_words, int32_t kf_cap, int32_t* common) {
    if (!v) return MSS_E_BADARG;
    mss_handle* h = v->h;
    h->err.clear();
    if (n_query_words < 0 || kf_cap < 0 || (n_query_words > 0 && !query_words) || (kf_cap > 0 && !common)) { h->err = "kfdb_common_words: bad arguments"; return MSS_E_BADARG; }
    for (int i = 0; i < n_query_words; ++i)
        if (query_words[i] < 0 || query_words[i] >= v->n_words) { h->err = "kfdb_common_words: word id out of range"; return MSS_E_BADARG; }
    MSS_CUDA(h, cudaSetDevice(h->device));
    const size_t bw = (size_t)(v->n_words + 31) / 32 + 1;
    const size_t o_q = 0, o_c = align_up((size_t)std::max(n_query_words, 1) * 4, 16), bytes = o_c + (size_t)std::max(kf_cap, 1) * 4;
    int rc;
    if ((rc = ensure(h, v->scratch, bytes + 16))) return rc;
    MSS_CUDA(h, cudaMemsetAsync(v->bitmap.p, 0, bw * 4, h->stream));
    MSS_CUDA(h, cudaMemsetAsync(v->scratch.p + o_c, 0, (size_t)std::max(kf_cap, 1) * 4, h->stream));
    if (n_query_words) {
        MSS_CUDA(h, cudaMemcpyAsync(v->scratch.p + o_q, query_words, (size_t)n_query_words * 4, cudaMemcpyHostToDevice, h->stream));
        MSS_CUDA(h, cudaStreamSynchronize(h->stream));
        mssb::db_mark<<<(n_query_words + 255) / 256, 256, 0, h->stream>>>((const int*)(v->scratch.p + o_q), n_query_words, v->bitmap.p);
        if (v->n_postings > 0)
            mssb::db_common<<<std::max(1, std::min((v->n_postings + 255) / 256, h->sm_count * 8)), 256, 0, h->stream>>>(v->postings.p, v->n_postings, v->bitmap.p,
                                                                                                                      (int*)(v->scratch.p + o_c), kf_cap);
        h->stats.kernel_launches += 2;
    }
    MSS_CUDA(h, cudaGetLastError());
    if (kf_cap) MSS_CUDA(h, cudaMemcpyAsync(common, v->scratch.p + o_c, (size_t)kf_cap * 4, cudaMemcpyDeviceToHost, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    return MSS_OK;
}

int32_t mss_kfdb_postings(const mss_vocabulary* v) { return v ? v->n_postings : 0; }

}  // extern "C"
