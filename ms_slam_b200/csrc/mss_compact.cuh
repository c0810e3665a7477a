// mss_compact.cuh -- stream compaction of the per-keypoint arrays of sparsified keyframes.
//
// KeyFrame::EraseBadDescriptor (/root/reference/src/KeyFrame.cc:311-361) rebuilds, on the host and row by row, the ORB
// descriptor matrix (cv::Mat, 32 bytes per row), the undistorted keypoints (cv::KeyPoint, 28 bytes), mvuRight and mvDepth of a
// keyframe so that only the rows whose slot still holds a map point survive, in order.  Here the same compaction runs on
// device-resident arrays, one CTA per keyframe, in place: rows are handled in chunks of 512 in ascending order; a chunk
// is read completely (registers) before any of its rows is written, and a row never moves up, so no row is overwritten
// before it has been read.  Pure data movement: 68 bytes read per row, 68 bytes written per surviving row -> HBM-bound.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mssc {

constexpr int kT = 256;
constexpr int kDescBytes = 32;       // ORB descriptor row
constexpr int kKeyBytes = 28;        // cv::KeyPoint: pt.x, pt.y, size, angle, response (float), octave, class_id (int)

struct KfPayload {
    int n;
    int kf;                          // mirror handle or -1
    const uint8_t* keep;             // [n] or nullptr (flags from the mirror: slot holds a map point)
    uint4* desc;                     // [n][2]
    uint32_t* keys;                  // [n][7]
    float* uright;
    float* depth;
};

constexpr int kRpt = 2;              // rows per thread and chunk: 512 rows = 34 KB are in flight per CTA between two barriers
constexpr int kChunk = kT * kRpt;

// flags_mirror: slot_mp of the mirror (keyframe-major, S per keyframe) or nullptr
__global__ void __launch_bounds__(kT) compact_keyframes_kernel(const KfPayload* P, int nkf, const int* slot_mp, int S, int* n_out) {
    __shared__ int s_w[kT / 32];
    __shared__ int s_new[kChunk];
    for (int q = blockIdx.x; q < nkf; q += gridDim.x) {
        const KfPayload p = P[q];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int out = 0;
        for (int c0 = 0; c0 < p.n; c0 += kChunk) {
            const int m = min(kChunk, p.n - c0);
            // thread t owns rows c0 + t * kRpt .. + kRpt - 1 (consecutive: row order = (thread, j) order)
            bool keep[kRpt];
            int x = 0;
#pragma unroll
            for (int j = 0; j < kRpt; ++j) {
                const int r = (int)threadIdx.x * kRpt + j, i = c0 + r;
                keep[j] = false;
                if (r < m) keep[j] = p.keep ? p.keep[i] != 0 : (slot_mp && p.kf >= 0 && i < S && slot_mp[(size_t)p.kf * S + i] >= 0);
                x += keep[j] ? 1 : 0;
            }
            const int mine = x;
            // the whole chunk goes into registers before anything is written: descriptors as 2 x 16 bytes per row, keypoints as
            // 7 words per row, both by flat index = coalesced
            uint4 d[2 * kRpt];
            uint32_t k[7 * kRpt];
            float ur[kRpt], dp[kRpt];
#pragma unroll
            for (int j = 0; j < 2 * kRpt; ++j) {
                const int it = j * kT + (int)threadIdx.x;
                if (p.desc && it < 2 * m) d[j] = p.desc[(size_t)c0 * 2 + it];
            }
#pragma unroll
            for (int j = 0; j < 7 * kRpt; ++j) {
                const int it = j * kT + (int)threadIdx.x;
                if (p.keys && it < 7 * m) k[j] = p.keys[(size_t)c0 * 7 + it];
            }
#pragma unroll
            for (int j = 0; j < kRpt; ++j) {
                const int r = (int)threadIdx.x * kRpt + j;
                ur[j] = 0.f; dp[j] = 0.f;
                if (r < m) { if (p.uright) ur[j] = p.uright[c0 + r]; if (p.depth) dp[j] = p.depth[c0 + r]; }
            }
            // exclusive scan of the per-thread counts over the chunk
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
            __syncthreads();                                   // (s_new / s_w of the previous chunk are no longer read)
            if (lane == 31) s_w[wid] = x;
            __syncthreads();
            int off = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kT / 32; ++w) { const int t = s_w[w]; if (w < wid) off += t; tot += t; }
            int pos = out + off + x - mine;
#pragma unroll
            for (int j = 0; j < kRpt; ++j) { s_new[(int)threadIdx.x * kRpt + j] = keep[j] ? pos : -1; pos += keep[j] ? 1 : 0; }
            __syncthreads();                                   // every row of the chunk is in registers, s_new is complete
#pragma unroll
            for (int j = 0; j < 2 * kRpt; ++j) {
                const int it = j * kT + (int)threadIdx.x;
                if (p.desc && it < 2 * m) { const int nr = s_new[it >> 1]; if (nr >= 0) p.desc[(size_t)nr * 2 + (it & 1)] = d[j]; }
            }
#pragma unroll
            for (int j = 0; j < 7 * kRpt; ++j) {
                const int it = j * kT + (int)threadIdx.x;
                if (p.keys && it < 7 * m) { const int r = it / 7, nr = s_new[r]; if (nr >= 0) p.keys[(size_t)nr * 7 + (it - r * 7)] = k[j]; }
            }
#pragma unroll
            for (int j = 0; j < kRpt; ++j) {
                if (!keep[j]) continue;
                const int nr = s_new[(int)threadIdx.x * kRpt + j];
                if (p.uright) p.uright[nr] = ur[j];
                if (p.depth) p.depth[nr] = dp[j];
            }
            out += tot;
        }
        if (threadIdx.x == 0) n_out[q] = out;
        __syncthreads();
    }
}

}  // namespace mssc

// ---------------------------------------------------------------------------------------------------------------------------
// The same compaction with the chunks staged by the TMA unit (sm_90+ bulk copies, sm_100a here): one elected thread issues
// cp.async.bulk global -> shared for the next chunk's arrays (descriptors, keypoints, uRight, depth, flags: five bulk copies
// completing on one mbarrier) while the CTA compacts the current chunk out of shared memory; the surviving descriptor rows are
// gathered into a contiguous shared buffer and leave with ONE cp.async.bulk shared -> global (16-byte aligned: 32-byte rows);
// keypoints (28-byte rows) and the two float arrays are written with ordinary coalesced stores.  No registers hold row data,
// two chunks are in flight per CTA, and nothing but address arithmetic is executed for the descriptor bytes.
// Eligibility (checked on the host): every array 16-byte aligned; the mirror's slot rows 16-byte aligned (S % 4 == 0).  Full
// chunks of 256 rows use the bulk path; the last partial chunk of a keyframe is staged with ordinary loads.
// ---------------------------------------------------------------------------------------------------------------------------
namespace mssc {

constexpr int kC = 256;              // rows per chunk of the TMA variant

struct __align__(128) TmaStage {
    uint4 desc[2 * kC];
    uint32_t keys[7 * kC];
    float ur[kC];
    float dp[kC];
    int flag[kC];                    // mirror: slot_mp values; caller flags: the first kC bytes hold keep[]
};
struct __align__(128) TmaSmem {
    TmaStage st[2];
    uint4 out_desc[2 * kC];
    unsigned long long bar[2];
    int s_w[kT / 32];
    int s_new[kC];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(kT) compact_keyframes_tma_kernel(const KfPayload* P, int nkf, const int* slot_mp, int S, int* n_out) {
    extern __shared__ __align__(128) unsigned char tma_raw[];
    TmaSmem& M = *reinterpret_cast<TmaSmem*>(tma_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&M.bar[0], 1);
        mbar_init(&M.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned uses[2] = {0u, 0u};                         // completed bulk loads per stage (mbarrier phase parity)
    for (int q = blockIdx.x; q < nkf; q += gridDim.x) {
        const KfPayload p = P[q];
        const bool mirror_flags = p.keep == nullptr;
        const int* mflags = (mirror_flags && slot_mp && p.kf >= 0) ? slot_mp + (size_t)p.kf * S : nullptr;
        const int nflag = mirror_flags ? min(p.n, S) : p.n;     // rows beyond the mirror's slots do not survive
        const int nchunks = (p.n + kC - 1) / kC;
        // stage chunk c into M.st[c & 1]: bulk copies for a full chunk, ordinary loads for the partial one
        auto stage_chunk = [&](int c) {
            TmaStage& T = M.st[c & 1];
            const int c0 = c * kC, m = min(kC, p.n - c0);
            const bool have_flags = mirror_flags ? (mflags != nullptr && c0 + m <= nflag) : true;
            if (m == kC && have_flags) {
                if (threadIdx.x == 0) {
                    const uint32_t bytes = (p.desc ? kC * 32u : 0u) + (p.keys ? kC * 28u : 0u) + (p.uright ? kC * 4u : 0u) + (p.depth ? kC * 4u : 0u) +
                                           (mirror_flags ? kC * 4u : (uint32_t)kC);
                    mbar_expect_tx(&M.bar[c & 1], bytes);
                    if (p.desc) bulk_g2s(T.desc, p.desc + (size_t)c0 * 2, kC * 32u, &M.bar[c & 1]);
                    if (p.keys) bulk_g2s(T.keys, p.keys + (size_t)c0 * 7, kC * 28u, &M.bar[c & 1]);
                    if (p.uright) bulk_g2s(T.ur, p.uright + c0, kC * 4u, &M.bar[c & 1]);
                    if (p.depth) bulk_g2s(T.dp, p.depth + c0, kC * 4u, &M.bar[c & 1]);
                    if (mirror_flags) bulk_g2s(T.flag, mflags + c0, kC * 4u, &M.bar[c & 1]);
                    else bulk_g2s(T.flag, p.keep + c0, (uint32_t)kC, &M.bar[c & 1]);
                }
                return true;
            }
            for (int i = threadIdx.x; i < 2 * m; i += kT) if (p.desc) T.desc[i] = p.desc[(size_t)c0 * 2 + i];
            for (int i = threadIdx.x; i < 7 * m; i += kT) if (p.keys) T.keys[i] = p.keys[(size_t)c0 * 7 + i];
            for (int i = threadIdx.x; i < m; i += kT) {
                if (p.uright) T.ur[i] = p.uright[c0 + i];
                if (p.depth) T.dp[i] = p.depth[c0 + i];
                if (mirror_flags) T.flag[i] = (mflags && c0 + i < nflag) ? mflags[c0 + i] : -1;
                else reinterpret_cast<uint8_t*>(T.flag)[i] = p.keep[c0 + i];
            }
            return false;
        };
        int out = 0;
        bool bulk_cur = nchunks > 0 ? stage_chunk(0) : false;
        for (int c = 0; c < nchunks; ++c) {
            TmaStage& T = M.st[c & 1];
            const int c0 = c * kC, m = min(kC, p.n - c0);
            // the other stage was consumed in iteration c - 1 (barrier at its end): refill it with chunk c + 1
            const bool bulk_next = c + 1 < nchunks ? stage_chunk(c + 1) : false;
            if (bulk_cur) { mbar_wait(&M.bar[c & 1], uses[c & 1] & 1u); ++uses[c & 1]; }
            else __syncthreads();                                // ordinary loads of this chunk (issued one iteration ago or just now)
            bool keep = false;
            if ((int)threadIdx.x < m) keep = mirror_flags ? T.flag[threadIdx.x] >= 0 : reinterpret_cast<const uint8_t*>(T.flag)[threadIdx.x] != 0;
            int x = keep ? 1 : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
            if (threadIdx.x == 0) bulk_wait_read();              // the previous chunk's descriptor store has read out_desc
            __syncthreads();                                     // (s_new / s_w / out_desc of the previous chunk are free)
            if (lane == 31) M.s_w[wid] = x;
            __syncthreads();
            int off = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kT / 32; ++w) { const int t = M.s_w[w]; if (w < wid) off += t; tot += t; }
            M.s_new[threadIdx.x] = keep ? off + x - 1 : -1;      // rank inside the chunk
            __syncthreads();
            if (p.desc) {
                for (int it = threadIdx.x; it < 2 * m; it += kT) { const int nr = M.s_new[it >> 1]; if (nr >= 0) M.out_desc[nr * 2 + (it & 1)] = T.desc[it]; }
            }
            if (p.keys) {
                for (int it = threadIdx.x; it < 7 * m; it += kT) { const int r = it / 7, nr = M.s_new[r]; if (nr >= 0) p.keys[(size_t)(out + nr) * 7 + (it - r * 7)] = T.keys[it]; }
            }
            if (keep) {
                const int nr = out + M.s_new[threadIdx.x];
                if (p.uright) p.uright[nr] = T.ur[threadIdx.x];
                if (p.depth) p.depth[nr] = T.dp[threadIdx.x];
            }
            fence_async_smem();                                  // out_desc (generic writes) -> visible to the bulk store
            __syncthreads();                                     // stage c & 1 and out_desc complete; stage free for chunk c + 2
            if (threadIdx.x == 0 && p.desc && tot > 0) bulk_s2g(p.desc + (size_t)out * 2, M.out_desc, (uint32_t)tot * 32u);
            out += tot;
            bulk_cur = bulk_next;
        }
        if (threadIdx.x == 0) { n_out[q] = out; bulk_wait_read(); }
        __syncthreads();
    }
    if (threadIdx.x == 0) bulk_wait_all();
}

}  // namespace mssc
