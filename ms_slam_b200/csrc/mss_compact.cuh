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
