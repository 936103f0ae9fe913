// cvad_resample.cuh -- 8 / 24 / 48 kHz -> 16 kHz, one frame-sized chunk at a time.
//
// The reference's resampler is AudioUtils.resample_audio
// (/root/reference/src/real_time_vad/utils/audio.py:19-55): scipy.signal.resample (FFT method)
// to int(len * 16000 / sr) samples, cast to float32.  It is not called on the reference's own
// path (vad_wrapper.py:621-624 is a placeholder); the north star adds it in front of the 16 kHz
// model.  For a fixed (N_in, 512) pair the FFT method is an exact LINEAR operator
// y = R x (SURVEY.md section 7 "Resampling parity"), so the unit of resampling is one chunk of
// N_in = 512 * sr / 16000 source samples (256 / 768 / 1536) -> one 512-sample model frame.
//
//   R[n][m] = (1/N_in) * ( D_K(theta) + nyq(n, m) ),  theta = 2 pi (n/512 - m/N_in),
//   D_K(theta) = 1 + 2 sum_{k=1}^{K-1} cos(k theta) = sin((2K-1) theta/2) / sin(theta/2),  K = min(512, N_in)/2
//   down-sampling (N_in > 512): nyq = 2 cos(pi 512 m / N_in) (-1)^n     (scipy doubles the folded Nyquist bin)
//   up-sampling   (N_in < 512): nyq = cos(K theta)                       (scipy halves the source Nyquist bin)
//
// The kernel is the same streamed-weight FFMA2 GEMM as the model layers: R^T [m][512] flows
// through the bulk-copy ring 8 rows at a time, the audio chunk is staged k-major 128 samples at
// a time, each thread owns 4 items x 8 outputs.  Output: float32 16 kHz audio in HBM, which the
// front-end kernel then frames exactly like native 16 kHz input (gate and NaN check included).
#pragma once
#include "cvad_common.cuh"

namespace cvad {

struct ResampleStep {
    const void *audio;   // device, source rate
    int pcm;
    long long stride;    // source elements between streams
    int n_in;            // source samples per frame: 256 / 768 / 1536
    int n_streams;
    int n_stiles;
    int max_frames;
    const int *n_frames; // or null
    const float *rt;     // R^T [n_in][512]
    float *out;          // [n_streams][max_frames * 512]
    // mixed-rate steps: this launch handles the streams listed in `list` (their number is read from the
    // device word `count`); null = all n_streams in order
    const int *list;
    const int *count;
    int osplit;          // resample_tc_kernel: a tile's four 128-sample output blocks are shared among 1, 2 or 4 CTAs
};

constexpr int kRsPiece = 128;  // source samples staged per piece
constexpr size_t kRsSmemBytes = (size_t)(kRingStages * kRingSlotFloats + 2 * kRsPiece * kTile) * 4 + 64 + 512;

__device__ __forceinline__ void rs_ring_issue(const WeightRing &r, uint32_t g, uint32_t period) {
    const uint32_t slot = g % kRingStages;
    const uint32_t ci = g % period;
    mbar_arrive_expect_tx(&r.bars[slot], kRingSlotFloats * 4u);
    bulk_g2s(r.buf + slot * kRingSlotFloats, r.gsrc + (size_t)ci * kRingSlotFloats, kRingSlotFloats * 4u, &r.bars[slot]);
}

__global__ void __launch_bounds__(kThreads, 1) resample_kernel(const ResampleStep p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring_buf = reinterpret_cast<float *>(smem_raw);
    float *xin = ring_buf + kRingStages * kRingSlotFloats;  // [2][128][32]
    uint64_t *bars = reinterpret_cast<uint64_t *>(xin + 2 * kRsPiece * kTile);
    int *s_valid = reinterpret_cast<int *>(bars + 8);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int tm = tid & 7;
    const int tn = tid >> 3;  // outputs 8tn..8tn+7
    const uint32_t period = (uint32_t)p.n_in / 8u;
    const int n_pieces = p.n_in / kRsPiece;

    WeightRing ring{ring_buf, bars, p.rt, 0u};
    if (tid == 0) {
        for (int i = 0; i < kRingStages; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (uint32_t i = 0; i < kRingStages; ++i) rs_ring_issue(ring, i, period);

    const int n_items = p.count ? *p.count : p.n_streams;
    const int n_stiles = p.count ? (n_items + kTile - 1) / kTile : p.n_stiles;
    const int n_tiles = p.max_frames * n_stiles;
    int *s_gi = s_valid + kTile;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int frame = tile / n_stiles;
        const int st = tile - frame * n_stiles;
        if (tid < kTile) {
            const int j = st * kTile + tid;
            int valid = 0, i = -1;
            if (j < n_items) {
                i = p.list ? p.list[j] : j;
                valid = frame < (p.n_frames ? p.n_frames[i] : p.max_frames);
            }
            s_valid[tid] = valid;
            s_gi[tid] = i;
        }
        const int any_valid = __syncthreads_or(tid < kTile ? s_valid[tid] : 0);
        if (!any_valid) continue;

        // piece loader: lane = item, each thread two float4 of the 128-sample piece
        const int s = lane;
        const int gi = s_gi[s];
        const bool valid = s_valid[s] != 0;
        const long long base = (long long)gi * p.stride + (long long)frame * p.n_in;
        auto load_piece = [&](int piece, float (&v)[2][4]) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int q = it * 16 + warp;  // float4 index within the piece, 0..31
                const long long off = base + (long long)piece * kRsPiece + 4 * q;
#pragma unroll
                for (int e = 0; e < 4; ++e) v[it][e] = 0.f;
                if (valid) {
                    if (p.pcm == 0) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[it][e] = __ldg(reinterpret_cast<const float *>(p.audio) + off + e);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float x = (float)__ldg(reinterpret_cast<const short *>(p.audio) + off + e);
                            v[it][e] = p.pcm == 1 ? __fdiv_rn(x, 32767.0f) : x * (1.0f / 32768.0f);
                        }
                    }
                }
            }
        };
        auto store_piece = [&](int buf, const float (&v)[2][4]) {
            float *dst = xin + buf * kRsPiece * kTile;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int q = it * 16 + warp;
#pragma unroll
                for (int e = 0; e < 4; ++e) dst[(4 * q + e) * kTile + s] = v[it][e];
            }
        };
        float stage[2][4];
        load_piece(0, stage);
        store_piece(0, stage);
        __syncthreads();

        float2 acc2[2][2][4];
        zero_tile(acc2[0]);
        zero_tile(acc2[1]);
        for (int piece = 0; piece < n_pieces; ++piece) {
            if (piece + 1 < n_pieces) load_piece(piece + 1, stage);  // in flight during the math below
            const float *xb = xin + (piece & 1) * kRsPiece * kTile + 4 * tm;
            for (int ci = 0; ci < kRsPiece / 8; ++ci) {
                const float *w = ring_wait(ring);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const float4 a = ld4(xb + (ci * 8 + kk) * kTile);
                    fma4x4(acc2[0], a, dup4(ld4(w + kk * 512 + 8 * tn)));
                    fma4x4(acc2[1], a, dup4(ld4(w + kk * 512 + 8 * tn + 4)));
                }
                __syncthreads();
                if (tid == 0) rs_ring_issue(ring, ring.g + kRingStages, period);
                ++ring.g;
            }
            if (piece + 1 < n_pieces) {
                store_piece((piece + 1) & 1, stage);
                __syncthreads();
            }
        }
        float v0[4][4], v1[4][4];
        unpack_tile(acc2[0], v0);
        unpack_tile(acc2[1], v1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int item = 4 * tm + i;
            const int gidx = s_gi[item];
            if (gidx >= 0 && s_valid[item]) {
                float *dst = p.out + (size_t)gidx * p.max_frames * 512 + (size_t)frame * 512 + 8 * tn;
                st4(dst, make_float4(v0[i][0], v0[i][1], v0[i][2], v0[i][3]));
                st4(dst + 4, make_float4(v1[i][0], v1[i][1], v1[i][2], v1[i][3]));
            }
        }
        __syncthreads();
    }
    for (int i = 0; i < kRingStages; ++i) {
        ring_wait(ring);
        ++ring.g;
    }
}

// ---- mixed-rate steps: per-rate stream lists, and the pass-through of streams that are already at 16 kHz
__device__ __host__ __forceinline__ int rate_slot(int rate) {
    return rate == 8000 ? 0 : rate == 24000 ? 1 : rate == 48000 ? 2 : rate == 16000 ? 3 : -1;
}

// lists[r][0 .. counts[r]) = indices of the streams whose source rate has slot r (order is irrelevant:
// streams are independent); a stream with an unknown rate gets status bit 2.  `init_status` (chained steps: no memset
// ahead of the step): this kernel, the first of the step to touch `status`, writes every stream's word
__global__ void rate_lists_kernel(const int *src_rates, int n, int *lists, int *counts, unsigned int *status,
                                  int init_status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = rate_slot(src_rates[i]);
    if (status && init_status) status[i] = r < 0 ? 2u : 0u;
    if (r < 0) {
        if (status && !init_status) atomicOr(&status[i], 2u);
        return;
    }
    lists[(size_t)r * n + atomicAdd(&counts[r], 1)] = i;
}

// 16 kHz streams of a mixed-rate step: PCM -> float32 frames in the resampler's output layout
__global__ void passthrough_kernel(const ResampleStep p) {
    const int n_items = *p.count;
    const long long per = (long long)p.max_frames * 512;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n_items * per;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx / per);
        const long long k = idx - j * per;
        const int i = p.list[j];
        if ((int)(k >> 9) >= (p.n_frames ? p.n_frames[i] : p.max_frames)) continue;
        const long long off = (long long)i * p.stride + k;
        float x;
        if (p.pcm == 0) x = __ldg(reinterpret_cast<const float *>(p.audio) + off);
        else {
            x = (float)__ldg(reinterpret_cast<const short *>(p.audio) + off);
            x = p.pcm == 1 ? __fdiv_rn(x, 32767.0f) : x * (1.0f / 32768.0f);
        }
        p.out[(size_t)i * per + k] = x;
    }
}

}  // namespace cvad
