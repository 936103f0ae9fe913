// cvad_v5.cuh -- Silero VAD v5 (16 kHz branch) as two fused sm_100a kernels.
//
// What the reference runs per frame through onnxruntime
// (/root/reference/src/real_time_vad/core/silero_model.py:433, graph = SURVEY.md 8a "S5"):
//   frame loader (split/zero-pad/denoise)  -> v5_frontend_kernel
//   STFT conv k256 s128 (3 columns) + |.|  -> v5_frontend_kernel
//   encoder.0..3 (conv k3 + ReLU)          -> v5_frontend_kernel   -> feat[128] per (stream, frame)
//   LSTMCell(128), ReLU, 1x1 conv, sigmoid -> v5_recurrent_kernel
//   start/end state machine (:790-923)     -> v5_recurrent_kernel
//
// The front end has no dependence between frames, so it is batched over
// (stream, frame) items; only the recurrent kernel walks frames in order.
// One CTA (512 threads) owns a tile of 32 items; activations live in shared memory
// k-major ([channel][time][item]) so that a thread's 4 items are one LDS.128, and each
// thread owns a 4-item x 4-output register tile for ALL time columns of the layer,
// which lets the k=3 convolutions skip their zero-pad taps statically.  Small layers
// split K over 2 or 4 thread groups and reduce through the output buffer.
// Arithmetic is FP32 (packed FFMA2) with ascending-k accumulation (SURVEY.md section 0
// fact 5: plain TF32/BF16 operands miss the 1e-4 bar).
#pragma once
#include "cvad_common.cuh"

namespace cvad {

// ---------------------------------------------------------------- packed layouts
// Front-end weight stream (floats), one period = 41 chunks:
//   [0      ,  65536) stft   [k 256][n 256]      n=0:re0 n=1:re128 n=2b:re_b n=2b+1:im_b (b=1..127)
//   [65536  , 115072) enc0   [c 129][tap 3][o 128]
//   [115072 , 139648) enc1   [c 128][tap 3][o 64]
//   [139648 , 147840) enc2   [c 64][tap 1..2][o 64]   (tap 0 only ever meets the zero pad)
//   [147840 , 156032) enc3   [c 64][o 128]            (centre tap; T=1 so the others meet zero pad)
constexpr int kFeStreamFloats = 156032;
constexpr int kFeChunks = 41;
// Front-end bias block: enc0[128] enc1[64] enc2[64] enc3[128]
constexpr int kFeBiasFloats = 384;
// Recurrent weight stream: [k 256][n' 512], k<128: W_ih, k>=128: W_hh; n' = 4*unit + gate(i,f,g,o)
constexpr int kRecStreamFloats = 131072;
constexpr int kRecChunks = 32;

__device__ __forceinline__ void fe_chunk(int ci, uint32_t &off, uint32_t &n) {
    if (ci < 16) { off = ci * 4096; n = 4096; }
    else if (ci < 29) { int e = ci - 16; off = 65536 + e * 3840; n = (e < 12) ? 3840 : 3456; }
    else if (ci < 37) { int e = ci - 29; off = 115072 + e * 3072; n = 3072; }
    else if (ci < 39) { int e = ci - 37; off = 139648 + e * 4096; n = 4096; }
    else { int e = ci - 39; off = 147840 + e * 4096; n = 4096; }
}

__device__ __forceinline__ void fe_ring_issue(const WeightRing &r, uint32_t g) {
    uint32_t off, n;
    fe_chunk(static_cast<int>(g % kFeChunks), off, n);
    const uint32_t slot = g % kRingStages;
    mbar_arrive_expect_tx(&r.bars[slot], n * 4u);
    bulk_g2s(r.buf + slot * kRingSlotFloats, r.gsrc + off, n * 4u, &r.bars[slot]);
}

__device__ __forceinline__ void rec_ring_issue(const WeightRing &r, uint32_t g) {
    const uint32_t slot = g % kRingStages;
    const uint32_t ci = g % kRecChunks;
    mbar_arrive_expect_tx(&r.bars[slot], kRingSlotFloats * 4u);
    bulk_g2s(r.buf + slot * kRingSlotFloats, r.gsrc + ci * kRingSlotFloats, kRingSlotFloats * 4u, &r.bars[slot]);
}

// ---------------------------------------------------------------- step parameters
// Resident LSTM state (h and c, 128 rows each: v5 = 128 hidden units, v4 = 2 layers x 64): [slot / 8][row][slot % 8] floats.
// A reader that walks 32 consecutive slots of one row touches four 32-byte sectors (as with a [row][slot] matrix), and the
// 8 streams an epilogue thread of the fused tensor-core kernel owns are ONE 32-byte row when their slots are consecutive
// and start at a multiple of 8 -- its 32 lanes (rows) then read 1 KB contiguous, straight into the registers that need
// them (no transposition through shared memory on the way in or out).  Capacity is rounded up to a multiple of 8 slots.
__host__ __device__ __forceinline__ size_t state_at(int row, int slot) {
    return ((size_t)(slot >> 3) * 128u + (size_t)row) * 8u + (size_t)(slot & 7);
}

struct V5Step {
    const void *audio;            // device
    int pcm;                      // CVAD_PCM_*
    long long stride;             // elements between streams
    int frame_len;                // valid samples per frame (<=512 used)
    int hop;
    int vec_ok;                   // 16-byte vector loads allowed
    int n_streams;
    int n_stiles;                 // ceil(n_streams / 32)
    int max_frames;
    int max_streams;              // engine capacity (leading dimension of transposed state)
    const int *slots;             // [n_streams] or null (identity)
    const int *n_frames;          // [n_streams] or null (max_frames)
    const unsigned char *denoise; // [max_streams]
    unsigned int *status;         // [n_streams]
    float *feat;                  // [max_frames * n_stiles][128][32]
    // front-end weights
    const float *w_fe;
    const float *b_fe;
    // recurrent weights
    const float *w_rec;
    const float *b_rec;           // [512] packed n' order, b_ih + b_hh
    const float *w_dec;           // [128] + bias at [128]
    // per-slot persistent state: [slot / 8][unit (128)][slot % 8] (state_at below)
    float *h_state;
    float *c_state;
    int *sm_active;
    int *sm_scount;
    int *sm_ecount;
    long long *frames_done;
    const double *start_p;
    const double *end_p;
    const int *n_start;
    const int *n_end;
    // outputs
    float *probs;                 // [n_streams][max_frames] or null
    unsigned char *flags;         // [n_streams][max_frames] or null
    void *events;                 // cvad_event[max_events] or null
    int max_events;
    int *n_events;                // device counter or null
    int *ev_ctr;                  // chained fused steps: engine-owned {event count, CTA ticket}, both zero between steps; or null
    int *step_ctr;                // chained fused steps: engine-owned {chained steps completed, CTA ticket}; or null
    int step_seq;                 // chained fused steps: how many chained steps were launched before this one
    int status_zero;              // chained fused steps: the kernel clears `status` itself (no memset ahead of it)
    int commit;                   // 0: do not write state back (debug)
    float *dbg;                   // front-end debug dump for tile 0, or null
    // tensor-core path (cvad_v5tc.cuh): BF16x3 weight tile streams, Nyquist-channel weights, gate biases, hand-off
    const unsigned char *w_fe_tc;
    const unsigned char *w_rec_tc;
    const float *nyq_w;           // [128][4]: encoder.0 weight of input channel 128, taps 0..2
    const float *b_rec_tc;        // [4][128] gate-major (i,f,g,o), b_ih + b_hh
    unsigned char *feat_tc;       // [max_frames * n_stiles][tc5::kFeatTileBytes]
    // FP16-split build of the fused kernel (CVAD_MATH_TC16): weight tile streams with two FP16 parts, scaled per layer
    const unsigned char *w_fe_h;
    const unsigned char *w_rec_h;
    float tc16_inv_w[8];          // 1 / weight scale of: stft, encoder.0..3, W_ih, W_hh
    long long *prof;              // optional clock64 marks of CTA 0 (cvad_set_profile), or null
    int v4_t2;                    // v4 8 kHz sub-model: two time steps reach the LSTM per frame
    float *v4_mag;                // v4 tensor-core path: |STFT| tiles [tile][129][8][16] written by v4tc_stft_kernel, or null
    const float *v4_fft;          // v4 FFT path: exact-basis STFT tiles [tile][128 cols][re | im][132] from v4_stft_fft_kernel, or null
};

struct EventRec {
    int stream, slot, frame, kind;
    long long stream_frame;
};

// Front-end shared memory (floats): ring | bufA (xT / e0T / e2T) | bufB (magT / e1T / scratch)
constexpr int kFeBufA = 512 * kTile;            // 16384 floats: xT[512][32]; e0T[128][3][32]=12288; e2T[64][32]
constexpr int kFeBufB = 129 * 3 * kTile;        // 12384 floats: magT[129][3][32]; e1T[64][2][32]; scratch[128][32]
constexpr size_t kFeSmemBytes =
    (size_t)(kRingStages * kRingSlotFloats + kFeBufA + kFeBufB) * 4 + 64 /*barriers*/ + 256 /*tile meta*/;

// debug dump layout (floats): magT | e0T | e1T | e2T | feat
constexpr int kDbgMag = 129 * 3 * 32, kDbgE0 = 128 * 3 * 32, kDbgE1 = 64 * 2 * 32, kDbgE2 = 64 * 32, kDbgFeat = 128 * 32;
constexpr int kDbgFloats = kDbgMag + kDbgE0 + kDbgE1 + kDbgE2 + kDbgFeat;

__device__ __forceinline__ void block_copy_to_global(float *dst, const float *src, int n) {
    for (int i = threadIdx.x; i < n; i += kThreads) dst[i] = src[i];
}

// =====================================================================================
// Front end: frame loader -> STFT -> magnitude -> encoder.0..3 -> feat
// =====================================================================================
__device__ __forceinline__ void fe_release(WeightRing &ring, int tid) {
    __syncthreads();
    if (tid == 0) fe_ring_issue(ring, ring.g + kRingStages);
    ++ring.g;
}

__global__ void __launch_bounds__(kThreads, 1) v5_frontend_kernel(const V5Step p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring_buf = reinterpret_cast<float *>(smem_raw);
    float *bufA = ring_buf + kRingStages * kRingSlotFloats;
    float *bufB = bufA + kFeBufA;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bufB + kFeBufB);
    int *s_slot = reinterpret_cast<int *>(bars + 8);  // [32]
    int *s_valid = s_slot + kTile;                    // [32]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int tm = tid & 7;   // item group: items 4tm..4tm+3
    const int tq = tid >> 3;  // 0..63

    WeightRing ring{ring_buf, bars, p.w_fe, 0u};
    if (tid == 0) {
        for (int i = 0; i < kRingStages; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (uint32_t i = 0; i < kRingStages; ++i) fe_ring_issue(ring, i);

    const int n_tiles = p.max_frames * p.n_stiles;
    const int flen = p.frame_len < 512 ? p.frame_len : 512;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int frame = tile / p.n_stiles;
        const int st = tile - frame * p.n_stiles;

        // ---- tile metadata
        if (tid < kTile) {
            const int i = st * kTile + tid;
            int valid = 0, slot = -1;
            if (i < p.n_streams) {
                slot = p.slots ? p.slots[i] : i;
                const int nf = p.n_frames ? p.n_frames[i] : p.max_frames;
                valid = frame < nf;
            }
            s_slot[tid] = slot;
            s_valid[tid] = valid;
        }
        const int any_valid = __syncthreads_or(tid < kTile ? s_valid[tid] : 0);
        if (!any_valid) continue;  // uniform: nothing consumed from the ring

        // ---- frame loader: xT[sample][item] = gate(pcm(audio[item][frame*hop + sample]))
        // (audio.py:164-190 split, :104-121 gate, silero_model.py:449-474 pad/truncate to 512)
        {
            float *xT = bufA;
            const int s = lane;
            const int i = st * kTile + s;
            const bool valid = s_valid[s] != 0;
            const int slot = s_slot[s];
            const bool dn = valid ? (p.denoise[slot] != 0) : false;
            const long long base = (long long)i * p.stride + (long long)frame * p.hop;
            bool bad = false;
#pragma unroll 4
            for (int it = 0; it < 8; ++it) {
                const int q = it * 16 + warp;  // float4 index 0..127
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (valid) {
                    const int k0 = 4 * q;
                    if (p.vec_ok && k0 + 4 <= flen) {
                        if (p.pcm == 0) {
                            const float4 t = __ldg(reinterpret_cast<const float4 *>(
                                reinterpret_cast<const float *>(p.audio) + base + k0));
                            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        } else {
                            const short4 t = __ldg(reinterpret_cast<const short4 *>(
                                reinterpret_cast<const short *>(p.audio) + base + k0));
                            v[0] = (float)t.x; v[1] = (float)t.y; v[2] = (float)t.z; v[3] = (float)t.w;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (k0 + e < flen) {
                                if (p.pcm == 0)
                                    v[e] = __ldg(reinterpret_cast<const float *>(p.audio) + base + k0 + e);
                                else
                                    v[e] = (float)__ldg(reinterpret_cast<const short *>(p.audio) + base + k0 + e);
                            }
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = v[e];
                        if (p.pcm == 1) x = __fdiv_rn(x, 32767.0f);
                        else if (p.pcm == 2) x = x * (1.0f / 32768.0f);
                        if (!isfinite(x)) bad = true;
                        if (dn && !(fabsf(x) > 0.01f)) x = 0.0f;
                        v[e] = x;
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) xT[(4 * q + e) * kTile + s] = v[e];
            }
            if (bad && p.status) atomicOr(&p.status[i], 1u);
        }
        __syncthreads();

        // ---- STFT: spec[n][t][item] = sum_k W[k][n] * x[128 t + k][item], n over 256 packed columns
        {
            const float *xT = bufA;
            float *magT = bufB;
            const int tn = tq;  // columns 4tn..4tn+3 = bins 2tn, 2tn+1
            float2 acc2[3][2][4];
#pragma unroll
            for (int t = 0; t < 3; ++t) zero_tile(acc2[t]);

            for (int ci = 0; ci < 16; ++ci) {
                const float *w = ring_wait(ring);
                const float *xk = xT + (ci * 16) * kTile + 4 * tm;
#pragma unroll 8
                for (int kk = 0; kk < 16; ++kk) {
                    const Dup4 wv = dup4(ld4(w + kk * 256 + 4 * tn));
#pragma unroll
                    for (int t = 0; t < 3; ++t) fma4x4(acc2[t], ld4(xk + (128 * t + kk) * kTile), wv);
                }
                fe_release(ring, tid);
            }
            // magnitude = sqrt(re^2 + im^2) with separately rounded squares (ONNX Pow, Pow, Add, Sqrt)
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                float v[4][4];
                unpack_tile(acc2[t], v);
                float m0[4], m1[4], m2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x = v[i][0], y = v[i][1], z = v[i][2], w = v[i][3];
                    if (tn == 0) {
                        m0[i] = sqrtf(__fmul_rn(x, x));                              // bin 0   (im == 0)
                        m2[i] = sqrtf(__fmul_rn(y, y));                              // bin 128 (im == 0)
                    } else {
                        m0[i] = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));  // bin 2tn
                        m2[i] = 0.f;
                    }
                    m1[i] = sqrtf(__fadd_rn(__fmul_rn(z, z), __fmul_rn(w, w)));      // bin 2tn+1
                }
                st4(magT + ((2 * tn) * 3 + t) * kTile + 4 * tm, make_float4(m0[0], m0[1], m0[2], m0[3]));
                st4(magT + ((2 * tn + 1) * 3 + t) * kTile + 4 * tm, make_float4(m1[0], m1[1], m1[2], m1[3]));
                if (tn == 0) st4(magT + (128 * 3 + t) * kTile + 4 * tm, make_float4(m2[0], m2[1], m2[2], m2[3]));
            }
        }
        __syncthreads();
        if (p.dbg && tile == 0) { block_copy_to_global(p.dbg, bufB, kDbgMag); }

        // ---- encoder.0: conv k3 s1 p1, 129 -> 128, T 3 -> 3, + ReLU.  split-K over 2 thread groups.
        {
            const float *magT = bufB;
            float *e0T = bufA;
            const int tn = tq & 31;   // outputs 4tn..4tn+3
            const int grp = tq >> 5;  // 0..1
            float2 acc2[3][2][4];
#pragma unroll
            for (int t = 0; t < 3; ++t) zero_tile(acc2[t]);

            for (int ci = 0; ci < 13; ++ci) {
                const float *w = ring_wait(ring);
                const int cbase = ci * 10 + 5 * grp;
                const int cnt = min(5, 129 - cbase);
                for (int cc = 0; cc < cnt; ++cc) {
                    const float *wl = w + ((5 * grp + cc) * 3) * 128 + 4 * tn;
                    const Dup4 w0 = dup4(ld4(wl)), w1 = dup4(ld4(wl + 128)), w2 = dup4(ld4(wl + 256));
                    const float *ar = magT + ((cbase + cc) * 3) * kTile + 4 * tm;
                    const float4 a0 = ld4(ar), a1 = ld4(ar + kTile), a2 = ld4(ar + 2 * kTile);
                    // out t: taps k with input time t+k-1 in [0,3)
                    fma4x4(acc2[0], a0, w1); fma4x4(acc2[0], a1, w2);
                    fma4x4(acc2[1], a0, w0); fma4x4(acc2[1], a1, w1); fma4x4(acc2[1], a2, w2);
                    fma4x4(acc2[2], a1, w0); fma4x4(acc2[2], a2, w1);
                }
                fe_release(ring, tid);
            }
            // reduce the two K halves through the output buffer, add bias, ReLU
            float v[3][4][4];
#pragma unroll
            for (int t = 0; t < 3; ++t) unpack_tile(acc2[t], v[t]);
            if (grp == 1) {
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int j = 0; j < 4; ++j) st4(e0T + ((4 * tn + j) * 3 + t) * kTile + 4 * tm, col4(v[t], j));
            }
            __syncthreads();
            if (grp == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float b = __ldg(p.b_fe + 4 * tn + j);
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        float *d = e0T + ((4 * tn + j) * 3 + t) * kTile + 4 * tm;
                        st4(d, bias_relu4(col4(v[t], j), ld4(d), b));
                    }
                }
            }
        }
        __syncthreads();
        if (p.dbg && tile == 0) { block_copy_to_global(p.dbg + kDbgMag, bufA, kDbgE0); }

        // ---- encoder.1: conv k3 s2 p1, 128 -> 64, T 3 -> 2, + ReLU.  split-K over 4 groups.
        {
            const float *e0T = bufA;
            float *e1T = bufB;
            const int tn = tq & 15;   // outputs 4tn..4tn+3 of 64
            const int grp = tq >> 4;  // 0..3
            float2 acc2[2][2][4];
            zero_tile(acc2[0]);
            zero_tile(acc2[1]);
            for (int ci = 0; ci < 8; ++ci) {
                const float *w = ring_wait(ring);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = ci * 16 + 4 * grp + cc;
                    const float *wl = w + ((4 * grp + cc) * 3) * 64 + 4 * tn;
                    const Dup4 w0 = dup4(ld4(wl)), w1 = dup4(ld4(wl + 64)), w2 = dup4(ld4(wl + 128));
                    const float *ar = e0T + (c * 3) * kTile + 4 * tm;
                    const float4 a0 = ld4(ar), a1 = ld4(ar + kTile), a2 = ld4(ar + 2 * kTile);
                    // out t' reads input time 2t'+k-1: t'=0 -> (k1,t0),(k2,t1); t'=1 -> (k0,t1),(k1,t2)
                    fma4x4(acc2[0], a0, w1); fma4x4(acc2[0], a1, w2);
                    fma4x4(acc2[1], a1, w0); fma4x4(acc2[1], a2, w1);
                }
                fe_release(ring, tid);
            }
            float v[2][4][4];
            unpack_tile(acc2[0], v[0]);
            unpack_tile(acc2[1], v[1]);
            for (int r = 1; r < 4; ++r) {
                if (grp == r) {
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float *d = e1T + ((4 * tn + j) * 2 + t) * kTile + 4 * tm;
                            st4(d, r > 1 ? add4(col4(v[t], j), ld4(d)) : col4(v[t], j));
                        }
                }
                __syncthreads();
            }
            if (grp == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float b = __ldg(p.b_fe + 128 + 4 * tn + j);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        float *d = e1T + ((4 * tn + j) * 2 + t) * kTile + 4 * tm;
                        st4(d, bias_relu4(col4(v[t], j), ld4(d), b));
                    }
                }
            }
        }
        __syncthreads();
        if (p.dbg && tile == 0) { block_copy_to_global(p.dbg + kDbgMag + kDbgE0, bufB, kDbgE1); }

        // ---- encoder.2: conv k3 s2 p1, 64 -> 64, T 2 -> 1, + ReLU (taps 1,2 see data).  split-K over 4.
        {
            const float *e1T = bufB;
            float *e2T = bufA;
            const int tn = tq & 15;
            const int grp = tq >> 4;
            float2 acc2[2][4];
            zero_tile(acc2);
            for (int ci = 0; ci < 2; ++ci) {
                const float *w = ring_wait(ring);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = ci * 32 + 8 * grp + cc;
                    const float *wl = w + ((8 * grp + cc) * 2) * 64 + 4 * tn;
                    const Dup4 w1 = dup4(ld4(wl)), w2 = dup4(ld4(wl + 64));
                    const float *ar = e1T + (c * 2) * kTile + 4 * tm;
                    fma4x4(acc2, ld4(ar), w1);
                    fma4x4(acc2, ld4(ar + kTile), w2);
                }
                fe_release(ring, tid);
            }
            float v[4][4];
            unpack_tile(acc2, v);
            for (int r = 1; r < 4; ++r) {
                if (grp == r) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float *d = e2T + (4 * tn + j) * kTile + 4 * tm;
                        st4(d, r > 1 ? add4(col4(v, j), ld4(d)) : col4(v, j));
                    }
                }
                __syncthreads();
            }
            if (grp == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float b = __ldg(p.b_fe + 192 + 4 * tn + j);
                    float *d = e2T + (4 * tn + j) * kTile + 4 * tm;
                    st4(d, bias_relu4(col4(v, j), ld4(d), b));
                }
            }
        }
        __syncthreads();
        if (p.dbg && tile == 0) { block_copy_to_global(p.dbg + kDbgMag + kDbgE0 + kDbgE1, bufA, kDbgE2); }

        // ---- encoder.3: conv k3 s1 p1, 64 -> 128, T 1 -> 1 (centre tap), + ReLU -> feat (HBM)
        {
            const float *e2T = bufA;
            float *scratch = bufB;
            const int tn = tq & 31;
            const int grp = tq >> 5;
            float2 acc2[2][4];
            zero_tile(acc2);
            for (int ci = 0; ci < 2; ++ci) {
                const float *w = ring_wait(ring);
#pragma unroll
                for (int cc = 0; cc < 16; ++cc) {
                    const int c = ci * 32 + 16 * grp + cc;
                    const Dup4 wv = dup4(ld4(w + (16 * grp + cc) * 128 + 4 * tn));
                    fma4x4(acc2, ld4(e2T + c * kTile + 4 * tm), wv);
                }
                fe_release(ring, tid);
            }
            float v[4][4];
            unpack_tile(acc2, v);
            if (grp == 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) st4(scratch + (4 * tn + j) * kTile + 4 * tm, col4(v, j));
            }
            __syncthreads();
            if (grp == 0) {
                float *fout = p.feat + (size_t)tile * (128 * kTile);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float b = __ldg(p.b_fe + 256 + 4 * tn + j);
                    const float4 r = bias_relu4(col4(v, j), ld4(scratch + (4 * tn + j) * kTile + 4 * tm), b);
                    st4(fout + (4 * tn + j) * kTile + 4 * tm, r);
                    if (p.dbg && tile == 0)
                        st4(p.dbg + kDbgMag + kDbgE0 + kDbgE1 + kDbgE2 + (4 * tn + j) * kTile + 4 * tm, r);
                }
            }
        }
        __syncthreads();  // bufA/bufB free for the next tile
    }

    // drain the kRingStages chunks that are always in flight
    for (int i = 0; i < kRingStages; ++i) {
        ring_wait(ring);
        ++ring.g;
    }
}

// =====================================================================================
// Recurrent part: LSTMCell(128) + decoder + start/end state machine, frames in order
// =====================================================================================
constexpr size_t kRecSmemBytes =
    (size_t)(kRingStages * kRingSlotFloats + 2 * 4096 /*xbuf*/ + 4096 /*hbuf*/) * 4 + 64 /*bars*/ +
    kTile * (4 * 4 + 2 * 8 + 8) + 64;

__global__ void __launch_bounds__(kThreads, 1) v5_recurrent_kernel(const V5Step p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring_buf = reinterpret_cast<float *>(smem_raw);
    float *xbuf = ring_buf + kRingStages * kRingSlotFloats;  // [2][128][32]
    float *hbuf = xbuf + 2 * 4096;                           // [128][32]
    uint64_t *bars = reinterpret_cast<uint64_t *>(hbuf + 4096);  // [4] ring + [2] feature
    uint64_t *xbars = bars + kRingStages;
    double *s_startp = reinterpret_cast<double *>(bars + 8);   // [32]
    double *s_endp = s_startp + kTile;                         // [32]
    int *s_slot = reinterpret_cast<int *>(s_endp + kTile);     // [32]
    int *s_nfr = s_slot + kTile;                               // [32]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int tm = tid & 7;   // items 4tm..4tm+3
    const int tn = tid >> 3;  // 0..63: hidden units 2tn, 2tn+1 (packed columns 8tn..8tn+7)
    const int st = blockIdx.x;

    if (tid < kTile) {
        const int i = st * kTile + tid;
        int slot = -1, nf = 0;
        if (i < p.n_streams) {
            slot = p.slots ? p.slots[i] : i;
            nf = p.n_frames ? p.n_frames[i] : p.max_frames;
            if (p.status && p.status[i] != 0u) nf = 0;  // NaN/Inf: the reference raises before any frame runs
            s_startp[tid] = p.start_p[slot];
            s_endp[tid] = p.end_p[slot];
        }
        s_slot[tid] = slot;
        s_nfr[tid] = nf;
    }
    __syncthreads();
    int tmax = 0;
    {
        int v = s_nfr[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        tmax = v;
    }
    if (tmax == 0) return;  // uniform

    WeightRing ring{ring_buf, bars, p.w_rec, 0u};
    if (tid == 0) {
        for (int i = 0; i < kRingStages; ++i) mbar_init(&bars[i], 1);
        mbar_init(&xbars[0], 1);
        mbar_init(&xbars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t i = 0; i < kRingStages; ++i) rec_ring_issue(ring, i);
        for (int j = 0; j < 2 && j < tmax; ++j) {
            mbar_arrive_expect_tx(&xbars[j], 4096 * 4u);
            bulk_g2s(xbuf + j * 4096, p.feat + ((size_t)j * p.n_stiles + st) * 4096, 4096 * 4u, &xbars[j]);
        }
    }

    // ---- resident state: h -> shared [unit][item], c -> registers
    for (int idx = tid; idx < 4096; idx += kThreads) {
        const int s = idx & 31, u = idx >> 5;
        const int slot = s_slot[s];
        hbuf[idx] = slot >= 0 ? p.h_state[state_at(u, slot)] : 0.f;
    }
    float creg[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int slot = s_slot[4 * tm + i];
#pragma unroll
        for (int u = 0; u < 2; ++u) creg[i][u] = slot >= 0 ? p.c_state[state_at(2 * tn + u, slot)] : 0.f;
    }
    // state machine words (warp 0, one lane per stream)
    int sm_active = 0, sm_sc = 0, sm_ec = 0, sm_ns = 1, sm_ne = 1;
    long long sm_f0 = 0;
    if (warp == 0 && s_slot[lane] >= 0) {
        const int slot = s_slot[lane];
        sm_active = p.sm_active[slot];
        sm_sc = p.sm_scount[slot];
        sm_ec = p.sm_ecount[slot];
        sm_ns = p.n_start[slot];
        sm_ne = p.n_end[slot];
        sm_f0 = p.frames_done[slot];
    }
    // gate biases for this thread's 8 packed columns
    const float4 bia0 = __ldg(reinterpret_cast<const float4 *>(p.b_rec + 8 * tn));
    const float4 bia1 = __ldg(reinterpret_cast<const float4 *>(p.b_rec + 8 * tn + 4));
    const float dec_b = __ldg(p.w_dec + 128);
    __syncthreads();

    for (int j = 0; j < tmax; ++j) {
        const float *xb = xbuf + (j & 1) * 4096;
        mbar_wait(&xbars[j & 1], (uint32_t)(j >> 1) & 1u);

        float2 acc2[2][2][4];  // [column half][item pair][column]
        zero_tile(acc2[0]);
        zero_tile(acc2[1]);

        for (int ci = 0; ci < kRecChunks; ++ci) {
            const float *w = ring_wait(ring);
            const float *abase = ((ci < 16) ? (xb + ci * 8 * kTile) : (hbuf + (ci - 16) * 8 * kTile)) + 4 * tm;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const float4 a = ld4(abase + kk * kTile);
                const Dup4 w0 = dup4(ld4(w + kk * 512 + 8 * tn)), w1 = dup4(ld4(w + kk * 512 + 8 * tn + 4));
                fma4x4(acc2[0], a, w0);
                fma4x4(acc2[1], a, w1);
            }
            __syncthreads();
            if (tid == 0) rec_ring_issue(ring, ring.g + kRingStages);
            ++ring.g;
        }
        // every thread is past its last read of xb / hbuf (barrier above)
        if (tid == 0 && j + 2 < tmax) {
            mbar_arrive_expect_tx(&xbars[j & 1], 4096 * 4u);
            bulk_g2s(xbuf + (j & 1) * 4096, p.feat + ((size_t)(j + 2) * p.n_stiles + st) * 4096, 4096 * 4u,
                     &xbars[j & 1]);
        }
        // LSTM cell (PyTorch LSTMCell == ONNX LSTM with both biases added)
        float g0[4][4], g1[4][4];
        unpack_tile(acc2[0], g0);
        unpack_tile(acc2[1], g1);
        const float bi[8] = {bia0.x, bia0.y, bia0.z, bia0.w, bia1.x, bia1.y, bia1.z, bia1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = 4 * tm + i;
            const bool live = j < s_nfr[row];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float (&g)[4][4] = u == 0 ? g0 : g1;
                const float ig = sigmoid_f(g[i][0] + bi[4 * u + 0]);
                const float fg = sigmoid_f(g[i][1] + bi[4 * u + 1]);
                const float gg = tanhf(g[i][2] + bi[4 * u + 2]);
                const float og = sigmoid_f(g[i][3] + bi[4 * u + 3]);
                const float cn = __fadd_rn(__fmul_rn(fg, creg[i][u]), __fmul_rn(ig, gg));
                const float hn = og * tanhf(cn);
                if (live) {
                    creg[i][u] = cn;
                    hbuf[(2 * tn + u) * kTile + row] = hn;
                }
            }
        }
        __syncthreads();
        // decoder: sigmoid(w . relu(h') + b), then the start/end state machine
        if (warp == 0 && j < s_nfr[lane]) {
            float a = 0.f;
#pragma unroll 8
            for (int u = 0; u < 128; ++u) a = fmaf(__ldg(p.w_dec + u), fmaxf(hbuf[u * kTile + lane], 0.f), a);
            const float prob = sigmoid_f(a + dec_b);
            const double pd = (double)prob;
            unsigned int fl = 0u;
            if (!sm_active) {
                if (pd >= s_startp[lane]) {
                    ++sm_sc;
                    if (sm_sc >= sm_ns && sm_ns <= 20) {  // deque(maxlen=20), silero_model.py:620-623
                        sm_active = 1; sm_sc = 0; sm_ec = 0; fl |= 1u;
                    }
                } else {
                    sm_sc = 0;
                }
            } else {
                fl |= 4u;
                if (pd < s_endp[lane]) {
                    ++sm_ec;
                    if (sm_ec >= sm_ne && sm_ne <= 100) {  // deque(maxlen=100), :625-628
                        sm_active = 0; sm_ec = 0; fl |= 2u;
                    }
                } else {
                    sm_ec = 0;
                }
            }
            const int i = st * kTile + lane;
            if (p.probs) p.probs[(size_t)i * p.max_frames + j] = prob;
            if (p.flags) p.flags[(size_t)i * p.max_frames + j] = (unsigned char)fl;
            if ((fl & 3u) && p.n_events) {
                EventRec *ev = reinterpret_cast<EventRec *>(p.events);
                for (unsigned int kind = 1u; kind <= 2u; kind <<= 1) {
                    if (fl & kind) {
                        const int at = atomicAdd(p.n_events, 1);
                        if (ev && at < p.max_events) {
                            ev[at].stream = i; ev[at].slot = s_slot[lane]; ev[at].frame = j;
                            ev[at].kind = (int)kind; ev[at].stream_frame = sm_f0 + j;
                        }
                    }
                }
            }
        }
        // next frame's GEMM only reads hbuf; the next write to it is 32 barriers away
    }

    // ---- write the resident state back
    if (p.commit) {
        __syncthreads();
        for (int idx = tid; idx < 4096; idx += kThreads) {
            const int s = idx & 31, u = idx >> 5;
            const int slot = s_slot[s];
            if (slot >= 0 && s_nfr[s] > 0) p.h_state[state_at(u, slot)] = hbuf[idx];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int slot = s_slot[4 * tm + i];
            if (slot >= 0 && s_nfr[4 * tm + i] > 0) {
#pragma unroll
                for (int u = 0; u < 2; ++u) p.c_state[state_at(2 * tn + u, slot)] = creg[i][u];
            }
        }
        if (warp == 0 && s_slot[lane] >= 0 && s_nfr[lane] > 0) {
            const int slot = s_slot[lane];
            p.sm_active[slot] = sm_active;
            p.sm_scount[slot] = sm_sc;
            p.sm_ecount[slot] = sm_ec;
            p.frames_done[slot] = sm_f0 + s_nfr[lane];
        }
    }
    for (int i = 0; i < kRingStages; ++i) {
        ring_wait(ring);
        ++ring.g;
    }
}

}  // namespace cvad
