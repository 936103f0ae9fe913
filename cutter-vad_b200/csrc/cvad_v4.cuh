// cvad_v4.cuh -- Silero VAD v4 (16 kHz branch of silero_vad.onnx) as two fused sm_100a kernels.
//
// Graph = SURVEY.md section 8a table "S4" (what onnxruntime runs for the reference at
// /root/reference/src/real_time_vad/core/silero_model.py:433 when model_version == V4):
//   reflect-pad 96|96 -> STFT conv k256 s64 (8 columns) -> |.| -> log(1 + 2^20 |.|) -> adaptive
//   normalisation (bin mean, reflect-pad 3, 7-tap filter, time mean) -> concat(mag, norm) 258 ch
//   -> first_layer (dw5+ReLU -> pw 258->16, + proj 258->16, ReLU) -> 1x1 s2 -> encoder.3 -> 1x1 s2
//   -> encoder.7 (identity residual) -> 1x1 s2 -> encoder.11 -> 1x1 -> feat[64]        (v4_frontend_kernel)
//   -> 2 x LSTM(64) (ONNX gate order i,o,f,c) -> ReLU -> 1x1 conv -> sigmoid -> state machine
//                                                                                       (v4_recurrent_kernel)
// 77% of the 689,640 MAC per frame are the STFT, which runs as a packed-column FFMA2 GEMM fed by
// the same bulk-copy weight ring as v5.  The front end works on tiles of 16 items (the 258x8
// spectrogram of 32 items would not fit shared memory); the recurrent kernel on 32 streams.
#pragma once
#include "cvad_v5.cuh"

namespace cvad {

constexpr int kV4Tile = 16;  // items per front-end tile

// ---- front-end weight stream (floats), one period = 24 chunks
//   [0, 65536)          stft    [k 256][n 256]   (same column packing as v5)          16 chunks x 4096
//   [65536, 75856)      first   [c 258][40] = dw_w[5] dw_b pad[2] pw[16] proj[16]     3 chunks x 3440 (86 ch)
//   [75856, 78368)  S0  c1T[16][16] c1_b[16] e3_dw[5][16] e3_dwb[16] e3_pwT[16][32] e3_pwb[32]
//                       e3_pjT[16][32] e3_pjb[32] c2T[32][32] c2_b[32]                 2512
//   [78368, 80672)  S1  e7_dw[5][32] e7_dwb[32] e7_pwT[32][32] e7_pwb[32] c3T[32][32] c3_b[32]   2304
//   [80672, 82976)  S2  e11_dw[5][32] e11_dwb[32] e11_pwT[32][64] e11_pwb[64]          2304
//   [82976, 85152)  S3  e11_pjT[32][64] e11_pjb[64] c4_b[64]                           2176
//   [85152, 89248)  S4  c4T[64][64]                                                    4096
//   [89248, 89296)      first_layer bias pw_b[16]+proj_b[16] summed [16], norm_filter[7], pad   (read with __ldg)
constexpr int kV4FeStreamFloats = 89296;
constexpr int kV4FeChunks = 24;
constexpr int kV4OffFirst = 65536, kV4OffS0 = 75856, kV4OffS1 = 78368, kV4OffS2 = 80672, kV4OffS3 = 82976,
              kV4OffS4 = 85152, kV4OffMisc = 89248;
// ---- recurrent weight stream: layer1 [k 128][256] then layer2 [k 128][256]; k<64 input, k>=64 hidden;
//      column = 4*unit + gate(i,o,f,c); 16 chunks x 4096.  Bias block [2][256] = Wb + Rb, dec_w[64], dec_b.
constexpr int kV4RecStreamFloats = 65536;
constexpr int kV4RecChunks = 16;

__device__ __forceinline__ void v4_fe_chunk(int ci, uint32_t &off, uint32_t &n) {
    if (ci < 16) { off = ci * 4096; n = 4096; }
    else if (ci < 19) { off = kV4OffFirst + (ci - 16) * 3440; n = 3440; }
    else if (ci == 19) { off = kV4OffS0; n = 2512; }
    else if (ci == 20) { off = kV4OffS1; n = 2304; }
    else if (ci == 21) { off = kV4OffS2; n = 2304; }
    else if (ci == 22) { off = kV4OffS3; n = 2176; }
    else { off = kV4OffS4; n = 4096; }
}

__device__ __forceinline__ void v4_fe_ring_issue(const WeightRing &r, uint32_t g) {
    uint32_t off, n;
    v4_fe_chunk(static_cast<int>(r.period ? r.first + g % r.period : g % kV4FeChunks), off, n);
    const uint32_t slot = g % kRingStages;
    mbar_arrive_expect_tx(&r.bars[slot], n * 4u);
    bulk_g2s(r.buf + slot * kRingSlotFloats, r.gsrc + off, n * 4u, &r.bars[slot]);
}

__device__ __forceinline__ void v4_fe_release(WeightRing &ring, int tid) {
    __syncthreads();
    if (tid == 0) v4_fe_ring_issue(ring, ring.g + kRingStages);
    ++ring.g;
}

__device__ __forceinline__ void v4_rec_ring_issue(const WeightRing &r, uint32_t g) {
    const uint32_t slot = g % kRingStages;
    const uint32_t ci = g % kV4RecChunks;
    mbar_arrive_expect_tx(&r.bars[slot], kRingSlotFloats * 4u);
    bulk_g2s(r.buf + slot * kRingSlotFloats, r.gsrc + ci * kRingSlotFloats, kRingSlotFloats * 4u, &r.bars[slot]);
}

// shared memory (floats): ring | bufX (xT[704][16] -> norm[129][8][16] -> small activations) | bufM (mag) | bufR
constexpr int kV4BufX = 129 * 8 * kV4Tile;  // 16512
constexpr int kV4BufM = 129 * 8 * kV4Tile;  // 16512
constexpr int kV4BufR = 16 * 8 * kV4Tile;   // 2048
constexpr size_t kV4FeSmemBytes =
    (size_t)(kRingStages * kRingSlotFloats + kV4BufX + kV4BufM + kV4BufR) * 4 + 64 + 256;

// debug dump (tile 0, frame 0): mag[129][8][16] norm[129][8][16] r3[16][8][16] r15[32][4][16] r27[32][2][16]
//                               r39[64][16] feat[64][16]
constexpr int kV4DbgMag = 16512, kV4DbgNorm = 16512, kV4DbgR3 = 2048, kV4DbgR15 = 2048, kV4DbgR27 = 1024,
              kV4DbgR39 = 1024, kV4DbgFeat = 1024;
constexpr int kV4DbgFloats = kV4DbgMag + kV4DbgNorm + kV4DbgR3 + kV4DbgR15 + kV4DbgR27 + kV4DbgR39 + kV4DbgFeat;

__device__ __forceinline__ float4 relu4(float4 a) {
    return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
}
__device__ __forceinline__ float4 fma4s(float w, const float4 &v, const float4 &acc) {
    return make_float4(fmaf(w, v.x, acc.x), fmaf(w, v.y, acc.y), fmaf(w, v.z, acc.z), fmaf(w, v.w, acc.w));
}

// ---- small layers: activations [ch][T][16], one work item = (out channel, time, 4-item group)
// depthwise conv k5 pad 2 + bias + ReLU; w is [5][C], b [C]
template <int C, int T>
__device__ __forceinline__ void v4_dw_relu(const float *in, const float *w, const float *b, float *out, int tid) {
    for (int wi = tid; wi < C * T * 4; wi += kThreads) {
        const int tm = wi & 3, t = (wi >> 2) % T, c = wi / (4 * T);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            const int ti = t + d - 2;
            if (ti >= 0 && ti < T) a = fma4s(w[d * C + c], ld4(in + (c * T + ti) * kV4Tile + 4 * tm), a);
        }
        const float bb = b[c];
        a.x += bb; a.y += bb; a.z += bb; a.w += bb;
        st4(out + (c * T + t) * kV4Tile + 4 * tm, relu4(a));
    }
}

// out[co][t] = ReLU( sum_ci w1T[ci][co] in1[ci][t*STRIDE] + b1[co]  (+ sum_ci w2T[ci][co] in2[ci][t*STRIDE] + b2[co])
//                    (+ res[co][t]) )
template <int CIN, int COUT, int TIN, int TOUT, int STRIDE>
__device__ __forceinline__ void v4_pw(const float *in1, const float *w1T, const float *b1, const float *in2,
                                      const float *w2T, const float *b2, const float *res, float *out,
                                      float *gout, int gstride, int tid, int gtstride = 0) {
    for (int wi = tid; wi < COUT * TOUT * 4; wi += kThreads) {
        const int tm = wi & 3, t = (wi >> 2) % TOUT, co = wi / (4 * TOUT);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int ci = 0; ci < CIN; ++ci)
            a = fma4s(w1T[ci * COUT + co], ld4(in1 + (ci * TIN + t * STRIDE) * kV4Tile + 4 * tm), a);
        float bb = b1[co];
        a.x += bb; a.y += bb; a.z += bb; a.w += bb;
        if (in2) {
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int ci = 0; ci < CIN; ++ci)
                r = fma4s(w2T[ci * COUT + co], ld4(in2 + (ci * TIN + t * STRIDE) * kV4Tile + 4 * tm), r);
            bb = b2[co];
            a.x += r.x + bb; a.y += r.y + bb; a.z += r.z + bb; a.w += r.w + bb;
        }
        if (res) {
            const float4 r = ld4(res + (co * TOUT + t) * kV4Tile + 4 * tm);
            a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
        }
        a = relu4(a);
        if (out) st4(out + (co * TOUT + t) * kV4Tile + 4 * tm, a);
        if (gout) st4(gout + t * gtstride + co * gstride + 4 * tm, a);
    }
}

// =====================================================================================
__global__ void __launch_bounds__(kThreads, 1) v4_frontend_kernel(const V5Step p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring_buf = reinterpret_cast<float *>(smem_raw);
    float *bufX = ring_buf + kRingStages * kRingSlotFloats;
    float *bufM = bufX + kV4BufX;
    float *bufR = bufM + kV4BufM;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bufR + kV4BufR);
    int *s_slot = reinterpret_cast<int *>(bars + 8);  // [16]
    int *s_valid = s_slot + 32;                       // [16]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    WeightRing ring{ring_buf, bars, p.w_fe, 0u};
    if (p.v4_mag) { ring.first = 16u; ring.period = 8u; }   // the STFT ran on the tensor cores: skip its 16 chunks
    if (tid == 0) {
        for (int i = 0; i < kRingStages; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (uint32_t i = 0; i < kRingStages; ++i) v4_fe_ring_issue(ring, i);

    const int n_ft = 2 * p.n_stiles;  // front-end tiles (16 items) per frame
    const int n_tiles = p.max_frames * n_ft;
    const int flen = p.frame_len < 512 ? p.frame_len : 512;
    const float *misc = p.w_fe + kV4OffMisc;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int frame = tile / n_ft;
        const int ft = tile - frame * n_ft;

        if (tid < kV4Tile) {
            const int i = ft * kV4Tile + tid;
            int valid = 0, slot = -1;
            if (i < p.n_streams) {
                slot = p.slots ? p.slots[i] : i;
                const int nf = p.n_frames ? p.n_frames[i] : p.max_frames;
                valid = frame < nf;
            }
            s_slot[tid] = slot;
            s_valid[tid] = valid;
        }
        const int any_valid = __syncthreads_or(tid < kV4Tile ? s_valid[tid] : 0);
        if (!any_valid) continue;

        if (p.v4_mag) {
            // |STFT| of this tile was produced by v4tc_stft_kernel (same [bin][t][item] layout as bufM)
            const float4 *src = reinterpret_cast<const float4 *>(p.v4_mag + (size_t)tile * kV4BufM);
            for (int idx = tid; idx < kV4BufM / 4; idx += kThreads) st4(bufM + 4 * idx, __ldg(src + idx));
        } else {
        // ---- frame loader into the middle of the reflect-padded buffer: xT[96 + k][item]
        {
            float *xT = bufX;
            const int s = lane & 15;
            const int half = lane >> 4;
            const int i = ft * kV4Tile + s;
            const bool valid = s_valid[s] != 0;
            const int slot = s_slot[s];
            const bool dn = valid ? (p.denoise[slot] != 0) : false;
            const long long base = (long long)i * p.stride + (long long)frame * p.hop;
            bool bad = false;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int q = it * 32 + warp * 2 + half;  // float4 index 0..127
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
                for (int e = 0; e < 4; ++e) xT[(96 + 4 * q + e) * kV4Tile + s] = v[e];
            }
            if (bad && p.status) atomicOr(&p.status[i], 1u);
        }
        __syncthreads();
        // reflect padding (no edge repeat): xp[j] = x[96-j], xp[608+j] = x[510-j], j < 96
        for (int idx = tid; idx < 192 * kV4Tile; idx += kThreads) {
            const int s = idx & 15, j = idx >> 4;
            if (j < 96) bufX[j * kV4Tile + s] = bufX[(192 - j) * kV4Tile + s];
            else bufX[(608 + (j - 96)) * kV4Tile + s] = bufX[(606 - (j - 96)) * kV4Tile + s];
        }
        __syncthreads();

        // ---- STFT: 8 columns, hop 64; thread = 4 items x 4 time columns x 4 packed outputs
        {
            const float *xT = bufX;
            const int tm = tid & 3, tn = (tid >> 2) & 63, th = tid >> 8;
            float2 acc2[4][2][4];
#pragma unroll
            for (int t = 0; t < 4; ++t) zero_tile(acc2[t]);
            for (int ci = 0; ci < 16; ++ci) {
                const float *w = ring_wait(ring);
                const float *xk = xT + (64 * 4 * th + ci * 16) * kV4Tile + 4 * tm;
#pragma unroll 4
                for (int kk = 0; kk < 16; ++kk) {
                    const Dup4 wv = dup4(ld4(w + kk * 256 + 4 * tn));
#pragma unroll
                    for (int t = 0; t < 4; ++t) fma4x4(acc2[t], ld4(xk + (64 * t + kk) * kV4Tile), wv);
                }
                v4_fe_release(ring, tid);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int tt = 4 * th + t;
                float v[4][4];
                unpack_tile(acc2[t], v);
                float m0[4], m1[4], m2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x = v[i][0], y = v[i][1], z = v[i][2], w = v[i][3];
                    if (tn == 0) {
                        m0[i] = sqrtf(__fmul_rn(x, x));
                        m2[i] = sqrtf(__fmul_rn(y, y));
                    } else {
                        m0[i] = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
                        m2[i] = 0.f;
                    }
                    m1[i] = sqrtf(__fadd_rn(__fmul_rn(z, z), __fmul_rn(w, w)));
                }
                st4(bufM + ((2 * tn) * 8 + tt) * kV4Tile + 4 * tm, make_float4(m0[0], m0[1], m0[2], m0[3]));
                st4(bufM + ((2 * tn + 1) * 8 + tt) * kV4Tile + 4 * tm, make_float4(m1[0], m1[1], m1[2], m1[3]));
                if (tn == 0) st4(bufM + (128 * 8 + tt) * kV4Tile + 4 * tm, make_float4(m2[0], m2[1], m2[2], m2[3]));
            }
        }
        }   // !p.v4_mag
        __syncthreads();  // mag complete; xT dead
        // ---- spect = log(1 + 2^20 mag)  (ONNX Mul, Add, Log)
        for (int idx = tid; idx < kV4BufM / 4; idx += kThreads) {
            const float4 m = ld4(bufM + 4 * idx);
            st4(bufX + 4 * idx, make_float4(logf(__fadd_rn(1.0f, __fmul_rn(m.x, 1048576.0f))),
                                            logf(__fadd_rn(1.0f, __fmul_rn(m.y, 1048576.0f))),
                                            logf(__fadd_rn(1.0f, __fmul_rn(m.z, 1048576.0f))),
                                            logf(__fadd_rn(1.0f, __fmul_rn(m.w, 1048576.0f)))));
        }
        __syncthreads();
        // ---- adaptive normalisation: mean over the 129 bins per (t, item) ...
        if (tid < 8 * kV4Tile) {
            float a = 0.f;
            for (int c = 0; c < 129; ++c) a += bufX[c * 8 * kV4Tile + tid];
            bufR[tid] = a / 129.0f;
        }
        __syncthreads();
        // ... reflect-pad 3, 7-tap filter, mean over time -> one scalar per item
        if (tid < kV4Tile) {
            float mp[14];
#pragma unroll
            for (int i = 0; i < 3; ++i) mp[i] = bufR[(3 - i) * kV4Tile + tid];
#pragma unroll
            for (int i = 0; i < 8; ++i) mp[3 + i] = bufR[i * kV4Tile + tid];
#pragma unroll
            for (int i = 0; i < 3; ++i) mp[11 + i] = bufR[(6 - i) * kV4Tile + tid];
            float mm = 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < 7; ++d) a += __ldg(misc + 16 + d) * mp[t + d];
                mm += a;
            }
            bufR[8 * kV4Tile + tid] = mm / 8.0f;
        }
        __syncthreads();
        for (int idx = tid; idx < kV4BufX / 4; idx += kThreads) {
            float4 v = ld4(bufX + 4 * idx);
            const float4 mm = ld4(bufR + 8 * kV4Tile + ((4 * idx) & 15));
            v.x -= mm.x; v.y -= mm.y; v.z -= mm.z; v.w -= mm.w;
            st4(bufX + 4 * idx, v);
        }
        __syncthreads();
        if (p.dbg && tile == 0) {
            block_copy_to_global(p.dbg, bufM, kV4DbgMag);
            block_copy_to_global(p.dbg + kV4DbgMag, bufX, kV4DbgNorm);
        }

        // ---- first_layer: x1 = [mag; norm] (258 ch): relu(dw5) -> pw 258->16, + proj 258->16; warp = K group
        {
            const int tm = lane & 3, t = lane >> 2;  // lane = (time 8, item group 4)
            float2 acc2[2][16];
#pragma unroll
            for (int o = 0; o < 16; ++o) { acc2[0][o] = make_float2(0.f, 0.f); acc2[1][o] = make_float2(0.f, 0.f); }
            for (int cj = 0; cj < 3; ++cj) {
                const float *w = ring_wait(ring);
                for (int cl = warp; cl < 86; cl += 16) {
                    const int c = cj * 86 + cl;
                    const float *row = (c < 129 ? bufM + c * 8 * kV4Tile : bufX + (c - 129) * 8 * kV4Tile) + 4 * tm;
                    const float *wr = w + cl * 40;
                    const float4 wd = ld4(wr), we = ld4(wr + 4);  // dw taps 0..3 | tap 4, bias, pad, pad
                    const float wdw[5] = {wd.x, wd.y, wd.z, wd.w, we.x};
                    float4 d = make_float4(0.f, 0.f, 0.f, 0.f), xc = d;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int ti = t + k - 2;
                        if (ti >= 0 && ti < 8) {
                            const float4 v = ld4(row + ti * kV4Tile);
                            d = fma4s(wdw[k], v, d);
                            if (k == 2) xc = v;
                        }
                    }
                    d.x += we.y; d.y += we.y; d.z += we.y; d.w += we.y;
                    d = relu4(d);
                    const float2 d0 = make_float2(d.x, d.y), d1 = make_float2(d.z, d.w);
                    const float2 x0 = make_float2(xc.x, xc.y), x1 = make_float2(xc.z, xc.w);
#pragma unroll
                    for (int o4 = 0; o4 < 4; ++o4) {
                        const Dup4 wp = dup4(ld4(wr + 8 + 4 * o4)), wj = dup4(ld4(wr + 24 + 4 * o4));
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int o = 4 * o4 + j;
                            acc2[0][o] = __ffma2_rn(d0, wp.d[j], acc2[0][o]);
                            acc2[1][o] = __ffma2_rn(d1, wp.d[j], acc2[1][o]);
                            acc2[0][o] = __ffma2_rn(x0, wj.d[j], acc2[0][o]);
                            acc2[1][o] = __ffma2_rn(x1, wj.d[j], acc2[1][o]);
                        }
                    }
                }
                v4_fe_release(ring, tid);
            }
            // reduce the 16 K groups (warps) through r3[o][t][item], then bias + ReLU
            float *r3 = bufR;
            for (int r = 0; r < 16; ++r) {
                if (warp == r) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) {
                        float *dst = r3 + (o * 8 + t) * kV4Tile + 4 * tm;
                        float4 v = make_float4(acc2[0][o].x, acc2[0][o].y, acc2[1][o].x, acc2[1][o].y);
                        if (r > 0) v = add4(v, ld4(dst));
                        if (r == 15) {
                            const float b = __ldg(misc + o);
                            v = relu4(make_float4(v.x + b, v.y + b, v.z + b, v.w + b));
                        }
                        st4(dst, v);
                    }
                }
                __syncthreads();
            }
        }
        if (p.dbg && tile == 0) block_copy_to_global(p.dbg + kV4DbgMag + kV4DbgNorm, bufR, kV4DbgR3);

        // ---- the small layers; activations now live in bufX / bufM (spectrogram is dead)
        float *r7 = bufX, *d3 = bufX + 1024, *r15 = bufX + 2048, *r19 = bufX + 4096, *d7 = bufX + 5120;
        // 8 kHz sub-model: the third 1x1 convolution has stride 1, so T = 2 survives to the LSTM
        const bool t2 = p.v4_t2 != 0;
        float *r27 = bufX + 6144, *r31 = bufX + 7168, *d11 = t2 ? bufX + 8192 : bufX + 7680,
              *r39 = t2 ? bufX + 9216 : bufX + 8192;
        {
            const float *w = ring_wait(ring);  // S0
            const float *c1T = w, *c1b = w + 256, *e3dw = w + 272, *e3dwb = w + 352, *e3pwT = w + 368,
                        *e3pwb = w + 880, *e3pjT = w + 912, *e3pjb = w + 1424, *c2T = w + 1456, *c2b = w + 2480;
            v4_pw<16, 16, 8, 4, 2>(bufR, c1T, c1b, nullptr, nullptr, nullptr, nullptr, r7, nullptr, 0, tid);
            __syncthreads();
            v4_dw_relu<16, 4>(r7, e3dw, e3dwb, d3, tid);
            __syncthreads();
            v4_pw<16, 32, 4, 4, 1>(d3, e3pwT, e3pwb, r7, e3pjT, e3pjb, nullptr, r15, nullptr, 0, tid);
            __syncthreads();
            v4_pw<32, 32, 4, 2, 2>(r15, c2T, c2b, nullptr, nullptr, nullptr, nullptr, r19, nullptr, 0, tid);
            v4_fe_release(ring, tid);
        }
        {
            const float *w = ring_wait(ring);  // S1
            const float *e7dw = w, *e7dwb = w + 160, *e7pwT = w + 192, *e7pwb = w + 1216, *c3T = w + 1248,
                        *c3b = w + 2272;
            v4_dw_relu<32, 2>(r19, e7dw, e7dwb, d7, tid);
            __syncthreads();
            v4_pw<32, 32, 2, 2, 1>(d7, e7pwT, e7pwb, nullptr, nullptr, nullptr, r19, r27, nullptr, 0, tid);
            __syncthreads();
            if (t2) v4_pw<32, 32, 2, 2, 1>(r27, c3T, c3b, nullptr, nullptr, nullptr, nullptr, r31, nullptr, 0, tid);
            else v4_pw<32, 32, 2, 1, 2>(r27, c3T, c3b, nullptr, nullptr, nullptr, nullptr, r31, nullptr, 0, tid);
            v4_fe_release(ring, tid);
        }
        if (p.dbg && tile == 0) {
            block_copy_to_global(p.dbg + kV4DbgMag + kV4DbgNorm + kV4DbgR3, r15, kV4DbgR15);
            block_copy_to_global(p.dbg + kV4DbgMag + kV4DbgNorm + kV4DbgR3 + kV4DbgR15, r27, kV4DbgR27);
        }
        {
            const float *w2 = ring_wait(ring);  // S2 (kept while S3 is consumed)
            const float *e11dw = w2, *e11dwb = w2 + 160, *e11pwT = w2 + 192, *e11pwb = w2 + 2240;
            if (t2) v4_dw_relu<32, 2>(r31, e11dw, e11dwb, d11, tid);
            else v4_dw_relu<32, 1>(r31, e11dw, e11dwb, d11, tid);
            __syncthreads();
            // S3 sits in the next ring slot
            const uint32_t slot3 = (ring.g + 1) % kRingStages;
            mbar_wait(&ring.bars[slot3], ((ring.g + 1) / kRingStages) & 1u);
            const float *w3 = ring.buf + slot3 * kRingSlotFloats;
            const float *e11pjT = w3, *e11pjb = w3 + 2048;
            if (t2) v4_pw<32, 64, 2, 2, 1>(d11, e11pwT, e11pwb, r31, e11pjT, e11pjb, nullptr, r39, nullptr, 0, tid);
            else v4_pw<32, 64, 1, 1, 1>(d11, e11pwT, e11pwb, r31, e11pjT, e11pjb, nullptr, r39, nullptr, 0, tid);
            v4_fe_release(ring, tid);  // S2
            v4_fe_release(ring, tid);  // S3
        }
        if (p.dbg && tile == 0) block_copy_to_global(
            p.dbg + kV4DbgMag + kV4DbgNorm + kV4DbgR3 + kV4DbgR15 + kV4DbgR27, r39, kV4DbgR39);
        {
            const float *w = ring_wait(ring);  // S4: c4T[64][64]; its bias travels in S3's global copy
            const float *c4b = p.w_fe + kV4OffS3 + 2112;
            // feat is tile-transposed for the recurrent kernel: [frame (x 2 time steps at 8 kHz)][rtile][64][32]
            float *gout = p.feat + ((size_t)frame * (t2 ? 2 : 1) * p.n_stiles + (ft >> 1)) * (64 * kTile) + (ft & 1) * kV4Tile;
            float *dbgf = (p.dbg && tile == 0)
                              ? p.dbg + kV4DbgMag + kV4DbgNorm + kV4DbgR3 + kV4DbgR15 + kV4DbgR27 + kV4DbgR39
                              : nullptr;
            if (t2) v4_pw<64, 64, 2, 2, 1>(r39, w, c4b, nullptr, nullptr, nullptr, nullptr, nullptr, gout, kTile, tid,
                                           p.n_stiles * 64 * kTile);
            else v4_pw<64, 64, 1, 1, 1>(r39, w, c4b, nullptr, nullptr, nullptr, nullptr, dbgf ? bufM : nullptr, gout,
                                        kTile, tid);
            v4_fe_release(ring, tid);
            if (dbgf) block_copy_to_global(dbgf, bufM, kV4DbgFeat);
        }
        __syncthreads();
    }
    for (int i = 0; i < kRingStages; ++i) {
        ring_wait(ring);
        ++ring.g;
    }
}

// =====================================================================================
// Recurrent part: 2 x LSTM(64) + decoder + state machine.  h/c rows 0..63 = layer 1, 64..127 = layer 2.
// =====================================================================================
constexpr size_t kV4RecSmemBytes =
    (size_t)(kRingStages * kRingSlotFloats + 2 * 2048 /*xbuf*/ + 2 * 2048 /*h1,h2*/) * 4 + 64 +
    kTile * (4 * 4 + 2 * 8 + 8) + 64;

__global__ void __launch_bounds__(kThreads, 1) v4_recurrent_kernel(const V5Step p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring_buf = reinterpret_cast<float *>(smem_raw);
    float *xbuf = ring_buf + kRingStages * kRingSlotFloats;  // [2][64][32]
    float *hbuf = xbuf + 2 * 2048;                           // [2 layers][64][32]
    uint64_t *bars = reinterpret_cast<uint64_t *>(hbuf + 2 * 2048);
    uint64_t *xbars = bars + kRingStages;
    double *s_startp = reinterpret_cast<double *>(bars + 8);
    double *s_endp = s_startp + kTile;
    int *s_slot = reinterpret_cast<int *>(s_endp + kTile);
    int *s_nfr = s_slot + kTile;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int tm = tid & 7;   // items 4tm..4tm+3
    const int tn = tid >> 3;  // hidden unit 0..63 (packed columns 4tn..4tn+3 = i,o,f,c)
    const int st = blockIdx.x;

    if (tid < kTile) {
        const int i = st * kTile + tid;
        int slot = -1, nf = 0;
        if (i < p.n_streams) {
            slot = p.slots ? p.slots[i] : i;
            nf = p.n_frames ? p.n_frames[i] : p.max_frames;
            if (p.status && p.status[i] != 0u) nf = 0;
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
    if (tmax == 0) return;

    WeightRing ring{ring_buf, bars, p.w_rec, 0u};
    if (tid == 0) {
        for (int i = 0; i < kRingStages; ++i) mbar_init(&bars[i], 1);
        mbar_init(&xbars[0], 1);
        mbar_init(&xbars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    // 8 kHz sub-model: every frame is two LSTM time steps ("virtual frames" jj = 2 j + t); the decoder output of
    // the frame is the mean of the two sigmoids (ONNX ReduceMean over T)
    const int T2 = p.v4_t2 ? 2 : 1;
    const int vmax = tmax * T2;
    if (tid == 0) {
        for (uint32_t i = 0; i < kRingStages; ++i) v4_rec_ring_issue(ring, i);
        for (int j = 0; j < 2 && j < vmax; ++j) {
            mbar_arrive_expect_tx(&xbars[j], 2048 * 4u);
            bulk_g2s(xbuf + j * 2048, p.feat + ((size_t)j * p.n_stiles + st) * 2048, 2048 * 4u, &xbars[j]);
        }
    }
    for (int idx = tid; idx < 4096; idx += kThreads) {
        const int s = idx & 31, u = idx >> 5;  // u = layer*64 + unit
        const int slot = s_slot[s];
        hbuf[idx] = slot >= 0 ? p.h_state[state_at(u, slot)] : 0.f;
    }
    float creg[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int slot = s_slot[4 * tm + i];
#pragma unroll
        for (int l = 0; l < 2; ++l) creg[l][i] = slot >= 0 ? p.c_state[state_at(l * 64 + tn, slot)] : 0.f;
    }
    int sm_active = 0, sm_sc = 0, sm_ec = 0, sm_ns = 1, sm_ne = 1;
    long long sm_f0 = 0;
    if (warp == 0 && s_slot[lane] >= 0) {
        const int slot = s_slot[lane];
        sm_active = p.sm_active[slot]; sm_sc = p.sm_scount[slot]; sm_ec = p.sm_ecount[slot];
        sm_ns = p.n_start[slot]; sm_ne = p.n_end[slot]; sm_f0 = p.frames_done[slot];
    }
    const float4 bias1 = __ldg(reinterpret_cast<const float4 *>(p.b_rec + 4 * tn));
    const float4 bias2 = __ldg(reinterpret_cast<const float4 *>(p.b_rec + 256 + 4 * tn));
    const float dec_b = __ldg(p.w_dec + 64);
    __syncthreads();

    float psum = 0.f;
    for (int jj = 0; jj < vmax; ++jj) {
        const int j = jj / T2;
        const float *xb = xbuf + (jj & 1) * 2048;
        mbar_wait(&xbars[jj & 1], (uint32_t)(jj >> 1) & 1u);
#pragma unroll
        for (int layer = 0; layer < 2; ++layer) {
            float2 acc2[2][4];
            zero_tile(acc2);
            const float *in_lo = layer == 0 ? xb : hbuf;             // k < 64
            const float *in_hi = layer == 0 ? hbuf : hbuf + 2048;     // k >= 64: this layer's own h
            for (int ci = 0; ci < 8; ++ci) {
                const float *w = ring_wait(ring);
                const float *abase = ((ci < 4) ? (in_lo + ci * 16 * kTile) : (in_hi + (ci - 4) * 16 * kTile)) + 4 * tm;
#pragma unroll
                for (int kk = 0; kk < 16; ++kk)
                    fma4x4(acc2, ld4(abase + kk * kTile), dup4(ld4(w + kk * 256 + 4 * tn)));
                __syncthreads();
                if (tid == 0) v4_rec_ring_issue(ring, ring.g + kRingStages);
                ++ring.g;
            }
            if (layer == 1 && tid == 0 && jj + 2 < vmax) {
                mbar_arrive_expect_tx(&xbars[jj & 1], 2048 * 4u);
                bulk_g2s(xbuf + (jj & 1) * 2048, p.feat + ((size_t)(jj + 2) * p.n_stiles + st) * 2048, 2048 * 4u,
                         &xbars[jj & 1]);
            }
            float g[4][4];
            unpack_tile(acc2, g);
            const float4 bi = layer == 0 ? bias1 : bias2;
            float *hdst = hbuf + layer * 2048 + tn * kTile;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = 4 * tm + i;
                const float ig = sigmoid_f(g[i][0] + bi.x);
                const float og = sigmoid_f(g[i][1] + bi.y);
                const float fg = sigmoid_f(g[i][2] + bi.z);
                const float cg = tanhf(g[i][3] + bi.w);
                const float cn = __fadd_rn(__fmul_rn(fg, creg[layer][i]), __fmul_rn(ig, cg));
                const float hn = og * tanhf(cn);
                if (j < s_nfr[row]) {
                    creg[layer][i] = cn;
                    hdst[row] = hn;
                }
            }
            __syncthreads();
        }
        if (warp == 0 && j < s_nfr[lane]) {
            float a = 0.f;
#pragma unroll 8
            for (int u = 0; u < 64; ++u) a = fmaf(__ldg(p.w_dec + u), fmaxf(hbuf[2048 + u * kTile + lane], 0.f), a);
            float prob = sigmoid_f(a + dec_b);
            if (T2 == 2) {
                if ((jj & 1) == 0) { psum = prob; continue; }   // first time step: wait for the second
                prob = (psum + prob) * 0.5f;
            }
            const double pd = (double)prob;
            unsigned int fl = 0u;
            if (!sm_active) {
                if (pd >= s_startp[lane]) {
                    ++sm_sc;
                    if (sm_sc >= sm_ns && sm_ns <= 20) { sm_active = 1; sm_sc = 0; sm_ec = 0; fl |= 1u; }
                } else {
                    sm_sc = 0;
                }
            } else {
                fl |= 4u;
                if (pd < s_endp[lane]) {
                    ++sm_ec;
                    if (sm_ec >= sm_ne && sm_ne <= 100) { sm_active = 0; sm_ec = 0; fl |= 2u; }
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
    }

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
                for (int l = 0; l < 2; ++l) p.c_state[state_at(l * 64 + tn, slot)] = creg[l][i];
            }
        }
        if (warp == 0 && s_slot[lane] >= 0 && s_nfr[lane] > 0) {
            const int slot = s_slot[lane];
            p.sm_active[slot] = sm_active; p.sm_scount[slot] = sm_sc; p.sm_ecount[slot] = sm_ec;
            p.frames_done[slot] = sm_f0 + s_nfr[lane];
        }
    }
    for (int i = 0; i < kRingStages; ++i) {
        ring_wait(ring);
        ++ring.g;
    }
}

}  // namespace cvad
