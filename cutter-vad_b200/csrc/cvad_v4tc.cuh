// cvad_v4tc.cuh -- the STFT of Silero VAD v4 on the tcgen05 tensor cores.
//
// 77 % of v4's 689,640 MAC per frame are the STFT (8 columns x 258 rows x 256 taps, SURVEY.md 8a "S4"); the rest
// of the front end is depthwise / 1x1 layers of 16..64 channels that stay on the CUDA cores (cvad_v4.cuh).  This
// kernel computes |STFT| of one tile of 16 (stream, frame) items with the same machinery as cvad_v5tc.cuh
// (three-way BF16 split, six products per MAC, FP32 accumulators in TMEM, weight tiles streamed by cp.async.bulk)
// and writes the magnitude tile [129][8][16] to HBM, where v4_frontend_kernel picks it up (V5Step::v4_mag).
//
// The hop (64 samples) equals one 64-element K block, so the sliding windows need no im2col copy: the
// reflect-padded frame (704 samples) is stored as 11 segments of 64 samples, B operand rows = segment*16 + item,
// and weight K block kb multiplies rows of segments kb .. kb+7: ONE MMA of N = 8 columns x 16 items = 128.
#pragma once
#include "cvad_resample.cuh"
#include "cvad_v5tc.cuh"

namespace cvad {
namespace tc5 {

constexpr int kV4tcRing = 8;
constexpr uint32_t kV4tcPart = 11 * 16 * 128;          // one part of the padded-audio operand: 176 rows x 128 B
constexpr size_t kV4StftStreamBytes = 24 * 16384;      // blk 0..1 x kb 0..3 x part 0..2 (same row packing as v5's STFT)
constexpr size_t kV4tcSmem = 1024 + 3 * (size_t)kV4tcPart + (size_t)kV4tcRing * kSlotBytes + (2 * kV4tcRing + 2) * 8 + 16 +
                             2 * 16 * 4 + 64;
constexpr int kV4MagTile = 129 * 8 * 16;               // floats per tile

// CORR = true (CVAD_MATH_FFT, v4's default): the STFT proper comes from v4_stft_fft_kernel (cvad_fftk.cuh: exact
// Hann x DFT-256 basis, double precision, float32 re / im in p.v4_fft); this kernel adds what the exact basis lacks --
// the product with delta = (the file's float32 basis) - (Hann x DFT), |delta| <= 7.7e-8 -- and forms the magnitude.
// delta * 2^24 is O(1), so ONE BF16 product per MAC (8 weight tiles instead of 24, bf16(x) only) gives the correction
// to 0.4 % of itself, i.e. to ~1e-9 of the frame's scale.
constexpr float kV4CorrScale = 16777216.0f;            // 2^24
constexpr size_t kV4CorrStreamBytes = 8 * 16384;       // blk 0..1 x kb 0..3, one part

template <bool CORR>
__global__ void __launch_bounds__(kThreadsTC, 1) v4tc_stft_kernel(const V5Step p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *aud = base;
    unsigned char *ring_buf = aud + 3 * kV4tcPart;
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring_buf + kV4tcRing * kSlotBytes);
    uint64_t *full = bars, *empty = bars + kV4tcRing, *act_ready = bars + 2 * kV4tcRing, *acc_ready = act_ready + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_ready + 1);
    int *s_slot = reinterpret_cast<int *>(tmem_slot + 4);
    int *s_valid = s_slot + 16;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kV4tcRing; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(act_ready, 1);
        mbar_init(acc_ready, 1);
        mbar_fence_init();
    }
    if (warp == kProducerWarp) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (*tmem_slot != 0u) __trap();

    const int n_ft = 2 * p.n_stiles;                   // 16-item tiles per frame (same indexing as v4_frontend_kernel)
    const int n_tiles = p.max_frames * n_ft;
    const uint32_t aud_s = smem_u32(aud), ring_s = smem_u32(ring_buf);

    // warp-uniform liveness of tile (frame, ft): lanes 0..15 look at the tile's 16 streams
    auto tile_live16 = [&](int frame, int ft, int *slot_out, int *valid_out) {
        const int i = ft * 16 + (lane & 15);
        int valid = 0, slot = -1;
        if (i < p.n_streams) {
            slot = p.slots ? p.slots[i] : i;
            valid = frame < (p.n_frames ? p.n_frames[i] : p.max_frames);
        }
        if (slot_out) *slot_out = slot;
        if (valid_out) *valid_out = valid;
        return __any_sync(0xffffffffu, valid) != 0;
    };

    if (warp == kProducerWarp) {
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int frame = tile / n_ft, ft = tile - frame * n_ft;
            if (!tile_live16(frame, ft, nullptr, nullptr)) continue;
            if (lane == 0) {
                for (int s = 0; s < (CORR ? 8 : 24); ++s, ++g) {
                    const uint32_t slot = g % kV4tcRing;
                    mbar_wait(&empty[slot], ((g / kV4tcRing) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&full[slot], kSlotBytes);
                    bulk_g2s(ring_buf + slot * kSlotBytes, p.w_fe_tc + (size_t)s * kSlotBytes, kSlotBytes, &full[slot]);
                }
            }
            __syncwarp();
        }
    } else if (warp == kMmaWarp) {
        uint32_t g = 0, act_phase = 0;
        const uint32_t idesc = tc::idesc_bf16_f32(128, 128);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int frame = tile / n_ft, ft = tile - frame * n_ft;
            if (!tile_live16(frame, ft, nullptr, nullptr)) continue;
            mbar_wait(act_ready, act_phase); act_phase ^= 1u;
            tc::fence_after_sync();
            for (int blk = 0; blk < 2; ++blk)
                for (int kb = 0; kb < 4; ++kb)
                    for (int wp = 0; wp < (CORR ? 1 : 3); ++wp, ++g) {
                        const uint32_t slot = g % kV4tcRing;
                        mbar_wait(&full[slot], (g / kV4tcRing) & 1u);
                        tc::fence_after_sync();
                        if (tc::elect_one()) {
                            // column t of the STFT = samples 64 t .. 64 t + 255 = segments t .. t+3: K block kb <-> segment t + kb
                            if (CORR) {
                                const uint64_t ad = tc::smem_desc_sw128(ring_s + slot * kSlotBytes);
                                const uint64_t bd = tc::smem_desc_sw128(aud_s + kb * 2048u);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    tc::mma_bf16(blk * 128u, ad + ks * 2, bd + ks * 2, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
                            } else {
                                issue_split(wp, ring_s + slot * kSlotBytes, aud_s + kb * 2048u, kV4tcPart, blk * 128u,
                                            256u + blk * 128u, idesc, kb == 0 && wp == 0);
                            }
                            tc::mma_commit(&empty[slot]);
                        }
                        __syncwarp();
                    }
            if (tc::elect_one()) tc::mma_commit(acc_ready);
            __syncwarp();
        }
    } else {
        const int q = warp & 3, cg = warp >> 2;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        uint32_t acc_phase = 0;
        const int flen = p.frame_len < 512 ? p.frame_len : 512;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int frame = tile / n_ft, ft = tile - frame * n_ft;
            int my_slot, my_valid;
            if (!tile_live16(frame, ft, &my_slot, &my_valid)) continue;
            if (warp == 0 && lane < 16) { s_slot[lane] = my_slot; s_valid[lane] = my_valid; }
            epi_bar();
            // ---- loader: frame (zero-padded / truncated to 512) -> gate -> reflect-pad 96|96 -> 3 BF16 parts.
            //      Work unit = 8 consecutive samples of the PADDED frame of one item (88 units per item).
            for (int unit = tid; unit < 16 * 88; unit += kEpiThreads) {
                const int s = unit / 88, c8 = unit - s * 88;
                const int i = ft * 16 + s;
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.f;
                if (s_valid[s]) {
                    const bool dn = p.denoise[s_slot[s]] != 0;
                    const long long b0 = (long long)i * p.stride + (long long)frame * p.hop;
                    bool bad = false;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int P = 8 * c8 + e;                                  // padded position 0..703
                        const int k = P < 96 ? 96 - P : (P < 608 ? P - 96 : 1118 - P);   // reflect, no edge repeat
                        float x = 0.f;
                        if (k < flen) {
                            x = p.pcm == 0 ? __ldg(reinterpret_cast<const float *>(p.audio) + b0 + k)
                                           : (float)__ldg(reinterpret_cast<const short *>(p.audio) + b0 + k);
                            if (p.pcm == 1) x = __fdiv_rn(x, 32767.0f);
                            else if (p.pcm == 2) x = x * (1.0f / 32768.0f);
                            if (!isfinite(x)) bad = true;
                            if (dn && !(fabsf(x) > 0.01f)) x = 0.0f;
                        }
                        v[e] = x;
                    }
                    if (bad && p.status) atomicOr(&p.status[i], 1u);
                }
                uint32_t w[3][4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split3x2(v[2 * e], v[2 * e + 1], w[0][e], w[1][e], w[2][e]);
                const uint32_t seg = (uint32_t)c8 >> 3, kk = ((uint32_t)c8 & 7u) * 8u;
                const uint32_t off = tc::sw128_offset(seg * 16u + (uint32_t)s, kk, 176u);
#pragma unroll
                for (int part = 0; part < (CORR ? 1 : 3); ++part)
                    *reinterpret_cast<uint4 *>(aud + part * kV4tcPart + off) = make_uint4(w[part][0], w[part][1], w[part][2], w[part][3]);
            }
            tc::fence_async_smem();
            epi_bar();
            if (tid == 0) mbar_arrive(act_ready);

            // ---- magnitude (ONNX Slice, Pow, Pow, Add, Sqrt) -> HBM tile [bin][t][item]
            mbar_wait(acc_ready, acc_phase); acc_phase ^= 1u;
            tc::fence_after_sync();
            {
                const int b = 32 * q + lane;
                float *mout = p.v4_mag + (size_t)tile * kV4MagTile;
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    const int c0 = cg * 32 + ch * 8;           // column = t*16 + item
                    float mr[8], mi[8], cr[8], ci[8], mag[8];
                    tmem_ld8(lane_addr + c0, mr);
                    tmem_ld8(lane_addr + 128 + c0, mi);
                    if (CORR) {
                        // the exact-basis STFT of this tile: [col][re | im][132]; lanes = bins, so every load is one coalesced row
                        const float *fin = p.v4_fft + (size_t)tile * (128 * 2 * 132) + (size_t)c0 * (2 * 132);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            cr[e] = __ldg(fin + e * 264 + b);
                            ci[e] = __ldg(fin + e * 264 + (b == 0 ? 128 : 132 + b));   // row 0 of the second block carries re[128]
                        }
                    } else {
                        tmem_ld8(lane_addr + 256 + c0, cr);
                        tmem_ld8(lane_addr + 384 + c0, ci);
                    }
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float re = CORR ? fmaf(mr[e], 1.0f / kV4CorrScale, cr[e]) : mr[e] + cr[e];
                        float im = CORR ? fmaf(mi[e], 1.0f / kV4CorrScale, ci[e]) : mi[e] + ci[e];
                        if (b == 0) {
                            mout[128 * 128 + c0 + e] = sqrtf(__fmul_rn(im, im));   // bin 128 rides in row 0 of the im block
                            im = 0.f;
                        }
                        mag[e] = sqrtf(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
                    }
                    float4 *dst = reinterpret_cast<float4 *>(mout + b * 128 + c0);
                    dst[0] = make_float4(mag[0], mag[1], mag[2], mag[3]);
                    dst[1] = make_float4(mag[4], mag[5], mag[6], mag[7]);
                }
            }
            tc::fence_before_sync();
            epi_bar();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kProducerWarp) tc::tmem_dealloc(0u, 512);
}

// =====================================================================================
// Resampler on the tensor cores: y[512] = R x[n_in] per frame-sized chunk (cvad_resample.cuh has the operator and
// the FP32 build).  A = R (4 blocks of 128 output samples x 64-sample K blocks, BF16x3 tiles streamed in the
// order kb-major / block / part), B = the chunk's K block for a tile of 64 streams (double-buffered: the loader
// converts K block kb+1 while the MMAs of kb run), D = 4 x 64 columns (+ the correction accumulators) in TMEM.
// Output: float32 16 kHz frames in HBM, exactly where resample_kernel puts them.
// =====================================================================================
constexpr int kRsTcTile = 64;
constexpr int kRsTcRing = 8;
constexpr uint32_t kRsTcBPart = kRsTcTile * 128;                 // one part of one B buffer: 64 rows x 128 B
constexpr size_t kRsTcSmem = 1024 + 2 * 3 * (size_t)kRsTcBPart + (size_t)kRsTcRing * kSlotBytes + (2 * kRsTcRing + 5) * 8 +
                             16 + 3 * kRsTcTile * 4 + 64;

// H16 (CVAD_MATH_TC16): the operator as two FP16 parts (scaled by a power of two on the host, inverse in `inv_w`), the
// chunk scaled per stream from its own maximum (a first pass over the chunk, which the second then finds in L2):
// three products per MAC instead of six, same FP32-level accuracy (cvad_v5tc.cuh, "FP16 two-way split").
template <bool H16>
__global__ void __launch_bounds__(kThreadsTC, 1) resample_tc_kernel(const ResampleStep p, const unsigned char *r_tiles,
                                                                    const float inv_w) {
    constexpr int NP = H16 ? 2 : 3;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *bbuf = base;                                   // [2][3 parts][64 rows][128 B]
    unsigned char *ring_buf = bbuf + 2 * 3 * kRsTcBPart;
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring_buf + kRsTcRing * kSlotBytes);
    uint64_t *full = bars, *empty = bars + kRsTcRing, *b_full = bars + 2 * kRsTcRing, *b_empty = b_full + 2,
             *acc_ready = b_empty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_ready + 1);
    int *s_gi = reinterpret_cast<int *>(tmem_slot + 4);           // [64] stream index of each tile row, or -1
    int *s_valid = s_gi + kRsTcTile;
    float *s_inv = reinterpret_cast<float *>(s_valid + kRsTcTile);   // H16: [64] 1 / (stream scale)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kRsTcRing; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(acc_ready, 1);
        mbar_fence_init();
    }
    if (warp == kProducerWarp) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (*tmem_slot != 0u) __trap();
    // chained steps launch the model kernel with programmatic stream serialization: its grid may be scheduled now (it sets
    // up, prefetches weights and blocks in griddepcontrol.wait until this grid has completed)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int n_items = p.count ? *p.count : p.n_streams;
    const int n_st = (n_items + kRsTcTile - 1) / kRsTcTile;
    const int n_kb = p.n_in / 64;
    // small batches: the four output blocks of a tile are shared among `osplit` CTAs (each reads the tile's audio)
    const int osplit = p.osplit > 0 ? p.osplit : 1, nblk = 4 / osplit;
    const int n_tiles = p.max_frames * n_st * osplit;
    const uint32_t b_s = smem_u32(bbuf), ring_s = smem_u32(ring_buf);

    // warp-uniform: does tile (frame, st) hold a live (stream, frame) item?  (every warp evaluates all 64 rows)
    auto tile_live64 = [&](int frame, int st, int *gi2, int *valid2) {
        bool any = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = st * kRsTcTile + h * 32 + lane;
            int valid = 0, gi = -1;
            if (j < n_items) {
                gi = p.list ? p.list[j] : j;
                valid = frame < (p.n_frames ? p.n_frames[gi] : p.max_frames);
            }
            if (gi2) { gi2[h] = gi; valid2[h] = valid; }
            any = any || __any_sync(0xffffffffu, valid);
        }
        return any;
    };

    if (warp == kProducerWarp) {
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int osp = tile % osplit, ft = tile / osplit;
            const int frame = ft / n_st, st = ft - frame * n_st;
            if (!tile_live64(frame, st, nullptr, nullptr)) continue;
            if (lane == 0) {
                for (int kb = 0; kb < n_kb; ++kb)
                    for (int s = osp * nblk * NP; s < (osp + 1) * nblk * NP; ++s, ++g) {
                        const uint32_t slot = g % kRsTcRing;
                        mbar_wait(&empty[slot], ((g / kRsTcRing) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(&full[slot], kSlotBytes);
                        bulk_g2s(ring_buf + slot * kSlotBytes, r_tiles + (size_t)(kb * 4 * NP + s) * kSlotBytes, kSlotBytes, &full[slot]);
                    }
            }
            __syncwarp();
        }
    } else if (warp == kMmaWarp) {
        uint32_t g = 0, nb = 0;                                   // weight slots and B buffers consumed so far
        const uint32_t idesc = H16 ? idesc_f16_f32(128, kRsTcTile) : tc::idesc_bf16_f32(128, kRsTcTile);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int ft = tile / osplit;
            const int frame = ft / n_st, st = ft - frame * n_st;
            if (!tile_live64(frame, st, nullptr, nullptr)) continue;
            for (int kb = 0; kb < n_kb; ++kb, ++nb) {
                const uint32_t bb = nb & 1u;
                mbar_wait(&b_full[bb], (nb >> 1) & 1u);
                tc::fence_after_sync();
                for (int blk = 0; blk < nblk; ++blk)
                    for (int wp = 0; wp < NP; ++wp, ++g) {
                        const uint32_t slot = g % kRsTcRing;
                        mbar_wait(&full[slot], (g / kRsTcRing) & 1u);
                        tc::fence_after_sync();
                        if (tc::elect_one()) {
                            if (H16)
                                issue_split_h(wp, ring_s + slot * kSlotBytes, b_s + bb * 3 * kRsTcBPart, kRsTcBPart, blk * 64u,
                                              256u + blk * 64u, idesc, kb == 0 && wp == 0);
                            else
                                issue_split(wp, ring_s + slot * kSlotBytes, b_s + bb * 3 * kRsTcBPart, kRsTcBPart, blk * 64u,
                                            256u + blk * 64u, idesc, kb == 0 && wp == 0);
                            tc::mma_commit(&empty[slot]);
                        }
                        __syncwarp();
                    }
                if (tc::elect_one()) tc::mma_commit(&b_empty[bb]);    // this K block's operand may be overwritten
                __syncwarp();
            }
            if (tc::elect_one()) tc::mma_commit(acc_ready);
            __syncwarp();
        }
    } else {
        const int q = warp & 3, cg = warp >> 2;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        uint32_t acc_phase = 0, nb = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int osp = tile % osplit, ft = tile / osplit;
            const int frame = ft / n_st, st = ft - frame * n_st;
            int gi2[2], valid2[2];
            if (!tile_live64(frame, st, gi2, valid2)) continue;
            if (warp == 0) {
                s_gi[lane] = gi2[0]; s_gi[32 + lane] = gi2[1];
                s_valid[lane] = valid2[0]; s_valid[32 + lane] = valid2[1];
            }
            epi_bar();
            // two work units per thread and K block: 4 consecutive source samples of streams tid / 16 and tid / 16 + 32
            // (16 neighbouring lanes read one stream's 256-byte K block: coalesced, one 16-byte load each when aligned)
            const int c4 = tid & 15;
            int su[2];
            bool valid[2];
            long long b0[2];
            uint32_t off[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                su[j] = (tid >> 4) + 32 * j;
                valid[j] = s_valid[su[j]] != 0;
                b0[j] = valid[j] ? (long long)s_gi[su[j]] * p.stride + (long long)frame * p.n_in + 4 * c4 : 0;
                off[j] = tc::sw128_offset((uint32_t)su[j], (uint32_t)c4 * 4u, (uint32_t)kRsTcTile);
            }
            const bool vec = p.pcm == 0 ? ((reinterpret_cast<uintptr_t>(p.audio) & 15u) == 0 && (p.stride & 3) == 0)
                                        : ((reinterpret_cast<uintptr_t>(p.audio) & 7u) == 0 && (p.stride & 3) == 0);
            auto fetch = [&](int kb, int j) -> float4 {       // raw sample values (int16 not yet divided)
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid[j] && kb < n_kb) {
                    const long long o = b0[j] + 64LL * kb;
                    if (p.pcm == 0) {
                        const float *src = reinterpret_cast<const float *>(p.audio) + o;
                        if (vec) r = __ldg(reinterpret_cast<const float4 *>(src));
                        else r = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
                    } else {
                        const short *src = reinterpret_cast<const short *>(p.audio) + o;
                        if (vec) {
                            const short4 t = __ldg(reinterpret_cast<const short4 *>(src));
                            r = make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
                        } else {
                            r = make_float4((float)__ldg(src), (float)__ldg(src + 1), (float)__ldg(src + 2), (float)__ldg(src + 3));
                        }
                    }
                }
                return r;
            };
            auto to_unit = [&](float4 r) -> float4 {          // PCM scaling exactly as the loaders of the model kernels do it
                if (p.pcm == 1) r = make_float4(__fdiv_rn(r.x, 32767.0f), __fdiv_rn(r.y, 32767.0f), __fdiv_rn(r.z, 32767.0f), __fdiv_rn(r.w, 32767.0f));
                else if (p.pcm == 2) r = make_float4(r.x * (1.0f / 32768.0f), r.y * (1.0f / 32768.0f), r.z * (1.0f / 32768.0f), r.w * (1.0f / 32768.0f));
                return r;
            };
            float sc[2] = {1.f, 1.f};
            if (H16) {
                // first pass: each stream's maximum over the chunk (its 16 lanes are neighbours) -> that stream's scale
                float mx[2] = {0.f, 0.f};
                for (int kb0 = 0; kb0 < n_kb; kb0 += 4) {          // n_in / 64 = 4, 12 or 24: eight loads in flight per thread
                    float4 r[4][2];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) r[i][j] = fetch(kb0 + i, j);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float4 u = to_unit(r[i][j]);
                            mx[j] = fmaxf(mx[j], fmaxf(fmaxf(fabsf(u.x), fabsf(u.y)), fmaxf(fabsf(u.z), fabsf(u.w))));
                        }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t mb = __float_as_uint(mx[j]);
#pragma unroll
                    for (int w = 1; w < 16; w <<= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, w));
                    const Scale S = scale_from_max(mb);
                    sc[j] = S.s;
                    if (c4 == 0) s_inv[su[j]] = S.inv * inv_w;   // read by the epilogue (the K-block barriers lie in between)
                }
            }
            // K block kb + 1 is fetched while K block kb is converted and its MMAs run
            float4 nxt[2] = {fetch(0, 0), fetch(0, 1)};
            for (int kb = 0; kb < n_kb; ++kb, ++nb) {
                float4 v[2] = {to_unit(nxt[0]), to_unit(nxt[1])};
                nxt[0] = fetch(kb + 1, 0);
                nxt[1] = fetch(kb + 1, 1);
                uint32_t w[2][3][2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (H16) {
                        split2x2_h(v[j].x * sc[j], v[j].y * sc[j], w[j][0][0], w[j][1][0]);
                        split2x2_h(v[j].z * sc[j], v[j].w * sc[j], w[j][0][1], w[j][1][1]);
                    } else {
                        split3x2(v[j].x, v[j].y, w[j][0][0], w[j][1][0], w[j][2][0]);
                        split3x2(v[j].z, v[j].w, w[j][0][1], w[j][1][1], w[j][2][1]);
                    }
                }
                const uint32_t bb = nb & 1u;
                mbar_wait(&b_empty[bb], ((nb >> 1) & 1u) ^ 1u);   // the MMAs that read this buffer two K blocks ago are done
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    unsigned char *dst = bbuf + bb * 3 * kRsTcBPart + off[j];
#pragma unroll
                    for (int part = 0; part < NP; ++part)
                        *reinterpret_cast<uint2 *>(dst + part * kRsTcBPart) = make_uint2(w[j][part][0], w[j][part][1]);
                }
                tc::fence_async_smem();
                epi_bar();
                if (tid == 0) mbar_arrive(&b_full[bb]);
            }
            // ---- epilogue: D[sample][stream] -> out[stream][frame*512 + sample]
            mbar_wait(acc_ready, acc_phase); acc_phase ^= 1u;
            tc::fence_after_sync();
#pragma unroll 1
            for (int blk = 0; blk < nblk; ++blk) {
                const int sample = (osp * nblk + blk) * 128 + 32 * q + lane;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const int c0 = cg * 16 + ch * 8;
                    float m[8], cr[8];
                    tmem_ld8(lane_addr + blk * 64 + c0, m);
                    tmem_ld8(lane_addr + 256 + blk * 64 + c0, cr);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int c = c0 + e;
                        if (s_valid[c])
                            p.out[(size_t)s_gi[c] * p.max_frames * 512 + (size_t)frame * 512 + sample] =
                                H16 ? (m[e] + cr[e]) * s_inv[c] : m[e] + cr[e];
                    }
                }
            }
            tc::fence_before_sync();
            epi_bar();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kProducerWarp) tc::tmem_dealloc(0u, 512);
}

}  // namespace tc5
}  // namespace cvad
