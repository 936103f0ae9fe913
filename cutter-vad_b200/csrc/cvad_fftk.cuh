// cvad_fftk.cuh -- the two kernels built on cvad_fft.cuh's double-precision complex-256 transform:
//
//   resample_fft_kernel<R>   8 / 24 / 48 kHz -> 16 kHz, one chunk of 256 R source samples -> one 512-sample frame,
//                            = scipy.signal.resample in float64 rounded once to float32 (AudioUtils.resample_audio,
//                            /root/reference/src/real_time_vad/utils/audio.py:19-55).  Same ResampleStep contract and
//                            same output buffer as resample_kernel / resample_tc_kernel (cvad_resample.cuh, cvad_v4tc.cuh).
//   v4_stft_fft_kernel       v4's STFT with the exact Hann x DFT-256 basis (the graph's reflect-pad 96|96, conv k256 s64,
//                            SURVEY.md 8a "S4" rows 1-2), float32 re / im per (item, column) to HBM; v4tc_stft_kernel<true>
//                            adds the product with (file basis - exact basis) and forms the magnitude.
//
// Work split: 16 threads per complex-256 transform, so a warp runs two transforms side by side.
//   resampler: a warp owns TWO frames (items 2u, 2u+1 of the launch's stream list, same frame index): its 2 (R+1)/2
//              forward transforms keep both half-warps busy, the spectral stage runs on all 32 lanes, the two inverse
//              transforms again on one half-warp each.  Shared memory: 2 (R+1)/2 buffers of 4,352 bytes per warp.
//   STFT:      a CTA owns one tile of 16 items (the tile indexing of v4_frontend_kernel), a warp one item = four
//              transforms (two windows each), two per half-warp.
#pragma once
#include "cvad_fft.cuh"
#include "cvad_resample.cuh"
#include "cvad_v5.cuh"

namespace cvad {

template <int R>
struct RsFft {
    static constexpr int NP = (R + 1) / 2;                              // forward transforms per frame
    static constexpr int WPC = R == 6 ? 8 : (R == 3 ? 12 : 16);         // warps per CTA (<= 209 KB of shared memory)
    static constexpr size_t kSmem = (size_t)WPC * 2 * NP * fft::kBuf * sizeof(double2);
};

template <int R>
__global__ void __launch_bounds__(RsFft<R>::WPC * 32, 1) resample_fft_kernel(const ResampleStep p, const double2 *__restrict__ T) {
    using namespace fft;
    constexpr int NP = RsFft<R>::NP, WPC = RsFft<R>::WPC, NX = 256 * R;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, t = lane & 15;
    double2 *wb = reinterpret_cast<double2 *>(smem_raw) + (size_t)warp * (2 * NP * kBuf);
    // chained steps launch the model kernel with programmatic stream serialization: let its grid set up now
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int n_items = p.count ? *p.count : p.n_streams;
    const int n_pairs = (n_items + 1) >> 1;
    const int n_units = p.max_frames * n_pairs;
    const bool vec = (p.stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.audio) & (p.pcm == 0 ? 15u : 7u)) == 0;
    for (int unit = blockIdx.x * WPC + warp; unit < n_units; unit += gridDim.x * WPC) {
        const int frame = unit / n_pairs, pair = unit - frame * n_pairs;
        int gi[2];
        bool valid[2];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const int j = 2 * pair + f;
            gi[f] = -1;
            valid[f] = false;
            if (j < n_items) {
                gi[f] = p.list ? p.list[j] : j;
                valid[f] = frame < (p.n_frames ? p.n_frames[gi[f]] : p.max_frames);
            }
        }
        if (!valid[0] && !valid[1]) continue;     // warp-uniform

        // ---- load: sample m of the chunk is element m / R of sub-sequence m % R; sub-sequences 2p, 2p+1 share transform p.
        //      All of a frame's loads are issued (16 or 8 bytes per lane each) before the first one is used.
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            if (!valid[f]) continue;
            double2 *fb = wb + f * NP * kBuf;
            const long long base = (long long)gi[f] * p.stride + (long long)frame * NX;
            constexpr int NV = NX / 128;                      // 4-sample groups per lane
            float q[NV][4];
            if (vec) {
                if (p.pcm == 0) {
                    const float4 *src = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.audio) + base);
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        const float4 v = __ldg(src + lane + 32 * i);
                        q[i][0] = v.x; q[i][1] = v.y; q[i][2] = v.z; q[i][3] = v.w;
                    }
                } else {
                    const short4 *src = reinterpret_cast<const short4 *>(reinterpret_cast<const short *>(p.audio) + base);
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        const short4 v = __ldg(src + lane + 32 * i);
                        q[i][0] = (float)v.x; q[i][1] = (float)v.y; q[i][2] = (float)v.z; q[i][3] = (float)v.w;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < NV; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const long long off = base + 4 * (lane + 32 * i) + e;
                        q[i][e] = p.pcm == 0 ? __ldg(reinterpret_cast<const float *>(p.audio) + off)
                                             : (float)__ldg(reinterpret_cast<const short *>(p.audio) + off);
                    }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float x = q[i][e];
                    if (p.pcm == 1) x = __fdiv_rn(x, 32767.0f);
                    else if (p.pcm == 2) x = x * (1.0f / 32768.0f);
                    const int m = 4 * (lane + 32 * i) + e;
                    const int r = m % R, n = m / R;
                    reinterpret_cast<double *>(fb + (r >> 1) * kBuf + pos_in(n))[r & 1] = (double)x;
                }
            if (R & 1)
                for (int n = lane; n < 256; n += 32) fb[(NP - 1) * kBuf + pos_in(n)].y = 0.0;
        }
        __syncwarp();
        // ---- forward transforms: job j = 2 jj + half lives in buffer j (frame j / NP, transform j % NP)
#pragma unroll 1
        for (int jj = 0; jj < NP; ++jj) {
            const int j = 2 * jj + half;
            const bool on = j >= NP ? valid[1] : valid[0];
            if (on) pass1<false>(wb + j * kBuf, t, T, 6);
            __syncwarp();
            if (on) pass2<false>(wb + j * kBuf, t);
            __syncwarp();
        }
        // ---- spectral stage: X[k], X[256 - k] -> the inverse transform's inputs, written over the frame's buffer 0
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            if (!valid[f]) continue;
            double2 *fb = wb + f * NP * kBuf;
            double2 Zk[5], Zm[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int k = lane + 32 * i;
                if (k <= 128) rs_spectrum<R>(fb, k, T, Zk[i], Zm[i]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int k = lane + 32 * i;
                if (k <= 128) {
                    fb[pos_in(k)] = Zk[i];
                    if (k > 0 && k < 128) fb[pos_in(256 - k)] = Zm[i];
                }
            }
        }
        __syncwarp();
        // ---- inverse transform of frame `half`: z[n] = y[2n] + i y[2n+1]
        {
            double2 *fb = wb + half * NP * kBuf;
            const bool on = half ? valid[1] : valid[0];
            if (on) pass1<true>(fb, t, T, 6);
            __syncwarp();
            if (on) pass2<true>(fb, t);
            __syncwarp();
        }
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            if (!valid[f]) continue;
            const double2 *fb = wb + f * NP * kBuf;
            float2 *dst = reinterpret_cast<float2 *>(p.out + (size_t)gi[f] * p.max_frames * 512 + (size_t)frame * 512);
            for (int n = lane; n < 256; n += 32) {
                const double2 z = fb[pos_out(n)];
                dst[n] = make_float2((float)z.x, (float)z.y);
            }
        }
        __syncwarp();
    }
}

// ---- v4 STFT.  Output per tile (16 items x 8 columns): [col = t * 16 + item][re | im][132] floats (bins 0..128)
constexpr int kV4FftRow = 132;
constexpr int kV4FftTile = 128 * 2 * kV4FftRow;
constexpr size_t kV4FftSmem = (size_t)16 * 704 * sizeof(float) + (size_t)32 * fft::kBuf * sizeof(double2) + 256 * sizeof(double);

__global__ void __launch_bounds__(512, 1) v4_stft_fft_kernel(const V5Step p, const double2 *__restrict__ T, float *__restrict__ fft_out) {
    using namespace fft;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *bufs = reinterpret_cast<double2 *>(smem_raw);
    float *xp_all = reinterpret_cast<float *>(smem_raw + (size_t)32 * kBuf * sizeof(double2));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, t = lane & 15;
    float *x = xp_all + warp * 704;
    double2 *buf = bufs + (warp * 2 + half) * kBuf;

    double *hann = reinterpret_cast<double *>(xp_all + 16 * 704);    // the window, once per CTA
    for (int n = threadIdx.x; n < 256; n += blockDim.x) hann[n] = hann256(T, n);
    __syncthreads();

    const int n_ft = 2 * p.n_stiles;
    const int n_tiles = p.max_frames * n_ft;
    const int flen = p.frame_len < 512 ? p.frame_len : 512;
    // this warp's item of a tile: raw samples k = lane + 32 j into registers (the NEXT tile's while the current one is transformed)
    struct Item { bool valid; bool dn; int i; };
    auto fetch = [&](int tile, float (&raw)[16]) -> Item {
        Item it{false, false, 0};
        if (tile >= n_tiles) return it;
        const int frame = tile / n_ft, ft = tile - frame * n_ft;
        it.i = ft * 16 + warp;
        if (it.i < p.n_streams) {
            const int slot = p.slots ? p.slots[it.i] : it.i;
            it.valid = frame < (p.n_frames ? p.n_frames[it.i] : p.max_frames);
            if (it.valid) it.dn = p.denoise[slot] != 0;
        }
        const long long base = (long long)it.i * p.stride + (long long)frame * p.hop;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = lane + 32 * j;
            float v = 0.f;
            if (it.valid && k < flen)
                v = p.pcm == 0 ? __ldg(reinterpret_cast<const float *>(p.audio) + base + k)
                               : (float)__ldg(reinterpret_cast<const short *>(p.audio) + base + k);
            raw[j] = v;
        }
        return it;
    };
    float cur[16], nxt[16];
    Item it = fetch(blockIdx.x, cur);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const Item it_next = fetch(tile + gridDim.x, nxt);
        // ---- frame (zero-padded / truncated to 512): PCM scaling, non-finite flag, gate; an item that is not live reads as silence
        {
            bool bad = false;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float v = cur[j];
                if (p.pcm == 1) v = __fdiv_rn(v, 32767.0f);
                else if (p.pcm == 2) v = v * (1.0f / 32768.0f);
                if (!isfinite(v)) bad = true;
                if (it.dn && !(fabsf(v) > 0.01f)) v = 0.0f;
                x[96 + lane + 32 * j] = v;
            }
            if (__any_sync(0xffffffffu, bad) && lane == 0 && p.status) atomicOr(&p.status[it.i], 1u);
        }
        __syncwarp();
        // reflect padding without edge repeat: xp[j] = x[96 - j], xp[608 + j] = x[510 - j]
        for (int j = lane; j < 96; j += 32) {
            x[j] = x[192 - j];
            x[608 + j] = x[606 - j];
        }
        __syncwarp();
        float *fo = fft_out + (size_t)tile * kV4FftTile;
#pragma unroll 1
        for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * jj + half;             // windows 2j and 2j+1: samples 128 j + n and 128 j + 64 + n
            double2 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int n = 16 * q + t;
                const double w = hann[n];
                v[q] = make_double2(w * (double)x[128 * j + n], w * (double)x[128 * j + 64 + n]);
            }
            pass1_regs<false>(v, buf, t, T, 6);
            __syncwarp();
            pass2<false>(buf, t);
            __syncwarp();
            float *fa = fo + (size_t)((2 * j) * 16 + warp) * (2 * kV4FftRow);
            float *fb = fo + (size_t)((2 * j + 1) * 16 + warp) * (2 * kV4FftRow);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const int k = t + 16 * q;
                if (k <= 128) {
                    double2 A, B;
                    unpack2(buf, k, A, B);
                    fa[k] = (float)A.x; fa[kV4FftRow + k] = (float)A.y;
                    fb[k] = (float)B.x; fb[kV4FftRow + k] = (float)B.y;
                }
            }
            __syncwarp();
        }
        it = it_next;
#pragma unroll
        for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
    }
}

}  // namespace cvad
