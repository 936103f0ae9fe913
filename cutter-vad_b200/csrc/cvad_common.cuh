// cvad_common.cuh -- shared device helpers: mbarrier + bulk-copy (TMA 1-D) weight ring.
//
// The weights of every layer are pre-packed on the host into the exact order the
// kernels consume them ("weight stream"): a sequence of <=16 KB chunks.  One thread
// per CTA pushes chunks into a shared-memory ring with cp.async.bulk (SASS: UBLKCP)
// completing on an mbarrier; all threads wait on the barrier, consume the chunk with
// LDS.128, and a __syncthreads() hands the slot back to the producer.  The stream is
// periodic (one period per tile / per frame), so prefetch runs across layer and tile
// boundaries and the ring is drained before the CTA exits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvad {

constexpr int kTile = 32;            // streams (items) per CTA tile
constexpr int kThreads = 512;        // threads per CTA (16 warps: 4 per scheduler)
constexpr int kIG = 4;               // items per thread (one LDS.128 of the k-major activations)
constexpr int kTM = kTile / kIG;     // 8 item groups per tile
constexpr int kRingStages = 4;       // weight ring depth
constexpr int kRingSlotFloats = 4096;// 16 KB per ring slot

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

// 1-D bulk async copy global -> shared, completion counted in bytes on `bar`.
// bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// Register tiles use the sm_100 packed FP32 FMA (PTX fma.rn.f32x2, SASS FFMA2): one
// instruction = two IEEE fp32 FMAs on a 64-bit register pair.  A 3-register scalar FFMA
// issues at most every other cycle per scheduler on Blackwell; FFMA2 restores the full
// 128 FMA/clk/SM.  Accumulators are paired along the ITEM axis (pairs come straight out
// of the LDS.128 of the activations); the weight is duplicated into both halves.
//   acc[ip][j] = (item 2ip, item 2ip+1) x output j
struct Dup4 {
    float2 d[4];
};
__device__ __forceinline__ Dup4 dup4(const float4 &w) {
    Dup4 r;
    r.d[0] = make_float2(w.x, w.x);
    r.d[1] = make_float2(w.y, w.y);
    r.d[2] = make_float2(w.z, w.z);
    r.d[3] = make_float2(w.w, w.w);
    return r;
}

// acc (4 items x 4 outputs) += a (4 items) (x) w (4 outputs): 8 FFMA2
__device__ __forceinline__ void fma4x4(float2 (&acc)[2][4], const float4 &a, const Dup4 &w) {
    const float2 a0 = make_float2(a.x, a.y), a1 = make_float2(a.z, a.w);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        acc[0][j] = __ffma2_rn(a0, w.d[j], acc[0][j]);
        acc[1][j] = __ffma2_rn(a1, w.d[j], acc[1][j]);
    }
}

__device__ __forceinline__ void zero_tile(float2 (&acc)[2][4]) {
#pragma unroll
    for (int ip = 0; ip < 2; ++ip)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[ip][j] = make_float2(0.f, 0.f);
}

// view the paired accumulators as v[item][output] (pure register renaming)
__device__ __forceinline__ void unpack_tile(const float2 (&acc)[2][4], float (&v)[4][4]) {
#pragma unroll
    for (int ip = 0; ip < 2; ++ip)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[2 * ip][j] = acc[ip][j].x;
            v[2 * ip + 1][j] = acc[ip][j].y;
        }
}

// column j of a 4x4 tile as a float4 over the 4 items
__device__ __forceinline__ float4 col4(const float (&v)[4][4], int j) {
    return make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
}
__device__ __forceinline__ float4 add4(const float4 &a, const float4 &b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 bias_relu4(const float4 &a, const float4 &b, float bias) {
    return make_float4(fmaxf(a.x + b.x + bias, 0.f), fmaxf(a.y + b.y + bias, 0.f), fmaxf(a.z + b.z + bias, 0.f),
                       fmaxf(a.w + b.w + bias, 0.f));
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// Periodic weight stream -> shared-memory ring.
struct WeightRing {
    float *buf;          // kRingStages * kRingSlotFloats
    uint64_t *bars;      // kRingStages "full" barriers
    const float *gsrc;   // packed weight stream in HBM/L2
    uint32_t g;          // chunks consumed so far by this CTA (uniform across threads)
    uint32_t first = 0;  // v4 front end fed with a ready-made spectrogram: the stream starts at chunk `first` ...
    uint32_t period = 0; // ... and repeats every `period` chunks (0 = the kernel's full period)
};

__device__ __forceinline__ const float *ring_wait(const WeightRing &r) {
    const uint32_t slot = r.g % kRingStages;
    mbar_wait(&r.bars[slot], (r.g / kRingStages) & 1u);
    return r.buf + slot * kRingSlotFloats;
}

}  // namespace cvad
