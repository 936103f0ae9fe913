// cvad_tc_dev.cuh -- development probes of the tcgen05 building blocks (hardware probe, MMA issue-rate and bulk-copy
// stream-rate measurements).  NOT part of the product: compiled only into libcvad_b200_dev.so (csrc/cvad_dev.cu), which
// tests/test_gpu_tc_probe.py and tools/tc_*.py load; libcvad_b200.so holds none of this.
#pragma once
#include "cvad_tc.cuh"

namespace cvad {
namespace tc {

// ---- hardware probe (tests/test_gpu_tc_probe.py): D[128][32] = A[128][256] * B[32][256]^T in BF16 -> FP32
constexpr int kProbeM = 128, kProbeN = 32, kProbeK = 256;
constexpr size_t kProbeSmem = (size_t)(kProbeM + kProbeN) * kProbeK * 2 + 1024 + 64;

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    unsigned char *sA = base;                                   // 128 x 256 bf16 = 64 KB
    unsigned char *sB = sA + kProbeM * kProbeK * 2;             // 32 x 256 bf16 = 16 KB
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + kProbeN * kProbeK * 2);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int idx = tid; idx < kProbeM * kProbeK / 8; idx += 128) {   // 16-byte chunks
        const int r = idx / (kProbeK / 8), k = (idx % (kProbeK / 8)) * 8;
        *reinterpret_cast<uint4 *>(sA + sw128_offset(r, k, kProbeM)) =
            *reinterpret_cast<const uint4 *>(A + (size_t)r * kProbeK + k);
    }
    for (int idx = tid; idx < kProbeN * kProbeK / 8; idx += 128) {
        const int r = idx / (kProbeK / 8), k = (idx % (kProbeK / 8)) * 8;
        *reinterpret_cast<uint4 *>(sB + sw128_offset(r, k, kProbeN)) =
            *reinterpret_cast<const uint4 *>(B + (size_t)r * kProbeK + k);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 32);
    fence_async_smem();          // generic-proxy stores above -> visible to the tensor core's async proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = idesc_bf16_f32(kProbeM, kProbeN);
        for (int kb = 0; kb < kProbeK / 64; ++kb)
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = smem_desc_sw128(smem_u32(sA) + kb * kProbeM * 128 + ks * 32);
                const uint64_t bd = smem_desc_sw128(smem_u32(sB) + kb * kProbeN * 128 + ks * 32);
                mma_bf16(tmem, ad, bd, idesc, (kb | ks) ? 1u : 0u);
            }
        mma_commit(bar);
    }
    {   // bounded wait: a wrong descriptor must not hang the probe (and the box)
        const uint32_t addr = smem_u32(bar);
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(0u) : "memory");
        if (!done) {
            if (tid == 0) D[0] = -12345.0f;
            return;
        }
    }
    fence_after_sync();
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(size_t)(warp * 32 + lane) * kProbeN + i] = v[i];
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}


// ---- MN-major B operand probe: D[128][0..95] = A[128][64] * B[96][64]^T with B stored MN-major SWIZZLE_64B (N contiguous),
// and D[128][96..159] = A * B[32..95]^T through a descriptor that starts one N atom further (what a conv tap does)
constexpr size_t kProbeMnSmem = 16384 + 12288 + 1024 + 64;
__global__ void __launch_bounds__(128, 1) tc_probe_mn_kernel(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = base;                                   // 128 x 64 bf16, K-major SW128
    unsigned char *sB = sA + 16384;                             // 96 x 64 bf16, MN-major SW64: lbo 512, sbo 1536
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + 12288);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int idx = tid; idx < 128 * 8; idx += 128) {
        const int r = idx / 8, k = (idx % 8) * 8;
        *reinterpret_cast<uint4 *>(sA + sw128_offset(r, k, 128)) = *reinterpret_cast<const uint4 *>(A + (size_t)r * 64 + k);
    }
    for (int idx = tid; idx < 96 * 64; idx += 128) {
        const int n = idx / 64, k = idx % 64;
        *reinterpret_cast<__nv_bfloat16 *>(sB + mn64_offset(n, k, 512u, 1536u)) = B[(size_t)n * 64 + k];
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t i96 = idesc_bf16_f32(128, 96) | kIdescBMajorMN, i64 = idesc_bf16_f32(128, 64) | kIdescBMajorMN;
        for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = smem_desc_sw128(smem_u32(sA) + ks * 32);
            mma_bf16(tmem, ad, smem_desc_mn64(smem_u32(sB) + ks * 2 * 1536, 512u, 1536u), i96, ks ? 1u : 0u);
            mma_bf16(tmem + 96, ad, smem_desc_mn64(smem_u32(sB) + 512 + ks * 2 * 1536, 512u, 1536u), i64, ks ? 1u : 0u);
        }
        mma_commit(bar);
    }
    {
        const uint32_t addr = smem_u32(bar);
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(0u) : "memory");
        if (!done) {
            if (tid == 0) D[0] = -12345.0f;
            return;
        }
    }
    fence_after_sync();
    for (int c = 0; c < 160; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) D[(size_t)(warp * 32 + lane) * 160 + c + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- issue-rate probe (tools/tc_mma_rate.py): how many cycles does one M x N x 16 BF16 MMA cost when both
// operands come from shared memory?  One thread issues `reps` groups of 4 MMAs (one 64-element K block),
// commits, and the CTA waits; cycles are taken with clock64 by the issuing thread.  out[0] = cycles,
// out[1] = number of MMAs.  a_tiles distinct A tiles are cycled so the A operand address changes like a
// weight stream's does.
__global__ void __launch_bounds__(128, 1) tc_rate_kernel(int M, int N, int reps, int a_tiles, int n_acc, int b_mn, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    unsigned char *sA = base;                                   // a_tiles x (128 x 64 bf16 = 16 KB)
    unsigned char *sB = sA + (size_t)a_tiles * 16384;           // 256 x 64 bf16 = 32 KB
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + 32768);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (a_tiles * 16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t *>(sA)[i] = 0x3C003C00u;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    long long t0 = 0, t1 = 0;
    if (tmem != 0) __trap();      // sole CTA on the SM, whole TMEM allocated: base is column 0 / lane 0
    if (warp == 0) {              // warp-uniform issue loop, one elected lane issues (keeps operands in uniform registers)
        // b_mn: B operand MN-major SWIZZLE_64B (N % 32 == 0), N atoms 512 B apart, K groups N / 32 * 512 B apart
        const uint32_t sbo = (uint32_t)(N / 32) * 512u;
        const uint32_t idesc = idesc_bf16_f32(M, N) | (b_mn ? kIdescBMajorMN : 0u);
        const uint64_t a_base = smem_desc_sw128(smem_u32(sA)),
                       b_base = b_mn ? smem_desc_mn64(smem_u32(sB), 512u, sbo) : smem_desc_sw128(smem_u32(sB));
        const uint64_t b_step = b_mn ? (uint64_t)((2u * sbo) >> 4) : 2u;
        int at = 0, acc = 0;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint64_t ad = a_base + (uint64_t)(at * (16384 >> 4));
            const uint32_t d = acc * N;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_bf16(d, ad + ks * 2, b_base + ks * b_step, idesc, 1u);
            }
            if (++at == a_tiles) at = 0;
            if (++acc == n_acc) acc = 0;
        }
        if (elect_one()) mma_commit(bar);
        __syncwarp();
    }
    mbar_wait(bar, 0);
    if (tid == 0) {
        t1 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = (long long)reps * 4;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}


// ---- bulk-copy stream probe (tools/tc_stream_rate.py): how fast can ONE SM pull 16 KB tiles from L2 into a
// shared-memory ring with cp.async.bulk when nothing consumes them?  out[0] = cycles, out[1] = tiles.
__global__ void __launch_bounds__(64, 1) bulk_rate_kernel(const unsigned char *src, size_t src_bytes, int tiles, int depth,
                                                          int tile_bytes, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + (size_t)depth * tile_bytes);
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t src_mask = (uint32_t)(src_bytes / tile_bytes) - 1u;   // power of two
        const long long t0 = clock64();
        int slot = 0;
        uint32_t round = 0;
        for (int g = 0; g < tiles + depth; ++g) {
            if (g >= depth) mbar_wait(&bars[slot], (round - 1u) & 1u);       // previous fill of this slot landed
            if (g < tiles) {
                mbar_arrive_expect_tx(&bars[slot], tile_bytes);
                bulk_g2s(base + (uint32_t)slot * (uint32_t)tile_bytes, src + (size_t)(((uint32_t)g & src_mask) * (uint32_t)tile_bytes),
                         tile_bytes, &bars[slot]);
            }
            if (++slot == depth) { slot = 0; ++round; }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = tiles; }
    }
}

}  // namespace tc
}  // namespace cvad
