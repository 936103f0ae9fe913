// cvad_tc.cuh -- tcgen05 (5th-gen tensor core) building blocks: TMEM allocation, shared-memory
// matrix descriptors for the canonical K-major SWIZZLE_128B layout, the BF16 MMA issue and
// the TMEM -> register load.  Bit layouts follow the PTX ISA "tcgen05" matrix / instruction
// descriptor tables (same fields as CUTLASS's cute/arch/mma_sm100_desc.hpp).
//
// Operand layout used everywhere here (both A and B are K-major, 2-byte elements):
//   a tile of R rows x 64 elements is R rows of 128 bytes; rows are grouped by 8 into 1024-byte
//   atoms; inside an atom, 16-byte chunk c of row r is stored at chunk position c ^ (r % 8).
//   Tiles for successive 64-element K blocks follow each other (R * 128 bytes apart).
//   One MMA consumes K = 16 elements = 32 bytes of every row: the descriptor's start address
//   advances by 32 bytes per K step inside a 64-element block.
#pragma once
#include <cuda_bf16.h>

#include "cvad_common.cuh"

namespace cvad {
namespace tc {

// byte offset of element (row r, k) inside an operand of R rows (R % 8 == 0), K-major SW128
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t k, uint32_t R) {
    const uint32_t kb = k >> 6, c = (k >> 3) & 7u, e = k & 7u;
    return kb * R * 128u + (r >> 3) * 1024u + (r & 7u) * 128u + ((c ^ (r & 7u)) << 4) + e * 2u;
}

// matrix descriptor: start address (>>4) | LBO (>>4) << 16 | SBO (>>4) << 32 | version 1 << 46 | SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1u << 16;                    // leading byte offset: unused for swizzled K-major (1)
    d |= (uint64_t)(1024u >> 4) << 32;          // stride byte offset: 8-row atom = 1024 bytes
    d |= (uint64_t)1u << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2u << 61;                    // SWIZZLE_128B
    return d;
}

// instruction descriptor, kind::f16: D = F32, A = B = BF16, both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_bf16_f32(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {      // true on exactly one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// this warp's 32 TMEM lanes x 32 consecutive columns -> 32 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

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


// ---- issue-rate probe (tools/tc_mma_rate.py): how many cycles does one M x N x 16 BF16 MMA cost when both
// operands come from shared memory?  One thread issues `reps` groups of 4 MMAs (one 64-element K block),
// commits, and the CTA waits; cycles are taken with clock64 by the issuing thread.  out[0] = cycles,
// out[1] = number of MMAs.  a_tiles distinct A tiles are cycled so the A operand address changes like a
// weight stream's does.
__global__ void __launch_bounds__(128, 1) tc_rate_kernel(int M, int N, int reps, int a_tiles, int n_acc, long long *out) {
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
        const uint32_t idesc = idesc_bf16_f32(M, N);
        const uint64_t a_base = smem_desc_sw128(smem_u32(sA)), b_base = smem_desc_sw128(smem_u32(sB));
        int at = 0, acc = 0;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint64_t ad = a_base + (uint64_t)(at * (16384 >> 4));
            const uint32_t d = acc * N;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_bf16(d, ad + ks * 2, b_base + ks * 2, idesc, 1u);
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
