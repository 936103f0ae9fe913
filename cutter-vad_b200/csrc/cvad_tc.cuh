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

// MN-major B operand, SWIZZLE_64B (2-byte elements): the contiguous dimension is N.  An atom is 8 K rows of 64 bytes
// (32 N elements); 16-byte chunk c of K row r sits at chunk position c ^ ((r >> 1) & 3) (atoms 512-byte aligned).  Atoms
// follow each other along N every `lbo` bytes and along K (groups of 8) every `sbo` bytes; one MMA (K = 16) reads two K
// groups, so the start address advances by 2 * sbo per K step.  Same canonical form as CUTLASS's
// make_umma_desc<Major::MN> with LayoutType::B64: ((8,4,n),(8,k)) : ((1,8,LBO),(32,SBO)) elements.
__host__ __device__ __forceinline__ uint32_t mn64_offset(uint32_t n, uint32_t k, uint32_t lbo, uint32_t sbo) {
    const uint32_t r = k & 7u, c = (n >> 3) & 3u;
    return (k >> 3) * sbo + (n >> 5) * lbo + r * 64u + ((c ^ ((r >> 1) & 3u)) << 4) + (n & 7u) * 2u;
}
__device__ __forceinline__ uint64_t smem_desc_mn64(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1u << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)4u << 61;                    // SWIZZLE_64B
    return d;
}
constexpr uint32_t kIdescBMajorMN = 1u << 16;   // instruction descriptor: B operand is MN-major

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

}  // namespace tc
}  // namespace cvad
