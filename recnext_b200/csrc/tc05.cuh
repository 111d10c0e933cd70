// tc05.cuh — thin inline-PTX layer over the Blackwell 5th-generation tensor cores (tcgen05.mma with TMEM accumulators),
// mbarriers and TMA bulk copies, as used by the channel-mixer GEMMs (ffn_tc.cu) and the tensor-core RecConv kernels
// (tconv*.cuh).  sm_100a only.  Descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor"
// tables (the same fields CUTLASS names in cute/arch/mma_sm100_desc.hpp).
//
// Shared-memory operand layouts used in this repo (all SWIZZLE_NONE, i.e. built from 8 x 16-byte "core matrices"
// whose 8 rows are contiguous 16-byte chunks):
//   K-major  operand (rows = M or N, 16-byte chunk = 8 consecutive K elements):
//       element (row, k) at  (row / 8) * SBO + (k / 8) * LBO + (row % 8) * 16 + (k % 8) * 2
//       With SBO = 128 the rows of one 8-wide K chunk are simply consecutive 16-byte chunks, so "start at row r" is
//       "start address + 16 r": this is what lets a depthwise stencil read its filter rows as row-shifted operands.
//   MN-major operand (16-byte chunk = 8 consecutive M/N elements of one K index):
//       element (k, n)   at  (n / 8) * SBO + (k / 8) * LBO + (k % 8) * 16 + (n % 8) * 2
// One tcgen05.mma of kind::f16 consumes K = 16, i.e. two core matrices along K (LBO apart).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace recnext {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- instruction / shared-memory descriptors -------------------------------------------------------------------
// kind::f16 instruction descriptor: fp32 accumulate, A/B both bf16 (fmt 1) or both fp16 (fmt 0)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int ab_fmt, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// SWIZZLE_NONE shared-memory matrix descriptor (version 1 = Blackwell); byte offsets must be multiples of 16
__host__ __device__ constexpr uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
// the same descriptor with the start address moved by `bytes` (multiple of 16; no carry out of the 14-bit field expected)
__device__ __forceinline__ uint64_t sdesc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// ---- TMEM --------------------------------------------------------------------------------------------------------
// executed by ONE full warp; the base address (lane 0, first column) is written to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t slot_saddr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_saddr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all tcgen05.mma issued so far by this thread arrive (once) on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint32_t bar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}

// 32 lanes x N consecutive 32-bit columns: thread t of warp w reads TMEM lane 32 (w % 4) + t
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait suspends the warp in hardware until the phase completes or the time hint (ns) expires: a waiting warp issues almost
// no instructions (a bare spin loop costs the working warps of the same scheduler a quarter of their issue slots, ncu)
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// non-blocking probe (mbarrier.test_wait): has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// A wait that outlives 2^22 probes (each probe sleeps up to the hint) can only be a protocol bug: trap (the launch fails with a
// CUDA error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins == (1u << 22)) __trap();
    }
}
// TMA bulk copy global -> shared (bytes % 16 == 0, both 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

}  // namespace tc
}  // namespace recnext
