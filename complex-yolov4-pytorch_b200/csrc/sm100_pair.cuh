// sm100_pair.cuh -- inline-PTX wrappers of the CTA-pair (tcgen05 cta_group::2) forms used by conv_pair.cu and conv_wgrad.cu:
// TMA loads that signal the LEADER CTA's mbarrier, the 2-CTA MMA, the commit that arrives on a barrier in both CTAs,
// pair-wide TMEM allocation.  Protocol after CUTLASS' 2-SM collectives (see conv_pair.cu).
#pragma once
#include "sm100.cuh"

namespace cy4 {
using namespace sm100;

constexpr uint32_t kPeerMask = 0xFEFFFFFFu;         // clears the CTA-rank bit of a shared-window address: the even CTA's copy


__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(const CUtensorMap *m, uint64_t *bar, void *dst, int c, int w, int h, int n,
                                                        uint16_t off_w, uint16_t off_h)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c), "r"(w), "r"(h), "r"(n),
          "h"(off_w), "h"(off_h)
        : "memory");
}
__device__ __forceinline__ void umma_f16_lohi_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                   uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once all MMAs issued so far by this thread have completed) on the barrier at this offset in both CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem)       // warp 1 of BOTH CTAs, same smem offset
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

}  // namespace cy4
