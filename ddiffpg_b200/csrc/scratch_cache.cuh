// L2 handling of the per-CTA scratch blocks of the fused chains (ELU' of the critic chain, the state partial of the
// sampler's layer 0).  Such a block is written and read back by the same thread within one tile and rewritten by the next
// tile; it never has to reach DRAM, but with default priorities the L2 writes back 40 % of it under the streaming traffic
// around it (critic chain: 185 MB written per launch against 2 MB of results).  Stores and loads carry an evict_last policy
// instead (26 MB written, same duration: profiles/r02/ab_scratch_policy.txt).  The same policy on the sampler's scratch
// changes nothing measurable (same file), so it is a per-kernel switch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ddp {
namespace tc {

__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <bool KEEP>
__device__ __forceinline__ void scratch_st(uint4* p, const uint4 w) {
    if (KEEP)
        asm volatile("st.global.cg.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
                     :: "l"(p), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w), "l"(l2_evict_last_policy()) : "memory");
    else __stcg(p, w);
}
template <bool KEEP>
__device__ __forceinline__ uint4 scratch_ld(const uint4* p) {
    if (KEEP) {
        uint4 v;
        asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(l2_evict_last_policy()));
        return v;
    }
    return __ldcg(p);
}
// Drop one 128-byte line from L2 without writing it back (its bytes are dead until rewritten).  One lane per line; the
// address is made to depend on a loaded word whose sign bit is known to be clear, so the discard cannot be issued
// before the whole warp's load of that line has returned.
__device__ __forceinline__ void scratch_discard(const void* line, uint32_t loaded) {
    const char* q = reinterpret_cast<const char*>(line) + ((size_t)(loaded >> 31) << 7);
    asm volatile("discard.global.L2 [%0], 128;" :: "l"(q) : "memory");
}

}  // namespace tc
}  // namespace ddp
