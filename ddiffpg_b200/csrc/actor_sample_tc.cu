// H1, bf16 tensor-core path (tcgen05 / TMEM / TMA) -- placeholder until the kernel lands.
#include "actor_layout.cuh"

namespace ddp {

int pack_actor_tc(const ActorLayout&, const float* const[12], void*, cudaStream_t) {
    DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 actor path is not built in this library");
}
size_t actor_sample_tc_workspace(const ActorLayout&, long) { return 0; }
int actor_sample_tc(const ActorLayout&, const void*, const float*, const float*, float*, long, void*, size_t,
                    cudaStream_t) {
    DDP_FAIL(DDP_ERR_UNSUPPORTED, "DDP_BF16 actor path is not built in this library");
}

}  // namespace ddp
