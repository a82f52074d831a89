// Micro-benchmark: TMEM -> register read bandwidth of tcgen05.ld.32x32b for different vector widths / warp counts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int X> struct Ld;
#define R4(b) "%" #b
template <> struct Ld<8> {
    static __device__ __forceinline__ void go(uint32_t a, uint32_t& acc) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) acc ^= v[i];
    }
};
template <> struct Ld<16> {
    static __device__ __forceinline__ void go(uint32_t a, uint32_t& acc) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(a) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) acc ^= v[i];
    }
};
template <> struct Ld<32> {
    static __device__ __forceinline__ void go(uint32_t a, uint32_t& acc) {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(a) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
};

// variant without a wait after every load: two loads in flight per warp
template <int X>
__global__ void bw_kernel(int warps, int reps, long long* out, uint32_t* sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    if (warp < warps) {
        const int half = (warp >> 2) * 256;             // warps 4..7 read the other half of the columns
        for (int r = 0; r < reps; ++r)
            for (int c = 0; c < 256; c += X) Ld<X>::go(tb + ((half + c) & 511), acc);
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tptr) : "memory");
}

template <int X> void run(int warps) {
    long long* d; uint32_t* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4096);
    const int reps = 200;
    bw_kernel<X><<<1, 256>>>(warps, reps, d, s);
    bw_kernel<X><<<1, 256>>>(warps, reps, d, s);
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * reps * 256 * 32 * 4;
    printf("x%-3d %d warps: %8lld clk, %6.1f B/clk/SM (%s)\n", X, warps, h, bytes / h, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(s);
}
int main() {
    for (int w : {1, 4, 8}) { run<8>(w); run<16>(w); run<32>(w); }
    return 0;
}
