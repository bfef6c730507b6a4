// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, SS) for a few (M, N) shapes and operand majors.
#include "../specforge_b200/csrc/sf_ptx.cuh"
#include <cstdio>
using namespace sf;

template <int N, int BMAJ>
__global__ void __launch_bounds__(128, 1) mma_loop(int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc<1>(smem_u32(&tslot), 512);
    // zero the operand area so MMAs read finite data
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (sbase - smem_u32(smem_raw)))[i] = 0;
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, 0, BMAJ);
        const uint64_t a = make_smem_desc_sw128(sbase, 0, 1024);
        const uint64_t b = make_smem_desc_sw128(sbase + 16384, BMAJ ? 8192 : 0, 1024);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16<1>(tmem, a + ((k * 32) >> 4), b + ((k * (BMAJ ? 2048 : 32)) >> 4), idesc, 1);
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = (t1 - t0);
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

template <int N, int BMAJ>
void run(const char* name) {
    long long* d; cudaMalloc(&d, 8);
    const int smem = 64 * 1024;
    cudaFuncSetAttribute(mma_loop<N, BMAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    mma_loop<N, BMAJ><<<148, 128, smem>>>(iters, d);
    cudaDeviceSynchronize();
    mma_loop<N, BMAJ><<<148, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s: %s  %.1f cycles per MMA (M=128, N=%d, K=16)\n", name, cudaGetErrorString(e), (double)h / (iters * 4.0), N);
    cudaFree(d);
}
int main() {
    run<64, 0>("N=64  B K-major ");
    run<128, 0>("N=128 B K-major ");
    run<256, 0>("N=256 B K-major ");
    run<64, 1>("N=64  B MN-major");
    run<128, 1>("N=128 B MN-major");
    run<32, 0>("N=32  B K-major ");
    return 0;
}
