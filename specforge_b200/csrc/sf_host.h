// sf_host.h — host-side shared declarations for libspecforge_b200 (errors, launch counter, op descs).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sf {

// Records a message retrievable through sf_last_error() and returns `code` (negative errno style).
int set_error(int code, const char* fmt, ...);
// Every kernel launch made by this library goes through here; bench.py reports the count.
void count_launch(int n = 1);

struct GemmDesc {
    const void* A; int64_t lda; int a_major;   // a_major/b_major: 0 = K-major, 1 = MN-major
    const void* B; int64_t ldb; int b_major;
    void* D; int64_t ldd;
    const void* R; int64_t ldr;                // residual (EPI_BF16_RESID) or null
    int M, N, K;
    int epi;                                   // sf::EPI_*
    int cta_group;                             // 0 = auto, 1, 2
};
int gemm(const GemmDesc& g, cudaStream_t stream);

#define SF_CUDA_CHECK_LAUNCH(what)                                                            \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) return ::sf::set_error(-5, "%s: %s", what, cudaGetErrorString(e__)); \
        ::sf::count_launch();                                                                 \
    } while (0)

}  // namespace sf
