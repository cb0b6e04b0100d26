// curand_host.cu -- ORACLE / TEST INFRASTRUCTURE ONLY.
// cuRAND's own Philox4x32-10 (third-party: cuRAND 10.3.10, CUDA 12.9 headers under
// /usr/local/cuda/include) compiled for the HOST through its QUALIFIERS hook
// (curand_philox4x32_x.h:84-86), so the oracle's restatement and the engine's kernels can
// be pinned word-for-word against the library the reference's call sites would use
// (curand_init(seed, idx, 0, ...) at inc/tool.cuh:194 with the Philox state type).
#define QUALIFIERS static inline __host__ __device__
#include <cstdint>
#include <curand_kernel.h>

extern "C" {

// 4 words of block `block` of subsequence `subsequence`: curand_init(seed, subsequence, 4*block) + curand4.
void curand_host_block(uint64_t seed, uint64_t subsequence, uint64_t block, uint32_t out[4])
{
    curandStatePhilox4_32_10_t s;
    curand_init(seed, subsequence, 4ull * block, &s);
    uint4 w = curand4(&s);
    out[0] = w.x; out[1] = w.y; out[2] = w.z; out[3] = w.w;
}

// Raw Philox4x32-10 bijection (Random123 known-answer tests).
void curand_host_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint4 c = make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]);
    uint2 k = make_uint2(key[0], key[1]);
    uint4 w = curand_Philox4x32_10(c, k);
    out[0] = w.x; out[1] = w.y; out[2] = w.z; out[3] = w.w;
}

// First `count` values of curand() and curand_normal() of a stream (host libm maths).
void curand_host_words(uint64_t seed, uint64_t subsequence, uint64_t offset, int count, uint32_t *out)
{
    curandStatePhilox4_32_10_t s;
    curand_init(seed, subsequence, offset, &s);
    for (int i = 0; i < count; ++i) out[i] = curand(&s);
}

void curand_host_normals(uint64_t seed, uint64_t subsequence, int count, float *out)
{
    curandStatePhilox4_32_10_t s;
    curand_init(seed, subsequence, 0, &s);
    for (int i = 0; i < count; ++i) out[i] = curand_normal(&s);
}

}  // extern "C"
