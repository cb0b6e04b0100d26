// pipe_microbench.cu -- measures the per-SM issue rates the European kernel's roofline rests on
// (SURVEY.md 8(d) assumes 128 thread-instr/clk/SM and 16 MUFU/clk/SM).  Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench tools/pipe_microbench.cu
// One CTA of 1024 threads per SM, ILP independent chains per thread, clock64() around the loop.
// Prints thread-instructions per clock per SM for each instruction mix.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 2048;
constexpr int kIlp = 8;

enum Mix { IMADW, LOP3, FFMA, EX2, LG2, SIN, SQRT, IMADW_LOP3, PHILOX_ROUND, FFMA_LOP3, FFMA_IMADW, I2F, FMNMX,
           IMAD_HI, IMAD_LO, FFMA2, FFMA2_IMADW, EX2_IMADW, FADD2, FFMA2_LOP3, EX2_FFMA2,
           FADD_IMADW, FMUL_IMADW, FFMA2X_IMADW, FFMA3X_IMADW, HFMA2_IMADW, FADD, FMUL, MIX_COUNT };
const char *kNames[] = {"imad.wide.u32", "lop3", "ffma", "mufu.ex2", "mufu.lg2", "mufu.sin(+fmul.rz)", "mufu.sqrt",
                        "imad.wide+lop3 (1:1)", "philox round (2 imad.wide + 2 lop3)", "ffma+lop3 (1:1)",
                        "ffma+imad.wide (1:1)", "i2fp.u32", "fmnmx", "imad.hi.u32", "imad (lo)", "ffma2 (packed f32x2)",
                        "ffma2+imad.wide (1:1)", "mufu.ex2+imad.wide (1:1)", "fadd2", "ffma2+lop3 (1:1)",
                        "mufu.ex2+ffma2 (1:1)", "fadd+imad.wide (1:1)", "fmul+imad.wide (1:1)", "2 ffma + imad.wide",
                        "3 ffma + imad.wide", "hfma2+imad.wide (1:1)", "fadd", "fmul"};
const int kInstrPerStep[] = {1, 1, 1, 1, 1, 2, 1, 2, 4, 2, 2, 1, 1, 1, 1, 1, 2, 2, 1, 2, 2, 2, 2, 3, 4, 2, 1, 1};

template <int MIX>
__global__ void __launch_bounds__(1024) bench(uint32_t seed, uint32_t *out, long long *cycles)
{
    uint32_t a[kIlp], b[kIlp];
    float f[kIlp], h[kIlp];
    unsigned long long g[kIlp];   // packed f32x2 chains
    unsigned long long gm, ga;
    asm("mov.b64 %0, {%1, %2};" : "=l"(gm) : "f"(0.999f), "f"(0.998f));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ga) : "f"(1e-3f), "f"(2e-3f));
#pragma unroll
    for (int j = 0; j < kIlp; ++j) {
        a[j] = seed + threadIdx.x * 7 + j;
        b[j] = seed * 3 + threadIdx.x + j * 5;
        f[j] = 1.0f + 1e-3f * (float)(threadIdx.x + j);
        h[j] = f[j] * 1.25f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(g[j]) : "f"(f[j]), "f"(f[j] + 0.5f));
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int j = 0; j < kIlp; ++j) {
            if (MIX == IMADW) {
                uint64_t p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[j]), "r"(0xD2511F53u));
                a[j] = (uint32_t)(p >> 32) + (uint32_t)p * 0;  // keep hi; lo folded away
                b[j] ^= 0;                                     // (no-op)
                asm volatile("" : "+r"(a[j]));
            } else if (MIX == LOP3) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b[j]), "r"(seed));
            } else if (MIX == FFMA) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(0.999f), "f"(1e-3f));
            } else if (MIX == EX2) {
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
            } else if (MIX == LG2) {
                asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
            } else if (MIX == SIN) {
                asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
            } else if (MIX == SQRT) {
                asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
            } else if (MIX == IMADW_LOP3) {
                uint64_t p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[j]), "r"(0xD2511F53u));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[j]) : "r"((uint32_t)(p >> 32)), "r"((uint32_t)p), "r"(seed));
            } else if (MIX == PHILOX_ROUND) {
                uint64_t p0, p1;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(a[j]), "r"(0xD2511F53u));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(b[j]), "r"(0xCD9E8D57u));
                uint32_t n0, n2;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n0) : "r"((uint32_t)(p1 >> 32)), "r"((uint32_t)p0), "r"(seed));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n2) : "r"((uint32_t)(p0 >> 32)), "r"((uint32_t)p1), "r"(seed));
                a[j] = n0;
                b[j] = n2;
            } else if (MIX == FFMA_LOP3) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(0.999f), "f"(1e-3f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b[j]), "r"(seed));
            } else if (MIX == FFMA_IMADW) {
                uint64_t p;
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(0.999f), "f"(1e-3f));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[j]), "r"(0xD2511F53u));
                a[j] = (uint32_t)(p >> 32);
                asm volatile("" : "+r"(a[j]));
            } else if (MIX == I2F) {
                float t;
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(t) : "r"(a[j]));
                a[j] = __float_as_uint(t);
            } else if (MIX == FMNMX) {
                asm volatile("max.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(__uint_as_float(b[j])));
            } else if (MIX == IMAD_HI) {
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(0xD2511F53u));
            } else if (MIX == IMAD_LO) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(0xD2511F53u), "r"(b[j]));
            } else if (MIX == FFMA2 || MIX == FFMA2_IMADW || MIX == FFMA2_LOP3 || MIX == EX2_FFMA2) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(g[j]) : "l"(gm), "l"(ga));
                if (MIX == FFMA2_IMADW) {
                    uint64_t p;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[j]), "r"(0xD2511F53u));
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[j]) : "r"((uint32_t)(p >> 32)), "r"((uint32_t)p), "r"(0u));
                } else if (MIX == FFMA2_LOP3) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b[j]), "r"(seed));
                } else if (MIX == EX2_FFMA2) {
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
                }
            } else if (MIX == EX2_IMADW) {
                uint64_t p;
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[j]), "r"(0xD2511F53u));
                a[j] = (uint32_t)(p >> 32) ^ (uint32_t)p;
            } else if (MIX == FADD_IMADW || MIX == FMUL_IMADW || MIX == FFMA2X_IMADW || MIX == FFMA3X_IMADW ||
                       MIX == HFMA2_IMADW) {
                uint64_t p;
                if (MIX == FADD_IMADW) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(1e-3f));
                if (MIX == FMUL_IMADW) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(0.9999f));
                if (MIX == FFMA2X_IMADW || MIX == FFMA3X_IMADW) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(0.999f), "f"(1e-3f));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(h[j]) : "f"(0.999f), "f"(1e-3f));
                    if (MIX == FFMA3X_IMADW) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(0.998f), "f"(2e-3f));
                }
                if (MIX == HFMA2_IMADW) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(b[j]) : "r"(0x3c003c00u), "r"(0x10001000u));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[j]), "r"(0xD2511F53u));
                a[j] = (uint32_t)(p >> 32) ^ (uint32_t)p;
            } else if (MIX == FADD) {
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(1e-3f));
            } else if (MIX == FMUL) {
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(0.9999f));
            } else if (MIX == FADD2) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(g[j]) : "l"(ga));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < kIlp; ++j) acc ^= a[j] ^ b[j] ^ __float_as_uint(f[j]) ^ __float_as_uint(h[j]) ^ (uint32_t)g[j] ^ (uint32_t)(g[j] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MIX>
void run(int sms, uint32_t *out, long long *cycles)
{
    bench<MIX><<<sms, 1024>>>(1234u, out, cycles);
    bench<MIX><<<sms, 1024>>>(1234u, out, cycles);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    const double instr = 1024.0 * kIters * kIlp * kInstrPerStep[MIX];
    printf("%-40s %8.2f thread-instr/clk/SM   (%.0f cycles)\n", kNames[MIX], instr / avg, avg);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs, clockRate %d kHz\n", prop.name, sms, prop.clockRate);
    uint32_t *out;
    long long *cycles;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
    cudaMalloc(&cycles, sizeof(long long) * sms);
    run<IMADW>(sms, out, cycles);
    run<LOP3>(sms, out, cycles);
    run<FFMA>(sms, out, cycles);
    run<EX2>(sms, out, cycles);
    run<LG2>(sms, out, cycles);
    run<SIN>(sms, out, cycles);
    run<SQRT>(sms, out, cycles);
    run<IMADW_LOP3>(sms, out, cycles);
    run<PHILOX_ROUND>(sms, out, cycles);
    run<FFMA_LOP3>(sms, out, cycles);
    run<FFMA_IMADW>(sms, out, cycles);
    run<I2F>(sms, out, cycles);
    run<FMNMX>(sms, out, cycles);
    run<IMAD_HI>(sms, out, cycles);
    run<IMAD_LO>(sms, out, cycles);
    run<FFMA2>(sms, out, cycles);
    run<FFMA2_IMADW>(sms, out, cycles);
    run<EX2_IMADW>(sms, out, cycles);
    run<FADD2>(sms, out, cycles);
    run<FFMA2_LOP3>(sms, out, cycles);
    run<EX2_FFMA2>(sms, out, cycles);
    run<FADD_IMADW>(sms, out, cycles);
    run<FMUL_IMADW>(sms, out, cycles);
    run<FFMA2X_IMADW>(sms, out, cycles);
    run<FFMA3X_IMADW>(sms, out, cycles);
    run<HFMA2_IMADW>(sms, out, cycles);
    run<FADD>(sms, out, cycles);
    run<FMUL>(sms, out, cycles);
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
