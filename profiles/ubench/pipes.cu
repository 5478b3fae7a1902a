// Micro-benchmark: issue cost of the instructions k_fold / k_hist / k_scatter lean on (sm_100a).
// Every kernel runs ITER dependent-free repetitions per thread with 8 warps x 4 CTAs per SM and reports
// SM cycles per warp-instruction per SM sub-partition.   nvcc -arch=sm_100a -O3 pipes.cu -o pipes && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int MODE>
__global__ void k(uint32_t *out, uint32_t seed)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t acc = 0;
    uint32_t v = seed + (MODE == 0 ? lane : MODE == 1 ? (lane >> 2) : MODE == 2 ? 0u : lane);
    double d0 = 1.0 + lane, d1 = 2.0 + lane, d2 = 3.0, d3 = 4.0;
    float f0 = 1.f + lane, f1 = 2.f, f2 = 3.f, f3 = 4.f;
    for (int i = 0; i < ITER; ++i)
    {
        if (MODE <= 2)
        {  // match.any: 32 distinct / 8 groups of 4 / all equal
            acc += __match_any_sync(0xffffffffu, v + i);
            acc += __match_any_sync(0xffffffffu, v ^ i);
            acc += __match_any_sync(0xffffffffu, v + 2 * i);
            acc += __match_any_sync(0xffffffffu, v + 3 * i);
        }
        else if (MODE == 3)
        {  // F2F.F64.F32
            d0 += (double)f0; d1 += (double)f1; d2 += (double)f2; d3 += (double)f3;
            f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f;
        }
        else if (MODE == 4)
        {  // F2F.F32.F64
            f0 += (float)d0; f1 += (float)d1; f2 += (float)d2; f3 += (float)d3;
            d0 += 1.0; d1 += 1.0; d2 += 1.0; d3 += 1.0;
        }
        else if (MODE == 5)
        {  // DFMA only
            d0 = fma(d0, 1.0000001, d1); d1 = fma(d1, 1.0000001, d2); d2 = fma(d2, 0.9999999, d3); d3 = fma(d3, 1.0000001, d0);
        }
        else if (MODE == 6)
        {  // POPC
            acc += __popc(v + i) + __popc(v ^ i) + __popc(v + 2 * i) + __popc(v + 3 * i);
        }
        else if (MODE == 7)
        {  // ballot (vote)
            acc += __ballot_sync(0xffffffffu, (v + i) & 1) + __ballot_sync(0xffffffffu, (v + i) & 2) +
                   __ballot_sync(0xffffffffu, (v + i) & 4) + __ballot_sync(0xffffffffu, (v + i) & 8);
        }
        else if (MODE == 8)
        {  // shfl
            acc += __shfl_xor_sync(0xffffffffu, v + i, 1) + __shfl_xor_sync(0xffffffffu, v + i, 2) +
                   __shfl_xor_sync(0xffffffffu, v + i, 4) + __shfl_xor_sync(0xffffffffu, v + i, 8);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (uint32_t)(d0 + d1 + d2 + d3) + (uint32_t)(f0 + f1 + f2 + f3);
}

template <int MODE>
void run(const char *name, int per_iter, uint32_t *d_out, int n_sm, double sm_hz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int ctas = n_sm * 4, threads = 256;
    k<MODE><<<ctas, threads>>>(d_out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<ctas, threads>>>(d_out, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM sub-partition: (4 CTAs * 8 warps / 4 SMSPs) = 8 warps, each ITER * per_iter instructions of interest
    const double inst_per_smsp = 8.0 * ITER * per_iter;
    const double cycles = ms * 1e-3 * sm_hz;
    printf("%-28s %8.3f ms  %6.2f cycles per warp-instruction per SMSP\n", name, ms, cycles / inst_per_smsp);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double hz = khz * 1e3;
    uint32_t *d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 4 * 256 * 4);
    printf("%s, %d SMs, %.0f MHz (nominal)\n", p.name, p.multiProcessorCount, hz / 1e6);
    run<0>("match.any (32 distinct)", 4, d, p.multiProcessorCount, hz);
    run<1>("match.any (8 groups)", 4, d, p.multiProcessorCount, hz);
    run<2>("match.any (all equal)", 4, d, p.multiProcessorCount, hz);
    run<3>("F2F.F64.F32 (+DADD,FADD)", 4, d, p.multiProcessorCount, hz);
    run<4>("F2F.F32.F64 (+FADD,DADD)", 4, d, p.multiProcessorCount, hz);
    run<5>("DFMA", 4, d, p.multiProcessorCount, hz);
    run<6>("POPC", 4, d, p.multiProcessorCount, hz);
    run<7>("VOTE.ballot", 4, d, p.multiProcessorCount, hz);
    run<8>("SHFL.bfly", 4, d, p.multiProcessorCount, hz);
    return 0;
}
