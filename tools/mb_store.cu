// mb_store.cu -- store-path microbenchmark for the G/H value scatter (sm_100a).
//
// Question it answers: how fast can 888 tiles x 5 per-variable chunks (the
// layout pcx_fill writes) be written, as
//   M0  direct coalesced 8 B STG           (upper bound of the per-thread path)
//   M1  the round-1 pattern: one slot of a 117-slot period per thread, looping
//       over sections (misaligned ~184 B segments -> partial sectors)
//   M2  values staged in shared memory, then cp.async.bulk (TMA) stores
//   M3  M2 in S sub-tiles per CTA (store of sub-tile i overlaps fill of i+1)
//   M5  16 B vector STG
// against an empty kernel (launch floor), on a ring of buffers larger than L2.
//
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_store mb_store.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int NVAR = 5;
constexpr int NSEC = 38;          // sections per tile
constexpr int PA = 23;            // slots per section per variable (misaligned on purpose)
constexpr int CH = NSEC * PA;     // doubles per chunk = 874
constexpr int CHP = 876;          // padded chunk stride in smem (16 B multiple)

__global__ void k_empty(double* out) {}

__global__ void __launch_bounds__(128) k_direct(double* out, int ntiles) {
    const int tile = blockIdx.x;
    for (int a = 0; a < NVAR; ++a) {
        double* o = out + ((size_t)a * ntiles + tile) * CH;
        for (int j = threadIdx.x; j < CH; j += 128) o[j] = (double)(j + a);
    }
}

__global__ void __launch_bounds__(128) k_period(double* out, int ntiles) {
    const int tile = blockIdx.x;
    const int u = threadIdx.x;
    if (u >= NVAR * PA) return;
    const int a = u / PA, l = u - a * PA;
    double* o = out + ((size_t)a * ntiles + tile) * CH + l;
#pragma unroll 4
    for (int s = 0; s < NSEC; ++s) o[s * PA] = (double)(s + u);
}

__global__ void __launch_bounds__(128) k_vec16(double* out, int ntiles) {
    const int tile = blockIdx.x;
    for (int a = 0; a < NVAR; ++a) {
        double2* o = reinterpret_cast<double2*>(out + ((size_t)a * ntiles + tile) * CH);
        for (int j = threadIdx.x; j < CH / 2; j += 128) o[j] = make_double2((double)j, (double)a);
    }
}

__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, int bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
}

// S sub-tiles per CTA; each sub-tile stages NVAR sub-chunks and bulk-stores them.
template <int S>
__global__ void __launch_bounds__(128) k_bulk(double* out, int ntiles) {
    extern __shared__ __align__(128) double sm[];
    const int tile = blockIdx.x;
    constexpr int SUB = CH / S / 2 * 2;            // even -> 16 B multiple
    for (int s = 0; s < S; ++s) {
        double* buf = sm + (size_t)s * NVAR * ((CHP / S + 4) / 2 * 2);
        const int len = (s == S - 1) ? CH - SUB * (S - 1) : SUB;
        // emulate the period-slot fill: thread u owns slot (a, l), loops sections
        for (int a = 0; a < NVAR; ++a)
            for (int j = threadIdx.x; j < len; j += 128) buf[a * ((CHP / S + 4) / 2 * 2) + j] = (double)(j + a);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < NVAR) {
            const int a = threadIdx.x;
            bulk_store(out + ((size_t)a * ntiles + tile) * CH + (size_t)s * SUB, buf + a * ((CHP / S + 4) / 2 * 2), len * 8);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x < NVAR) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// read phase (9 doubles per node, 113 nodes) + staged bulk stores
template <int S>
__global__ void __launch_bounds__(128) k_rw(const double* __restrict__ in, double* out, int ntiles, int nn) {
    extern __shared__ __align__(128) double sm[];
    const int tile = blockIdx.x;
    double acc = 0.0;
    if (threadIdx.x < nn)
        for (int a = 0; a < 9; ++a) acc += in[((size_t)a * ntiles + tile) * nn + threadIdx.x];
    constexpr int SUB = CH / S / 2 * 2;
    for (int s = 0; s < S; ++s) {
        double* buf = sm + (size_t)s * NVAR * ((CHP / S + 4) / 2 * 2);
        const int len = (s == S - 1) ? CH - SUB * (S - 1) : SUB;
        for (int a = 0; a < NVAR; ++a)
            for (int j = threadIdx.x; j < len; j += 128) buf[a * ((CHP / S + 4) / 2 * 2) + j] = acc + (double)(j + a);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < NVAR) {
            const int a = threadIdx.x;
            bulk_store(out + ((size_t)a * ntiles + tile) * CH + (size_t)s * SUB, buf + a * ((CHP / S + 4) / 2 * 2), len * 8);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x < NVAR) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <class F>
static float timeit(F launch, int steps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 20; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < steps; ++i) launch(i);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms * 1000.f / steps;
}

int main(int argc, char** argv) {
    const int R = 6, steps = 300;
    for (int ntiles : {888, 1776, 3552}) {
        const size_t n = (size_t)NVAR * ntiles * CH;
        double* bufs[R]; double* ins[R];
        for (int r = 0; r < R; ++r) {
            CK(cudaMalloc(&bufs[r], (n + 16) * 8));   
            CK(cudaMalloc(&ins[r], (size_t)9 * ntiles * 128 * 8));
            CK(cudaMemset(ins[r], 0, (size_t)9 * ntiles * 128 * 8));
        }
        const double mb = n * 8 / 1e6;
        printf("ntiles %d  bytes written per launch %.2f MB\n", ntiles, mb);
        float t;
        t = timeit([&](int i) { k_empty<<<ntiles, 128>>>(bufs[i % R]); }, steps);
        printf("  empty            %7.2f us\n", t);
        t = timeit([&](int i) { k_direct<<<ntiles, 128>>>(bufs[i % R], ntiles); }, steps);
        printf("  M0 direct 8B     %7.2f us  %7.0f GB/s\n", t, mb / t * 1e3);
        t = timeit([&](int i) { k_vec16<<<ntiles, 128>>>(bufs[i % R], ntiles); }, steps);
        printf("  M5 vec 16B       %7.2f us  %7.0f GB/s\n", t, mb / t * 1e3);
        t = timeit([&](int i) { k_period<<<ntiles, 128>>>(bufs[i % R], ntiles); }, steps);
        printf("  M1 period slots  %7.2f us  %7.0f GB/s\n", t, mb / t * 1e3);
        const int smem1 = NVAR * (CHP + 16) * 8;
        CK(cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
        CK(cudaFuncSetAttribute(k_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1 + 256));
        CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1 + 512));
        CK(cudaFuncSetAttribute(k_rw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
        CK(cudaFuncSetAttribute(k_rw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1 + 256));
        t = timeit([&](int i) { k_bulk<1><<<ntiles, 128, smem1>>>(bufs[i % R], ntiles); }, steps);
        printf("  M2 bulk S=1      %7.2f us  %7.0f GB/s\n", t, mb / t * 1e3);
        t = timeit([&](int i) { k_bulk<2><<<ntiles, 128, smem1 + 256>>>(bufs[i % R], ntiles); }, steps);
        printf("  M3 bulk S=2      %7.2f us  %7.0f GB/s\n", t, mb / t * 1e3);
        t = timeit([&](int i) { k_bulk<4><<<ntiles, 128, smem1 + 512>>>(bufs[i % R], ntiles); }, steps);
        printf("  M3 bulk S=4      %7.2f us  %7.0f GB/s\n", t, mb / t * 1e3);
        t = timeit([&](int i) { k_rw<1><<<ntiles, 128, smem1>>>(ins[i % R], bufs[i % R], ntiles, 113); }, steps);
        printf("  M4 read+bulk S=1 %7.2f us  %7.0f GB/s (+%.1f MB read)\n", t, mb / t * 1e3, 9.0 * ntiles * 113 * 8 / 1e6);
        t = timeit([&](int i) { k_rw<2><<<ntiles, 128, smem1 + 256>>>(ins[i % R], bufs[i % R], ntiles, 113); }, steps);
        printf("  M4 read+bulk S=2 %7.2f us  %7.0f GB/s\n", t, mb / t * 1e3);
        for (int r = 0; r < R; ++r) { CK(cudaFree(bufs[r])); CK(cudaFree(ins[r])); }
    }
    // plain cudaMemsetAsync of the same size as reference
    {
        const size_t bytes = 35200000;
        double* b[R];
        for (int r = 0; r < R; ++r) CK(cudaMalloc(&b[r], bytes));
        float t = timeit([&](int i) { cudaMemsetAsync(b[i % R], 0, bytes); }, steps);
        printf("cudaMemsetAsync 35.2 MB: %7.2f us %7.0f GB/s\n", t, bytes / 1e6 / t * 1e3);
    }
    return 0;
}
