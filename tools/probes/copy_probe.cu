// Calibration probe (development aid): what does a plain streaming copy / 9-in-9-out SoA copy reach on this GPU?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CKR(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int UNROLL>
__global__ void copy_vec(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * UNROLL;
    const size_t stride = (size_t)gridDim.x * blockDim.x * UNROLL;
    for (; i + UNROLL <= n; i += stride) {
        double2 v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) v[j] = in[i + j];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) out[i + j] = v[j];
    }
}
// 9 planes in, 9 planes out, one element (8 B) per plane per thread: the LBM access shape without arithmetic/shift
__global__ void copy_soa9(const double* __restrict__ in, double* __restrict__ out, size_t plane) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane) return;
    double v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = in[k * plane + i];
#pragma unroll
    for (int k = 0; k < 9; ++k) out[k * plane + i] = v[k];
}
int main() {
    const size_t plane = 4096ull * 4096ull, bytes = plane * 9 * 8;
    double *a, *b;
    CKR(cudaMalloc(&a, bytes)); CKR(cudaMalloc(&b, bytes));
    CKR(cudaMemset(a, 0, bytes)); CKR(cudaMemset(b, 0, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* name, float ms, int reps) {
        printf("%-32s %.4f ms  %.0f GB/s\n", name, ms / reps, 2.0 * bytes / (ms / reps) / 1e6);
    };
    const int reps = 50; float ms;
    for (int blocks_per_sm : {4, 8, 16, 32}) {
        for (int i = 0; i < 3; ++i) copy_vec<2><<<148 * blocks_per_sm, 256>>>((double2*)a, (double2*)b, bytes / 16);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) copy_vec<2><<<148 * blocks_per_sm, 256>>>((double2*)a, (double2*)b, bytes / 16);
        cudaEventRecord(e1); CKR(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        char nm[64]; snprintf(nm, 64, "copy double2 x2, %d CTA/SM", blocks_per_sm); report(nm, ms, reps);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) copy_vec<4><<<148 * blocks_per_sm, 256>>>((double2*)a, (double2*)b, bytes / 16);
        cudaEventRecord(e1); CKR(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        snprintf(nm, 64, "copy double2 x4, %d CTA/SM", blocks_per_sm); report(nm, ms, reps);
    }
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) copy_soa9<<<(unsigned)((plane + 255) / 256), 256>>>(a, b, plane);
    cudaEventRecord(e1); CKR(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    report("soa9 copy (1 thread/node)", ms, reps);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) CKR(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice));
    cudaEventRecord(e1); CKR(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    report("cudaMemcpy D2D", ms, reps);
    return 0;
}
