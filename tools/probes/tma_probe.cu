// Standalone TMA bisect probe (development aid): loads one 2-D box with cp.async.bulk.tensor and copies it out.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../latticeboltzmannsimulations_b200/csrc/lbm_tma.cuh"
using namespace lbm;
#define CKR(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <typename T, int BX, int BY>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, T* out, int c0, int c1) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)raw + 127) & ~uintptr_t(127));
    T* tile = (T*)base;
    uint64_t* bar = (uint64_t*)(base + BX * BY * sizeof(T));
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, BX * BY * sizeof(T));
        tma_load_2d(tile, &tmap, c0, c1, bar);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T, int BX, int BY>
int run(const char* name, CUtensorMapDataType dt) {
    const int W = 512, H = 64;
    std::vector<T> h(W * H);
    for (int i = 0; i < W * H; ++i) h[i] = (T)i;
    T *d, *o;
    CKR(cudaMalloc(&d, W * H * sizeof(T)));
    CKR(cudaMalloc(&o, BX * BY * sizeof(T)));
    CKR(cudaMemcpy(d, h.data(), W * H * sizeof(T), cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CKR(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    CUtensorMap tm;
    cuuint64_t dims[2] = {W, H}; cuuint64_t strides[1] = {W * sizeof(T)};
    cuuint32_t box[2] = {BX, BY}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("%s encode -> %d\n", name, (int)r); fflush(stdout);
    if (r) return 1;
    size_t smem = BX * BY * sizeof(T) + 8 + 128;
    CKR(cudaFuncSetAttribute(probe<T, BX, BY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int c0 : {0, 8, -(int)(16 / sizeof(T)), 16 * 31, (int)(16 / sizeof(T)) * 127}) {
        probe<T, BX, BY><<<1, 128, smem>>>(tm, o, c0, 3);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%s c0=%d -> %s\n", name, c0, cudaGetErrorString(e)); fflush(stdout);
        if (e != cudaSuccess) return 1;
        std::vector<T> res(BX * BY);
        CKR(cudaMemcpy(res.data(), o, BX * BY * sizeof(T), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int y = 0; y < BY; ++y) for (int x = 0; x < BX; ++x) {
            int gx = c0 + x, gy = 3 + y;
            T want = (gx < 0 || gx >= W) ? (T)0 : (T)(gy * W + gx);
            if (res[y * BX + x] != want) ++bad;
        }
        printf("%s c0=%d mismatches %d\n", name, c0, bad);
    }
    return 0;
}

int main() {
    if (run<float, 64, 4>("f32 64x4", CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return 1;
    if (run<float, 256, 4>("f32 256x4", CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return 1;
    if (run<double, 64, 4>("f64 64x4", CU_TENSOR_MAP_DATA_TYPE_FLOAT64)) return 1;
    if (run<double, 128, 4>("f64 128x4", CU_TENSOR_MAP_DATA_TYPE_FLOAT64)) return 1;
    printf("PROBE_OK\n");
    return 0;
}
