// TEST INFRASTRUCTURE (tests/host_emu): a stand-in for <cuda_runtime.h> that lets g++ compile the product's kernel
// SOURCE (csrc/lbm_device.cuh, lbm_kernels.cuh, lbm_aa.cuh) as plain host code, so that the CPU-only test suite can run
// the very statements of the one-thread-per-node kernels -- index arithmetic, wall flags, side buffers, collision --
// node by node in a loop and compare them with the oracle.  Nothing in the product includes this file; the product has
// no CPU path (lbm_create fails without a CUDA device).  Only kernels without shared memory, barriers or shuffles are
// executed through it; the shims for those exist so that the rest of the header parses.
#pragma once
#include <cmath>
#include <math.h>       // the C++ wrapper: ::fabs / ::sqrt get their float overloads, as in CUDA device code
#include <cstdint>
#include <cstdlib>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__

struct emu_uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static thread_local emu_uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }

// IEEE round-to-nearest fused multiply-add and the packed fp32 forms (each half = the scalar operation)
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }

// never executed by the emulated kernels
static inline void __syncthreads() { std::abort(); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int) { std::abort(); return v; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int) { std::abort(); return v; }
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p += v; return o; }
