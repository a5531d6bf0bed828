// common.cuh -- shared device helpers for libreid_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include "../../include/reid_b200.h"

#define REID_CHECK_LAUNCH()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return REID_E_CUDA;               \
  } while (0)

#define REID_NEG_INF (-INFINITY)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit global load that does not allocate in L1 (read-once data)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// The ONE fp32 dot-product routine of the library.  reid_pos_scores, reid_rescore_topk and
// reid_retrieve_exact all call it, so a (query, gallery-row) pair has bit-identical scores in
// every kernel.  Order: lane l accumulates elements 4l..4l+3 of each 128-element block with
// sequential FMAs, then a xor-butterfly (16,8,4,2,1).  Every lane returns the full sum.
// d must be a multiple of 4; rows 16-byte aligned.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_dot(const float* __restrict__ a, const float* __restrict__ b,
                                          int d, int lane) {
  float acc = 0.f;
  if (d == 512) {
    // the feature width of the whole reference: all eight 128-bit loads are issued before the first FMA (one
    // memory round trip per dot instead of four); the FMA order is the one of the generic loop, so the result is
    // bit-identical
    float4 x[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[i] = *reinterpret_cast<const float4*>(a + lane * 4 + 128 * i);
      y[i] = *reinterpret_cast<const float4*>(b + lane * 4 + 128 * i);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc = fmaf(x[i].x, y[i].x, acc); acc = fmaf(x[i].y, y[i].y, acc);
      acc = fmaf(x[i].z, y[i].z, acc); acc = fmaf(x[i].w, y[i].w, acc);
    }
    return warp_sum(acc);
  }
  for (int c = lane * 4; c < d; c += 128) {
    const float4 x = *reinterpret_cast<const float4*>(a + c);
    const float4 y = *reinterpret_cast<const float4*>(b + c);
    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
    acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
  }
  return warp_sum(acc);
}

// two dots against the same `a` with all loads of both rows in flight (d == 512); each result is bit-identical
// to warp_dot(a, b?, 512, lane)
__device__ __forceinline__ void warp_dot2_512(const float* __restrict__ a, const float* __restrict__ b0,
                                              const float* __restrict__ b1, int lane, float& r0, float& r1) {
  float4 x[4], y0[4], y1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[i] = *reinterpret_cast<const float4*>(a + lane * 4 + 128 * i);
    y0[i] = *reinterpret_cast<const float4*>(b0 + lane * 4 + 128 * i);
    y1[i] = *reinterpret_cast<const float4*>(b1 + lane * 4 + 128 * i);
  }
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a0 = fmaf(x[i].x, y0[i].x, a0); a0 = fmaf(x[i].y, y0[i].y, a0); a0 = fmaf(x[i].z, y0[i].z, a0); a0 = fmaf(x[i].w, y0[i].w, a0);
    a1 = fmaf(x[i].x, y1[i].x, a1); a1 = fmaf(x[i].y, y1[i].y, a1); a1 = fmaf(x[i].z, y1[i].z, a1); a1 = fmaf(x[i].w, y1[i].w, a1);
  }
  r0 = warp_sum(a0); r1 = warp_sum(a1);
}

// total order used by every top-list in the library: score descending, index ascending
__device__ __forceinline__ bool ranks_before(float sa, int ia, float sb, int ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}

__host__ __device__ __forceinline__ int64_t reid_min64(int64_t a, int64_t b) { return a < b ? a : b; }
