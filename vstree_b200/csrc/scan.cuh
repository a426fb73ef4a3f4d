// scan.cuh -- warp-shuffle and block-level scan / reduce primitives used by every kernel family.
#pragma once

#include "common.cuh"

#define FULL_MASK 0xffffffffu

template <typename T> __device__ __forceinline__ T warp_inclusive_sum(T v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    T t = __shfl_up_sync(FULL_MASK, v, o);
    if (lane >= o)
    {
      v += t;
    }
  }
  return v;
}

template <typename T> __device__ __forceinline__ T warp_inclusive_max(T v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    T t = __shfl_up_sync(FULL_MASK, v, o);
    if (lane >= o && t > v)
    {
      v = t;
    }
  }
  return v;
}

// Exclusive prefix sum over the first `BLOCK` threads of the block (BLOCK a multiple of 32, <= 1024).
// `warp_sums` is shared scratch of BLOCK/32 + 1 elements.  Every thread of the block must call it
// (threads with threadIdx.x >= BLOCK pass 0 and get garbage).  Returns the exclusive sum; *total gets the sum.
template <typename T, int BLOCK> __device__ __forceinline__ T block_exclusive_sum(T v, T *warp_sums, T *total)
{
  constexpr int NW = BLOCK / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = warp_inclusive_sum(v);
  if (lane == 31 && warp < NW)
  {
    warp_sums[warp] = incl;
  }
  __syncthreads();
  if (warp == 0)
  {
    T s = lane < NW ? warp_sums[lane] : (T) 0;
    T si = warp_inclusive_sum(s);
    if (lane < NW)
    {
      warp_sums[lane] = si - s;
    }
    if (lane == NW - 1)
    {
      warp_sums[NW] = si;
    }
  }
  __syncthreads();
  T res = (warp < NW ? warp_sums[warp] : (T) 0) + incl - v;
  if (total != nullptr)
  {
    *total = warp_sums[NW];
  }
  __syncthreads();  // scratch may be reused by the caller right away
  return res;
}

// Inclusive running maximum over the threads of the block (all BLOCK threads participate).
template <typename T, int BLOCK> __device__ __forceinline__ T block_inclusive_max(T v, T *warp_scratch)
{
  constexpr int NW = BLOCK / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = warp_inclusive_max(v);
  if (lane == 31)
  {
    warp_scratch[warp] = incl;
  }
  __syncthreads();
  if (warp == 0)
  {
    T s = lane < NW ? warp_scratch[lane] : (T) 0;
    T si = warp_inclusive_max(s);
    if (lane < NW)
    {
      warp_scratch[lane] = si;
    }
  }
  __syncthreads();
  if (warp > 0)
  {
    T prev = warp_scratch[warp - 1];
    if (prev > incl)
    {
      incl = prev;
    }
  }
  __syncthreads();
  return incl;
}

template <typename T> __device__ __forceinline__ T warp_reduce_sum(T v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    v += __shfl_down_sync(FULL_MASK, v, o);
  }
  return v;
}

template <typename T> __device__ __forceinline__ T warp_reduce_max(T v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    T t = __shfl_down_sync(FULL_MASK, v, o);
    if (t > v)
    {
      v = t;
    }
  }
  return v;
}

