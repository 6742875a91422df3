// hs_compat.cuh -- lets the device functions of hs_tmat.cuh / hs_ampl.cuh also compile as plain host C++
// (HS_HOST_EMU) for the single-lane logic tests under tests/emu/.  The product always builds with nvcc;
// nothing in the shipped library defines HS_HOST_EMU.
#pragma once
#ifdef HS_HOST_EMU
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cmath>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline void __syncwarp() {}
static inline void __syncthreads() {}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
using std::isfinite;
#else
#include <cuda_runtime.h>
#endif
