// Host stand-ins for the CUDA builtins used by chain_kernel.cuh / philox.cuh, so
// that tests/host_emu can compile the device state machine with g++ (T = 1) and
// diff it against the oracle without a GPU.  Test infrastructure only.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cmath>
#include <cstdint>

inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
inline int __shfl_sync(unsigned, int v, int) { return v; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline void __syncthreads() {}
using std::isinf;
using std::isnan;
inline void __syncwarp() {}
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { unsigned int o = *p; *p += v; return o; }
