// Host-side helpers shared by the translation units of libvmm_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vmm_b200.h"

namespace vmm {

// Per-thread error text; every extern "C" entry point returns 0 or a negative VMM_E_* code
// and never throws.
void set_error(const char* fmt, ...);
int device_sm_count();

#define VMM_CHECK_CUDA(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::vmm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VMM_E_CUDA;                                                                          \
    }                                                                                             \
  } while (0)

#define VMM_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::vmm::set_error(__VA_ARGS__);    \
      return VMM_E_INVALID;             \
    }                                   \
  } while (0)

#define VMM_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != 0) return _r;      \
  } while (0)

// Every kernel launch site ends with VMM_LAUNCHED(): counts the launch (vmm_launch_count) and
// surfaces launch-configuration errors.
void count_launch();
#define VMM_LAUNCHED()                     \
  do {                                     \
    ::vmm::count_launch();                 \
    VMM_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace vmm
