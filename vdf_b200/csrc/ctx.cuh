// Library context shared by the C-ABI translation units (api_*.cu).  One process drives one GPU.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../../include/vdfgpu.h"
#include "launch.cuh"
#include "msm.cuh"

namespace vdf {

struct ArgError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct StateError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// one in-flight host-scalar MSM of the asynchronous API (vdfgpu_msm_submit / vdfgpu_msm_wait)
struct AsyncSlot {
  fe* d_scalars = nullptr;
  size_t cap = 0;
  jac_t* d_out = nullptr;
  cudaEvent_t copied = nullptr, done = nullptr;
  bool busy = false;
};
constexpr int VDF_ASYNC_SLOTS = 4;

struct Context {
  std::mutex mu;          // serialises library calls (re-entrant use from several host threads)
  bool ready = false;
  int device = -1;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // stream in use (own_stream or the caller's)
  cudaStream_t copy_stream = nullptr;       // H2D of scalar chunks, overlapped with compute (vdfgpu_msm)
  cudaEvent_t chunk_ev[8] = {};             // chunk k of the scalars has arrived
  cudaEvent_t start_ev = nullptr;
  AsyncSlot slots[VDF_ASYNC_SLOTS];
  uint64_t launches = 0;
  StageProfile prof;      // stage timing of the most recent MSM (vdfgpu_profile_*)
};

Context& ctx();
void set_error(const std::string& msg);
void upload_constants_r1cs();   // api_r1cs.cu's copy of the constant-memory tables
void require_ready();   // throws StateError unless vdfgpu_init succeeded (initialises lazily on device 0)

// RAII device buffer on the context stream (stream-ordered)
template <class T>
struct DevBuf {
  T* p = nullptr;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(size_t count, cudaStream_t st) : s(st) {
    void* q = nullptr;
    VDF_CUDA_CHECK(cudaMallocAsync(&q, count * sizeof(T) + 16, st));
    p = reinterpret_cast<T*>(q);
  }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), s(o.s) { o.p = nullptr; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p; s = o.s; o.p = nullptr;
    }
    return *this;
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
  }
  ~DevBuf() { release(); }
};

inline void h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes) VDF_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
}
inline void d2h(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes) VDF_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
}
inline void d2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes) VDF_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
}

// wraps a C-ABI body: locks, maps exceptions to status codes + thread-local message
template <class Body>
int guarded(Body body) {
  try {
    std::lock_guard<std::mutex> lk(ctx().mu);
    body();
    return VDFGPU_OK;
  } catch (const ArgError& e) {
    set_error(e.what());
    return VDFGPU_ERR_ARG;
  } catch (const StateError& e) {
    set_error(e.what());
    return VDFGPU_ERR_STATE;
  } catch (const std::exception& e) {
    set_error(e.what());
    return VDFGPU_ERR_CUDA;
  }
}

}  // namespace vdf

// handles ------------------------------------------------------------------------------------------
struct vdfgpu_gens {
  int curve = 0;
  size_t n = 0;            // points per level
  uint32_t flags = 0;
  uint32_t c = 0;          // window bits of the table (table mode)
  uint32_t W = 0;          // levels stored (1 in plain mode)
  vdf::affine_t* pts = nullptr;  // [W][n]
};

// internal cross-TU entry: MSM over device scalars into a device point, on the context stream
namespace vdf {
void msm_on_device(vdfgpu_gens* g, size_t first, const fe* d_scalars, size_t n, jac_t* d_out, bool is_mont);
void msm_batch_on_device(vdfgpu_gens* g, const fe* const* d_scalars, const size_t* lens, uint32_t k, jac_t* d_out);
}
