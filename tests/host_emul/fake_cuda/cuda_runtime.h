// Fake <cuda_runtime.h> for the HOST-EMULATED build of the library (tests/host_emul/build_emul_lib.py): "device" memory
// is host memory, streams and events are inert, every call succeeds immediately. TEST INFRASTRUCTURE ONLY — it lets
// the CPU test-suite run the product's host logic (drivers, planners, launch code) together with its kernels (run by
// emul_lib_prelude.hpp one emulated thread at a time) against the oracle. Nothing under numericalnim_b200/ refers to it.
#pragma once
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "not supported under host emulation"); }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

struct dim3 { unsigned x, y, z; explicit dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };

struct EmulStreamObj { int id; };
typedef EmulStreamObj* cudaStream_t;
struct EmulEventObj { std::chrono::steady_clock::time_point t; };
typedef EmulEventObj* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocMapped = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };

inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; int l2CacheSize; int clockRate; };
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 1; p->l2CacheSize = 126 << 20; p->clockRate = 1965000; return cudaSuccess; }  // one SM: cooperative grids are one CTA
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }

inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new EmulStreamObj{1}; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmulEventObj{std::chrono::steady_clock::now()}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }

template <class T> inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  void* q = nullptr;   // exactly `bytes` (no padding): under AddressSanitizer a kernel reading one element past the end is caught
  if (posix_memalign(&q, 256, bytes ? bytes : 1) != 0) q = nullptr;
  *p = static_cast<T*>(q);
  return q ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
template <class T> inline cudaError_t cudaHostAlloc(T** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
template <class T> inline cudaError_t cudaHostGetDevicePointer(T** dev, T* host, unsigned) { *dev = host; return cudaSuccess; }  // mapped memory: same address
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }

typedef void* cudaMemPool_t;
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }

struct cudaIpcMemHandle_t { char reserved[64]; };
// ranks are threads of one process (multi_rank_emul.py): a handle is the pointer itself
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** out, cudaIpcMemHandle_t h, unsigned) { std::memcpy(out, h.reserved, sizeof(void*)); return *out ? cudaSuccess : cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

template <class K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 1; return cudaSuccess; }
