// cuda_runtime.h -- TEST INFRASTRUCTURE: a host stand-in for the CUDA runtime API, just wide enough for pangene_b200/csrc/pgx_*.cuh.
// Together with simt_emu.hpp (CUDA threads = OS threads, one CTA at a time) it lets tests/emu/build.py compile the WHOLE of libpgx -- engine,
// kernels, C ABI -- for the host as tests/emu/_build/libpgx_emu.so: "device" memory is host memory, streams execute in program order
// (every call completes before it returns, which is one legal schedule of the stream/event graph the engine builds), events are timestamps.
// Nothing here is used by the product; libpgx.so itself has no CPU path.
#pragma once
#include "simt_emu.hpp"
#include <chrono>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef struct emu_stream_s { int id; } *cudaStream_t;
typedef struct emu_event_s { double t; } *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess? "no error" : "emulation error"; }
static inline const char *cudaGetErrorName(cudaError_t e) { return e == cudaSuccess? "cudaSuccess" : "cudaErrorEmu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof(*p)); strcpy(p->name, "SIMT shim (host)"); p->major = 10; p->minor = 0; p->multiProcessorCount = 148; p->totalGlobalMem = (size_t)8 << 30; return cudaSuccess; }
template<class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { *p = (T*)malloc(bytes? bytes : 1); return *p? cudaSuccess : cudaErrorEmu; }
template<class T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { *p = (T*)malloc(bytes? bytes : 1); return *p? cudaSuccess : cudaErrorEmu; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { if (n) memmove(dst, src, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(dst, src, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { if (n) memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emu_stream_s{0}; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new emu_stream_s{0}; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event_s{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event_s{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template<class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

namespace emu {
// a launch of any shape: blocks that are not a multiple of 32 threads (the <<<1, 1>>> bookkeeping kernels) run their threads one after the
// other -- legal for kernels without barriers or warp collectives, which is what such launches are
template<class K> static void launch_any(long grid, int block, K kernel)
{
	static std::mutex one_launch_at_a_time;                // the shim's CTA state is global; the engine launches from its calling thread only
	std::lock_guard<std::mutex> lk(one_launch_at_a_time);
	if (grid <= 0 || block <= 0) return;
	if (block % 32 == 0) { launch((int)grid, block, kernel); return; }
	blockDim.x = block; gridDim.x = (int)grid;
	for (int b = 0; b < (int)grid; ++b) for (int t = 0; t < block; ++t) { blockIdx.x = b; threadIdx.x = t; kernel(); }
}
}
