// simt_emu.hpp -- TEST INFRASTRUCTURE: a minimal SIMT execution shim that lets tests/emu/build.py compile the sweep kernels (and, for
// tests/test_prim_emu.py, the radix-sort / scan / segmented-max kernels of pgx_sort.cuh and pgx_scan.cuh) of
// pangene_b200/csrc/pgx_kernels.cuh (the text between its [emu-begin sweep] / [emu-end sweep] markers, unchanged) for the host,
// so that the control flow of the production kernel -- tile staging, sparse table, job bucketing, warp-cooperative parking --
// can be checked against the plain thread-per-hit variant without a GPU (tests/test_sweep_emu.py).  One CTA runs at a time;
// each CUDA thread is an OS thread, __syncthreads() and the warp collectives are barriers.  Nothing here is used by the product.
#pragma once
#include <pthread.h>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <mutex>
#include <vector>

#define PGX_SIMT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct int2 { int32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int4 { int32_t x, y, z, w; };
static inline int2 make_int2(int32_t x, int32_t y) { int2 r; r.x = x; r.y = y; return r; }
struct emu_dim3 { int x = 0, y = 0, z = 0; };
static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;
alignas(128) static unsigned char emu_dyn_smem[256 << 10];

namespace emu {
constexpr int MAX_THREADS = 1024;
static pthread_barrier_t cta_bar, warp_bar[MAX_THREADS / 32];
static uint64_t warp_slot[MAX_THREADS / 32][32];
static int cta_red[2];
static inline int tid() { return threadIdx.x; }
static inline void warp_sync() { pthread_barrier_wait(&warp_bar[tid() >> 5]); }
template<class T> static inline uint64_t to_u64(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&u, &v, sizeof(T)); return u; }
template<class T> static inline T from_u64(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
// every lane deposits, everybody reads, nobody overwrites before all have read
template<class T, class F> static inline T exchange(T v, F pick)
{
	const int w = tid() >> 5, lane = tid() & 31;
	warp_slot[w][lane] = to_u64(v);
	warp_sync();
	T r = pick(warp_slot[w], lane);
	warp_sync();
	return r;
}
}

static inline void __syncthreads() { pthread_barrier_wait(&emu::cta_bar); }
static inline int __syncthreads_or(int p)
{
	emu::cta_red[0] = 0;   // (every thread writes the same 0 before the barrier)
	__syncthreads();
	if (p) __atomic_store_n(&emu::cta_red[0], 1, __ATOMIC_RELAXED);
	__syncthreads();
	int r = emu::cta_red[0];
	__syncthreads();
	return r;
}
static inline uint32_t __ballot_sync(uint32_t, bool p)
{
	return emu::exchange<uint32_t>(p? 1u : 0u, [](const uint64_t *s, int) { uint32_t b = 0; for (int l = 0; l < 32; ++l) if (s[l]) b |= 1u << l; return b; });
}
static inline bool __any_sync(uint32_t m, bool p) { return __ballot_sync(m, p) != 0; }
template<class T> static inline T __shfl_sync(uint32_t, T v, int src) { return emu::exchange<T>(v, [src](const uint64_t *s, int) { return emu::from_u64<T>(s[src & 31]); }); }
template<class T> static inline T __shfl_xor_sync(uint32_t, T v, int m) { return emu::exchange<T>(v, [m](const uint64_t *s, int lane) { return emu::from_u64<T>(s[(lane ^ m) & 31]); }); }
template<class T> static inline T __shfl_up_sync(uint32_t, T v, int d) { return emu::exchange<T>(v, [d](const uint64_t *s, int lane) { return emu::from_u64<T>(s[lane >= d? lane - d : lane]); }); }
static inline uint32_t __match_any_sync(uint32_t, uint32_t v)
{
	return emu::exchange<uint32_t>(v, [](const uint64_t *s, int lane) { uint32_t b = 0; for (int l = 0; l < 32; ++l) if ((uint32_t)s[l] == (uint32_t)s[lane]) b |= 1u << l; return b; });
}
static inline void __syncwarp() { emu::warp_sync(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x == 0? 32 : __builtin_clz((uint32_t)x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
template<class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template<class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template<class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template<class T> static inline T atomicCAS(T *p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
template<class T> static inline T atomicMax(T *p, T v) { T old = __atomic_load_n(p, __ATOMIC_RELAXED); while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { } return old; }
template<class T> static inline T atomicMin(T *p, T v) { T old = __atomic_load_n(p, __ATOMIC_RELAXED); while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { } return old; }

namespace emu {
// run kernel(args...) over grid x block threads; one CTA at a time, the OS threads persist over the CTAs
template<class K> static void launch(int grid, int block, K kernel)
{
	blockDim.x = block; gridDim.x = grid;
	const int n_warp = (block + 31) / 32;
	if (block % 32) { fprintf(stderr, "emu: block size must be a multiple of 32\n"); abort(); }
	pthread_barrier_init(&cta_bar, nullptr, block);
	for (int w = 0; w < n_warp; ++w) pthread_barrier_init(&warp_bar[w], nullptr, 32);
	std::vector<std::thread> th;
	for (int t = 0; t < block; ++t) th.emplace_back([=] {
		threadIdx.x = t;
		for (int b = 0; b < grid; ++b) {
			blockIdx.x = b;
			kernel();
			pthread_barrier_wait(&cta_bar);   // static __shared__ state belongs to one CTA at a time
		}
	});
	for (auto &x : th) x.join();
	pthread_barrier_destroy(&cta_bar);
	for (int w = 0; w < n_warp; ++w) pthread_barrier_destroy(&warp_bar[w]);
}
}
