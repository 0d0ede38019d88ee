// pgx_util.cuh -- error handling, device buffers, launch accounting (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <map>
#include <unordered_map>
#include <mutex>

#define PGX_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	fprintf(stderr, "[pgx] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); abort(); } } while (0)

#define PGX_FATAL(...) do { fprintf(stderr, "[pgx] FATAL: " __VA_ARGS__); fprintf(stderr, "\n"); abort(); } while (0)

namespace pgx {

// launch accounting: every kernel of this library goes through PGX_LAUNCH
struct LaunchStats { uint64_t launches = 0; };
extern LaunchStats g_launch;

#define PGX_LAUNCH(kernel, grid, block, smem, stream, ...) do { \
	kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
	++pgx::g_launch.launches; \
	cudaError_t e_ = cudaGetLastError(); \
	if (e_ != cudaSuccess) { fprintf(stderr, "[pgx] launch of %s failed at %s:%d: %s\n", #kernel, __FILE__, __LINE__, cudaGetErrorString(e_)); abort(); } } while (0)

static inline unsigned grid_for(int64_t n, int block) { int64_t g = (n + block - 1) / block; return (unsigned)(g > 0? g : 1); }

// Device memory is recycled across sessions: cudaMalloc/cudaFree of ~60 buffers cost ~100-200 ms per pg_post_process/
// pg_graph_gen pair (measured), which is more than the kernels take.  Freed blocks are kept in a size-indexed list and
// handed out again when a request of similar size (within 25 %) arrives; pgx_trim() really frees them.
struct DevCache {
	std::multimap<size_t, void*> free_;
	std::unordered_map<void*, size_t> size_of;
	std::mutex mu;
	uint64_t n_malloc = 0;            // driver allocations so far (a steady-state run makes none)
	static size_t round(size_t b) { size_t g = b >= ((size_t)1 << 20)? ((size_t)1 << 20) : 512; return (b + g - 1) / g * g; }
	void *get(size_t bytes)
	{
		bytes = round(bytes? bytes : 1);
		std::lock_guard<std::mutex> lk(mu);
		auto it = free_.lower_bound(bytes);
		if (it != free_.end() && it->first <= bytes + bytes / 4 + 4096) { void *p = it->second; free_.erase(it); return p; }
		void *p = nullptr;
		++n_malloc;
		cudaError_t e = cudaMalloc(&p, bytes);
		if (e != cudaSuccess) {   // give cached blocks back and retry once
			cudaGetLastError();
			for (auto &kv : free_) { size_of.erase(kv.second); cudaFree(kv.second); }
			free_.clear();
			PGX_CHECK(cudaMalloc(&p, bytes));
		}
		size_of[p] = bytes;
		return p;
	}
	void put(void *p)
	{
		if (!p) return;
		std::lock_guard<std::mutex> lk(mu);
		auto it = size_of.find(p);
		if (it == size_of.end()) { cudaFree(p); return; }   // not one of ours: never cache a block of unknown size
		free_.emplace(it->second, p);
	}
	void trim() { std::lock_guard<std::mutex> lk(mu); for (auto &kv : free_) { size_of.erase(kv.second); cudaFree(kv.second); } free_.clear(); }
	static DevCache &get_cache() { static DevCache c; return c; }
};

// typed device buffer
template<class T> struct DBuf {
	T *p = nullptr; size_t n = 0;
	DBuf() = default;
	DBuf(const DBuf&) = delete;
	DBuf &operator=(const DBuf&) = delete;
	~DBuf() { release(); }
	void alloc(size_t count) { release(); n = count; if (count) p = (T*)DevCache::get_cache().get(count * sizeof(T)); }
	void ensure(size_t count) { if (count > n) alloc(count + count / 8 + 16); }
	void release() { if (p) DevCache::get_cache().put(p); p = nullptr; n = 0; }
	void zero(cudaStream_t s) { if (n) PGX_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
	void fill_byte(int v, size_t count, cudaStream_t s) { if (count) PGX_CHECK(cudaMemsetAsync(p, v, count * sizeof(T), s)); }
	void swap(DBuf &o) { T *tp = p; p = o.p; o.p = tp; size_t tn = n; n = o.n; o.n = tn; }
};

// bump allocator for per-stage temporaries; one cudaMalloc for the whole session
struct Arena {
	char *base = nullptr; size_t cap = 0, top = 0;
	void init(size_t bytes) { release(); cap = bytes; top = 0; base = (char*)DevCache::get_cache().get(bytes); }
	void release() { if (base) DevCache::get_cache().put(base); base = nullptr; cap = top = 0; }
	template<class T> T *get(size_t count) {
		size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
		if (top + bytes > cap) PGX_FATAL("scratch arena exhausted: need %zu more bytes (cap %zu)", bytes, cap);
		T *r = (T*)(base + top); top += bytes; return r;
	}
	size_t mark() const { return top; }
	void reset(size_t m) { top = m; }
	~Arena() { release(); }
};

} // namespace pgx
