// pgx_hostpipe.hpp -- the host side of the host<->HBM boundary: a small persistent thread pool and a ring of
// pinned staging buffers through which the AoS <-> packed-record conversion and the PCIe copies are pipelined.
//
// Why: the reference hands us 88-byte pg_hit_t records in pageable malloc'd arrays and wants 88-byte records back
// in fresh malloc'd arrays (graph.c:45 / hit.c:62 free() them).  Copying those arrays verbatim costs 2 x 88 B/hit
// over PCIe from/into pageable memory (measured: 245 ms up + 460 ms down for 20 M hits, 90 % of the end-to-end
// time).  Instead: host threads pack the 12 words the device needs (48 B/hit) straight into pinned chunks that the
// copy engine drains while the next chunk is being packed; on the way back only the 6 words that changed
// (24 B/hit: permutation, rank, score_dom, pid_dom, pid_dom0, flag word) cross the bus and host threads rebuild
// the records.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include "pgx_util.cuh"

namespace pgx {

class ThreadPool {
	// One Job per parallel_for call.  A worker captures the shared_ptr of the job that is current when it wakes, under the mutex, and
	// from then on touches only that object: a worker that wakes late finds either no job or a newer one, never the cursor of one call
	// paired with the function of another (the function object lives on the caller's stack and dies when its call returns).
	struct Job {
		const std::function<void(int64_t)> *f;
		int64_t n;
		std::atomic<int64_t> next{0}, done{0};
		Job(const std::function<void(int64_t)> *f_, int64_t n_) : f(f_), n(n_) { }
	};
	std::vector<std::thread> workers;
	std::mutex mu;
	std::condition_variable cv_work, cv_done;
	std::shared_ptr<Job> job;
	uint64_t epoch = 0;
	bool stop = false;
	static void drain(Job &j)
	{
		int64_t mine = 0;
		for (int64_t i; (i = j.next.fetch_add(1)) < j.n;) { (*j.f)(i); ++mine; }
		if (mine) j.done.fetch_add(mine);
	}
	void loop()
	{
		uint64_t seen = 0;
		for (;;) {
			std::shared_ptr<Job> j;
			{
				std::unique_lock<std::mutex> lk(mu);
				cv_work.wait(lk, [&] { return stop || epoch != seen; });
				if (stop) return;
				seen = epoch; j = job;
			}
			if (!j) continue;
			drain(*j);
			if (j->done.load() >= j->n) { std::unique_lock<std::mutex> lk(mu); cv_done.notify_all(); }
		}
	}
public:
	int n_threads;
	ThreadPool()
	{
		const char *e = getenv("PGX_THREADS");
		int hw = (int)std::thread::hardware_concurrency();
		n_threads = e? atoi(e) : (hw > 34? 32 : hw - 2);   // two cores stay free for the caller's other threads (CUDA's own, a sampler, the rep thread)
		if (n_threads < 1) n_threads = 1;
		for (int t = 1; t < n_threads; ++t) workers.emplace_back([this] { loop(); });
	}
	~ThreadPool()
	{
		{ std::unique_lock<std::mutex> lk(mu); stop = true; }
		cv_work.notify_all();
		for (auto &w : workers) w.join();
	}
	// run f(0..n-1) on the pool; the caller works too; returns when every item has been executed (not merely claimed)
	void parallel_for(int64_t n, const std::function<void(int64_t)> &f)
	{
		if (n <= 0) return;
		if (n == 1 || workers.empty()) { for (int64_t i = 0; i < n; ++i) f(i); return; }
		std::shared_ptr<Job> j = std::make_shared<Job>(&f, n);
		{
			std::unique_lock<std::mutex> lk(mu);
			job = j; ++epoch;
		}
		cv_work.notify_all();
		drain(*j);
		std::unique_lock<std::mutex> lk(mu);
		cv_done.wait(lk, [&] { return j->done.load() >= j->n; });
		if (job == j) job.reset();
	}
	static ThreadPool &get() { static ThreadPool p; return p; }
};

// ring of pinned staging buffers; one per process, allocated on first use
struct PinnedRing {
	static constexpr int NB = 3;
	static constexpr size_t CAP = (size_t)48 << 20;
	char *buf[NB] = {nullptr, nullptr, nullptr};
	cudaEvent_t ev[NB];
	bool ready = false;
	void init()
	{
		if (ready) return;
		for (int b = 0; b < NB; ++b) { PGX_CHECK(cudaMallocHost((void**)&buf[b], CAP)); PGX_CHECK(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming)); }
		ready = true;
	}
	static PinnedRing &get() { static PinnedRing r; r.init(); return r; }
};

// dev[0..total) <- bytes produced by fill(byte_offset, n_bytes, dst) in chunks that are multiples of `unit`
template<class Fill> static void staged_h2d(void *dev, size_t total, size_t unit, Fill fill, cudaStream_t st)
{
	PinnedRing &r = PinnedRing::get();
	const size_t chunk = PinnedRing::CAP / unit * unit;
	int c = 0;
	for (size_t off = 0; off < total; off += chunk, ++c) {
		const int b = c % PinnedRing::NB;
		const size_t len = total - off < chunk? total - off : chunk;
		if (c >= PinnedRing::NB) PGX_CHECK(cudaEventSynchronize(r.ev[b]));
		fill(off, len, r.buf[b]);
		PGX_CHECK(cudaMemcpyAsync((char*)dev + off, r.buf[b], len, cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaEventRecord(r.ev[b], st));
	}
	for (int b = 0; b < PinnedRing::NB && b < c; ++b) PGX_CHECK(cudaEventSynchronize(r.ev[b]));
}

// consume(byte_offset, n_bytes, src) sees dev[0..total) chunk by chunk, copies running ahead of the consumer
template<class Consume> static void staged_d2h(const void *dev, size_t total, size_t unit, Consume consume, cudaStream_t st)
{
	PinnedRing &r = PinnedRing::get();
	const size_t chunk = PinnedRing::CAP / unit * unit;
	const int n_chunks = (int)((total + chunk - 1) / chunk);
	auto issue = [&](int c) {
		const size_t off = (size_t)c * chunk, len = total - off < chunk? total - off : chunk;
		const int b = c % PinnedRing::NB;
		PGX_CHECK(cudaMemcpyAsync(r.buf[b], (const char*)dev + off, len, cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaEventRecord(r.ev[b], st));
	};
	for (int c = 0; c < n_chunks && c < PinnedRing::NB - 1; ++c) issue(c);
	for (int c = 0; c < n_chunks; ++c) {
		if (c + PinnedRing::NB - 1 < n_chunks) issue(c + PinnedRing::NB - 1);   // its buffer was consumed in iteration c-1
		const int b = c % PinnedRing::NB;
		PGX_CHECK(cudaEventSynchronize(r.ev[b]));
		const size_t off = (size_t)c * chunk, len = total - off < chunk? total - off : chunk;
		consume(off, len, r.buf[b]);
	}
}

} // namespace pgx
