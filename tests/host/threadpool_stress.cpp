// Stress test of pgx::ThreadPool (pgx_hostpipe.hpp): many tiny parallel_for calls back to back -- the pattern of the tail chunks of
// staged_h2d/staged_d2h and of fix_tied_ordinals -- each with a function object that dies when its call returns.  Built with
// -fsanitize=thread by tests/test_host_threadpool.py; exits 0 when every item of every call ran exactly once.
#include <cstdio>
#include <vector>
#include <atomic>
#include "pgx_hostpipe.hpp"

int main()
{
	pgx::ThreadPool &pool = pgx::ThreadPool::get();
	long bad = 0;
	for (int round = 0; round < 20000; ++round) {
		const int n = 2 + round % 7;
		std::vector<std::atomic<int>> hit(n);
		for (auto &h : hit) h.store(0);
		{
			std::vector<int> payload(n, round);          // captured by reference: must outlive every worker's use of the lambda
			pool.parallel_for(n, [&](int64_t i) { if (payload[i] == round) hit[i].fetch_add(1); });
		}
		for (int i = 0; i < n; ++i) if (hit[i].load() != 1) ++bad;
	}
	// a long job right after a short one, and a short one right after a long one
	std::atomic<long> sum{0};
	pool.parallel_for(3, [&](int64_t i) { sum += i; });
	pool.parallel_for(100000, [&](int64_t i) { sum += i & 1; });
	pool.parallel_for(2, [&](int64_t i) { sum += 10 * (i + 1); });
	if (sum.load() != 3 + 50000 + 30) ++bad;
	printf("threads=%d bad=%ld\n", pool.n_threads, bad);
	return bad != 0;
}
