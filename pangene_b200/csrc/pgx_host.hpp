// pgx_host.hpp -- the two small SEQUENTIAL pieces of the path that stay on the host (replicated on every rank):
//   * the exact permutation of the reference's in-place MSD byte radix sort (ksort.h:52-87), needed because
//     ties between equal keys decide gene.rep_pid (hit.c:209-217) and the order of the greedy vertex
//     election (vertex.c:54-60) -- SURVEY.md H1 b,c;
//   * the greedy, order-dependent vertex election itself (vertex.c:60-80): O(n_gene * n_genome) with a
//     loop-carried dependence through the "claimed" marks.
// Inputs are a few hundred KB copied back from the device; everything data-parallel is in pgx_kernels.cuh.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <thread>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <time.h>
#include "pangene_abi.h"
#include "pgx_hostpipe.hpp"

namespace pgx {

// Reproduce the permutation of an American-flag sort (8-bit digits from the top byte down; buckets of <= 64
// records finished by a stable insertion sort; ranges of <= 64 records never enter the digit pass at all).
class FlagSortReplay {
	static void insertion(pg128_t *a, int64_t n)
	{
		for (int64_t i = 1; i < n; ++i) {
			if (!(a[i].x < a[i - 1].x)) continue;
			pg128_t t = a[i];
			int64_t j = i;
			while (j > 0 && t.x < a[j - 1].x) { a[j] = a[j - 1]; --j; }
			a[j] = t;
		}
	}
	struct Range { int64_t beg, n; int shift; };
	// one digit pass over a range (ksort.h:56-78); the sub-ranges that need another pass are appended to `todo`, small buckets are finished here
	static void pass(pg128_t *a, const Range &r, std::vector<Range> &todo)
	{
		int64_t nxt[256], end[256], beg[256];
		pg128_t *p = a + r.beg;
		const int sh = r.shift;
		std::fill(end, end + 256, 0);
		for (int64_t i = 0; i < r.n; ++i) ++end[p[i].x >> sh & 255];
		{   // every record in one bucket (a byte all keys share: the high bytes of small keys): the permutation below is the identity
			bool one = false;
			for (int k = 0; k < 256 && !one; ++k) one = end[k] == r.n;
			if (one) { if (sh > 0) todo.push_back(Range{r.beg, r.n, sh > 8? sh - 8 : 0}); return; }
		}
		int64_t acc = 0;
		for (int k = 0; k < 256; ++k) { beg[k] = nxt[k] = acc; acc += end[k]; end[k] = acc; }
		// permute in place: walk the buckets in order; whatever sits at the write cursor of bucket k and does not
		// belong there is carried from bucket to bucket, each time swapped against that bucket's cursor, until a
		// record of bucket k turns up.  Records already at home are stepped over one by one.
		for (int k = 0; k < 256;) {
			if (nxt[k] == end[k]) { ++k; continue; }
			int d = (int)(p[nxt[k]].x >> sh & 255);
			if (d == k) { ++nxt[k]; continue; }
			pg128_t hold = p[nxt[k]];
			while (d != k) { std::swap(hold, p[nxt[d]]); ++nxt[d]; d = (int)(hold.x >> sh & 255); }
			p[nxt[k]++] = hold;
		}
		if (sh == 0) return;
		const int sub = sh > 8? sh - 8 : 0;
		for (int k = 0; k < 256; ++k) {
			int64_t m = end[k] - beg[k];
			if (m > 64) todo.push_back(Range{r.beg + beg[k], m, sub});
			else if (m > 1) insertion(p + beg[k], m);
		}
	}
	static void finish(pg128_t *a, Range r0)
	{
		std::vector<Range> todo;
		todo.push_back(r0);
		while (!todo.empty()) { Range r = todo.back(); todo.pop_back(); pass(a, r, todo); }
	}
public:
	// Reproduce the permutation of an American-flag sort (8-bit digits from the top byte down; buckets of <= 64 records finished by a
	// stable insertion sort; ranges of <= 64 records never enter the digit pass at all).  Ranges are disjoint and their order of
	// processing does not matter, so once the first real split has produced many of them they are finished on the host thread pool.
	static void run(pg128_t *a, int64_t n)
	{
		if (n <= 64) { insertion(a, n); return; }
		std::vector<Range> todo;
		todo.push_back(Range{0, n, 56});
		while (todo.size() == 1) { Range r = todo.back(); todo.pop_back(); pass(a, r, todo); }   // identity levels, then the first split
		if (todo.empty()) return;
		// Ranges are independent.  Small inputs stay on the calling thread; large ones (the 80 k proteins of hit.c:209) are split into a few
		// balanced groups on FRESH threads -- not the sleeping pool: measured on the 16-core box, waking the pool for < 1 ms of work stalls the
		// caller for a whole scheduler period (~6 ms) in one call out of three, a new thread is placed on an idle core at once.
		if (n < (1 << 15)) { for (const Range &r : todo) finish(a, r); return; }
		const int n_grp = 6;
		std::sort(todo.begin(), todo.end(), [](const Range &x, const Range &y) { return x.n > y.n; });
		std::vector<std::vector<Range>> grp(n_grp);
		int64_t load[n_grp] = {0};
		for (const Range &r : todo) { int g = 0; for (int k = 1; k < n_grp; ++k) if (load[k] < load[g]) g = k; grp[g].push_back(r); load[g] += r.n; }
		std::vector<std::thread> th;
		for (int g = 1; g < n_grp; ++g) th.emplace_back([&grp, a, g]() { for (const Range &r : grp[g]) finish(a, r); });
		for (const Range &r : grp[0]) finish(a, r);
		for (auto &t : th) t.join();
	}
};

// hit.c:205-217: per-protein count / mean adjusted score, then the representative protein of every gene
struct RepResult { std::vector<int32_t> prot_n, prot_avg, prot_rep, gene_rep_pid; bool replayed = false; };
static inline void choose_representatives(const unsigned long long *z, const int32_t *prot_gid, int32_t n_prot, int32_t n_gene, RepResult &r, bool force_replay = false)
{
	r.prot_n.assign(n_prot, 0); r.prot_avg.assign(n_prot, 0); r.prot_rep.assign(n_prot, 0); r.gene_rep_pid.assign(n_gene, -1);
	for (int32_t i = 0; i < n_prot; ++i) {
		r.prot_n[i] = (int32_t)(uint32_t)z[i];
		r.prot_avg[i] = r.prot_n[i]? (int32_t)((double)(z[i] >> 32) / r.prot_n[i] + .499) : 0;
	}
	// The reference sorts the proteins by z (unstable radix sort) and walks the result from the end: the representative of a gene is its
	// protein with the LARGEST z, and the sort's tie order only matters where two proteins of one gene share that largest value (hit.c:209-217).
	// One pass finds the maxima and whether any gene has such a twin; only then is the exact permutation replayed (2-3 ms for 80 k proteins,
	// replicated on every rank: the largest serial piece of a sharded run before this shortcut).
	bool tie = force_replay;
	if (!tie) {
		std::vector<uint8_t> twin(n_gene, 0);
		for (int32_t i = 0; i < n_prot; ++i) {
			const int32_t g = prot_gid[i], b = r.gene_rep_pid[g];
			if (b < 0 || z[i] > z[b]) { r.gene_rep_pid[g] = i; twin[g] = 0; }
			else if (z[i] == z[b]) twin[g] = 1;
		}
		for (int32_t g = 0; g < n_gene && !tie; ++g) tie = twin[g] != 0;
	}
	if (tie) {
		std::fill(r.gene_rep_pid.begin(), r.gene_rep_pid.end(), -1);
		std::vector<pg128_t> s(n_prot);
		for (int32_t i = 0; i < n_prot; ++i) { s[i].x = z[i]; s[i].y = (uint64_t)i; }
		FlagSortReplay::run(s.data(), n_prot);
		for (int32_t i = n_prot - 1; i >= 0; --i) {
			int32_t pid = (int32_t)s[i].y, gid = prot_gid[pid];
			if (r.gene_rep_pid[gid] < 0) r.gene_rep_pid[gid] = pid;
		}
	}
	for (int32_t g = 0; g < n_gene; ++g) if (r.gene_rep_pid[g] >= 0) r.prot_rep[r.gene_rep_pid[g]] = 1;
	r.replayed = tie;
}

struct SegHost { int32_t gid, n_dom, n_sub; };

// vertex.c:52-98, the greedy election, from the compact form the device hands over (k_vtx_cells): per-gene counts and the shadowed cells only.  The reference
// walks a whole row of the [gene][genome] table for every accepted gene (vertex.c:74-77); the only cells that matter there are the
// shadowed ones, and the only thing read of the claimed gene is whether it is dominant in that genome -- both are in the cell record.
// Same decisions in the same order (the order is the exact replay of the reference's radix sort).
struct VtxCellHost { int32_t g, j, t, t_dom; };
static inline void elect_vertices_cells(const int32_t *cnt2, const VtxCellHost *cells, size_t n_cells, int32_t n_genome, int32_t n_gene, const int32_t *gene_rep_pid, const int32_t *prot_avg,
	const uint8_t *gene_pref, const uint8_t *gene_incl, double min_vertex_ratio, bool dump, const pg_gene_t *names, std::vector<SegHost> &out)
{
	struct timespec ts0, ts1, ts2, ts3;
	clock_gettime(CLOCK_MONOTONIC, &ts0);
	std::vector<pg128_t> cnt(n_gene);
	for (int32_t g = 0; g < n_gene; ++g) {
		uint64_t x = (uint64_t)(int64_t)prot_avg[gene_rep_pid[g]] + ((uint64_t)(uint32_t)cnt2[2 * g] << 32), y = (uint64_t)g + ((uint64_t)(uint32_t)cnt2[2 * g + 1] << 32);   // vertex.c:19-20,49-50
		if (gene_pref[g]) x |= 1ull << 63;                                                // vertex.c:55-57
		cnt[g].x = x; cnt[g].y = y;
	}
	// cells grouped by their gene (counting sort; the order inside a group is irrelevant: every cell of a row names a different genome)
	std::vector<uint32_t> start(n_gene + 1, 0);
	for (size_t i = 0; i < n_cells; ++i) ++start[cells[i].g + 1];
	for (int32_t g = 0; g < n_gene; ++g) start[g + 1] += start[g];
	std::vector<uint32_t> fill(start.begin(), start.end() - 1), by_gene(n_cells);
	for (size_t i = 0; i < n_cells; ++i) by_gene[fill[cells[i].g]++] = (uint32_t)i;
	clock_gettime(CLOCK_MONOTONIC, &ts1);
	FlagSortReplay::run(cnt.data(), n_gene);                                                  // vertex.c:60
	clock_gettime(CLOCK_MONOTONIC, &ts2);
	std::vector<uint64_t> claimed(((size_t)n_genome * n_gene + 63) / 64, 0);
	std::vector<int32_t> n_claimed(n_gene, 0);
	std::vector<uint8_t> accepted(n_gene, 0);
	std::vector<int32_t> n_dom_out(n_gene, 0);                                                // as the reference reads it back from the sort key (a negative mean score leaks into the field)
	for (int32_t i = n_gene - 1; i >= 0; --i) {
		const int32_t n_dom = (int32_t)(cnt[i].x << 1 >> 33), n_sub = (int32_t)(cnt[i].y >> 32), gid = (int32_t)cnt[i].y;
		const int32_t x = cnt2[2 * gid], y = n_claimed[gid];                                  // vertex.c:63-65
		if (dump && names) printf("g\t%s\t%d\t%d\t%d\t%d\t%c\t%c\n", names[gid].name, (int32_t)cnt[i].x, x, y, n_sub, "NY"[gene_incl[gid]], "NY"[gene_pref[gid]]);   // vertex.c:66-67
		if (gene_incl[gid] || (n_dom >= n_genome * min_vertex_ratio && y < x)) {             // vertex.c:69
			accepted[gid] = 1; n_dom_out[gid] = n_dom;
			if (x > 0 && n_sub > 0)                                                           // vertex.c:74-77
				for (uint32_t q = start[gid]; q < start[gid + 1]; ++q) {
					const VtxCellHost &c = cells[by_gene[q]];
					const size_t bit = (size_t)c.t * n_genome + c.j;
					if (!(claimed[bit >> 6] >> (bit & 63) & 1ull)) { claimed[bit >> 6] |= 1ull << (bit & 63); if (c.t_dom) ++n_claimed[c.t]; }
				}
		}
	}
	out.clear();
	for (int32_t g = 0; g < n_gene; ++g) if (accepted[g]) out.push_back(SegHost{g, n_dom_out[g], cnt2[2 * g + 1]});   // vertex.c:85-94: sorted by gene id
	clock_gettime(CLOCK_MONOTONIC, &ts3);
	if (getenv("PGX_TIMING")) {
		auto ms = [](const timespec &a, const timespec &b) { return (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6; };
		fprintf(stderr, "[pgx timing] election: keys + cell index %.2f ms, order replay %.2f ms, greedy + segments %.2f ms\n", ms(ts0, ts1), ms(ts1, ts2), ms(ts2, ts3));
	}
}

} // namespace pgx
