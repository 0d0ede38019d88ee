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
public:
	static void run(pg128_t *a, int64_t n)
	{
		if (n <= 64) { insertion(a, n); return; }
		std::vector<Range> todo;
		todo.push_back(Range{0, n, 56});
		int64_t nxt[256], end[256], beg[256];
		while (!todo.empty()) {
			Range r = todo.back(); todo.pop_back();
			pg128_t *p = a + r.beg;
			const int sh = r.shift;
			std::fill(end, end + 256, 0);
			for (int64_t i = 0; i < r.n; ++i) ++end[p[i].x >> sh & 255];
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
			if (sh == 0) continue;
			const int sub = sh > 8? sh - 8 : 0;
			for (int k = 0; k < 256; ++k) {
				int64_t m = end[k] - beg[k];
				if (m > 64) todo.push_back(Range{r.beg + beg[k], m, sub});
				else if (m > 1) insertion(p + beg[k], m);
			}
		}
	}
};

// hit.c:205-217: per-protein count / mean adjusted score, then the representative protein of every gene
struct RepResult { std::vector<int32_t> prot_n, prot_avg, prot_rep, gene_rep_pid; };
static inline void choose_representatives(const unsigned long long *z, const int32_t *prot_gid, int32_t n_prot, int32_t n_gene, RepResult &r)
{
	r.prot_n.assign(n_prot, 0); r.prot_avg.assign(n_prot, 0); r.prot_rep.assign(n_prot, 0); r.gene_rep_pid.assign(n_gene, -1);
	std::vector<pg128_t> s(n_prot);
	for (int32_t i = 0; i < n_prot; ++i) {
		s[i].x = z[i]; s[i].y = (uint64_t)i;
		r.prot_n[i] = (int32_t)(uint32_t)z[i];
		r.prot_avg[i] = r.prot_n[i]? (int32_t)((double)(z[i] >> 32) / r.prot_n[i] + .499) : 0;
	}
	FlagSortReplay::run(s.data(), n_prot);
	for (int32_t i = n_prot - 1; i >= 0; --i) {
		int32_t pid = (int32_t)s[i].y, gid = prot_gid[pid];
		if (r.gene_rep_pid[gid] < 0) r.gene_rep_pid[gid] = pid, r.prot_rep[pid] = 1;
	}
}

// vertex.c:52-98.  aux[g*n_genome+j] (gene-major): -2 = g has an unshadowed rank-0 hit in genome j, >=0 = the gene shadowing
// its only rank-0 hit there, -1 = absent.  Returns the accepted genes as segments sorted by gene id.
// The reference recounts, for every gene, in how many of its dominant genomes it has already been claimed by an accepted
// gene (vertex.c:63-65: O(n_gene * n_genome) with a stride of n_gene words).  Here that count is kept incrementally -- it only
// changes when an accepted gene claims a (genome, gene) cell for the first time -- so the pass touches only the rows of
// accepted genes.  Same decisions, same order.
struct SegHost { int32_t gid, n_dom, n_sub; };
static inline void elect_vertices(const int32_t *aux, int32_t n_genome, int32_t n_gene, const int32_t *gene_rep_pid, const int32_t *prot_avg,
	const uint8_t *gene_pref, const uint8_t *gene_incl, double min_vertex_ratio, bool dump, const pg_gene_t *names, std::vector<SegHost> &out)
{
	struct timespec ts0, ts1, ts2, ts3;
	clock_gettime(CLOCK_MONOTONIC, &ts0);
	std::vector<pg128_t> cnt(n_gene);
	std::vector<uint8_t> claimed((size_t)n_genome * n_gene, 0);
	std::vector<int32_t> n_dom_of(n_gene, 0), n_claimed(n_gene, 0);
	const int64_t GB = 512;
	ThreadPool::get().parallel_for((n_gene + GB - 1) / GB, [&](int64_t blk) {                 // per-gene counts: independent rows
		for (int32_t g = (int32_t)(blk * GB); g < n_gene && g < (int32_t)((blk + 1) * GB); ++g) {
			const int32_t *row = aux + (size_t)g * n_genome;
			uint64_t x = (uint64_t)(int64_t)prot_avg[gene_rep_pid[g]], y = (uint64_t)g;      // vertex.c:19-20
			int32_t nd = 0;
			for (int32_t j = 0; j < n_genome; ++j) {
				if (row[j] == -2) x += 1ull << 32, ++nd;                                      // vertex.c:49
				else if (row[j] >= 0) y += 1ull << 32;                                        // vertex.c:50
			}
			if (gene_pref[g]) x |= 1ull << 63;                                                // vertex.c:55-57
			cnt[g].x = x; cnt[g].y = y; n_dom_of[g] = nd;
		}
	});
	clock_gettime(CLOCK_MONOTONIC, &ts1);
	FlagSortReplay::run(cnt.data(), n_gene);                                                  // vertex.c:60
	clock_gettime(CLOCK_MONOTONIC, &ts2);
	out.clear();
	for (int32_t i = n_gene - 1; i >= 0; --i) {
		int32_t n_dom = (int32_t)(cnt[i].x << 1 >> 33), n_sub = (int32_t)(cnt[i].y >> 32), gid = (int32_t)cnt[i].y;
		int32_t x = n_dom_of[gid], y = n_claimed[gid];                                       // vertex.c:63-65
		if (dump && names) printf("g\t%s\t%d\t%d\t%d\t%d\t%c\t%c\n", names[gid].name, (int32_t)cnt[i].x, x, y, n_sub, "NY"[gene_incl[gid]], "NY"[gene_pref[gid]]);   // vertex.c:66-67
		if (gene_incl[gid] || (n_dom >= n_genome * min_vertex_ratio && y < x)) {             // vertex.c:69
			out.push_back(SegHost{gid, n_dom, n_sub});
			if (x > 0 && n_sub > 0) {                                                         // vertex.c:74-77 (n_sub == 0: no cell of this row names a shadowing gene)
				const int32_t *row = aux + (size_t)gid * n_genome;
				for (int32_t j = 0; j < n_genome; ++j) {
					int32_t t = row[j];
					if (t < 0) continue;
					uint8_t &c = claimed[(size_t)t * n_genome + j];
					if (!c) { c = 1; if (aux[(size_t)t * n_genome + j] == -2) ++n_claimed[t]; }
				}
			}
		}
	}
	std::sort(out.begin(), out.end(), [](const SegHost &a, const SegHost &b) { return a.gid < b.gid; });   // vertex.c:85-94
	clock_gettime(CLOCK_MONOTONIC, &ts3);
	if (getenv("PGX_TIMING")) {
		auto ms = [](const timespec &a, const timespec &b) { return (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6; };
		fprintf(stderr, "[pgx timing] election: counts %.2f ms, order replay %.2f ms, greedy+sort %.2f ms\n", ms(ts0, ts1), ms(ts1, ts2), ms(ts2, ts3));
	}
}

} // namespace pgx
