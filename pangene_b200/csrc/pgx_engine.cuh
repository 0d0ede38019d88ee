// pgx_engine.cuh -- host orchestration of the device-resident pipeline (one Session = one pg_data_t on one GPU).
//
// Layout in HBM: raw pg_hit_t records as uploaded (kept for the final materialisation), one
// packed-int32 column per field in (genome, contig, cs) order for ALL hits ("full"), and the same columns for
// the hits that survive the load-time filters ("active", typically 14-60 % of the input), plus exon lists
// re-laid contiguously in hit order for both.  Hits are physically sorted ONCE; the (contig, cm) order the
// reference re-sorts into 33 times is an index permutation of the active set.
#pragma once
#include <vector>
#include <cstring>
#include <cmath>
#include <cstddef>
#include <time.h>
#include <sys/mman.h>
#include <malloc.h>
#include <mutex>
#include "pgx.h"
#include "pgx_util.cuh"
#include "pgx_scan.cuh"
#include "pgx_sort.cuh"
#include "pgx_kernels.cuh"
#include "pgx_round.cuh"
#include "pgx_comm.cuh"
#include "pgx_host.hpp"
#include "pgx_hostpipe.hpp"
#include <atomic>
#include <algorithm>
#include <thread>

namespace pgx {

struct Stats {
	double sweep_ms = 0, sweep_bytes = 0, run_ms[3] = {0, 0, 0};
	int64_t n_hit = 0, n_active = 0, tie[4] = {0, 0, 0, 0}, h2d = 0, d2h = 0;
	int32_t n_seg = 0, n_arc = 0, fast_rounds = 0, redo = 0;
};
extern Stats g_stats;
extern int g_sweep_stat;

constexpr int TB = 256;   // default block size for the streaming kernels
constexpr int SWEEP_VARIANT_DEFAULT = 1;   // k_sweep_tiled variant (launch_sweep; PGX_SWEEP_VARIANT overrides)

// fresh hit arrays for the caller (it will free() them): 2 MB-aligned and advised huge so that the ~430 k first-touch page
// faults of a 1.8 GB result become ~900 where transparent huge pages are available (no effect otherwise)
static inline pg_hit_t *alloc_hit_array(size_t n)
{
	const size_t bytes = sizeof(pg_hit_t) * (n > 0? n : 1), HP = (size_t)2 << 20;
	if (bytes >= 4 * HP) {
		void *p = nullptr;
		const size_t sz = (bytes + HP - 1) / HP * HP;
		if (posix_memalign(&p, HP, sz) == 0) { madvise(p, sz, MADV_HUGEPAGE); return (pg_hit_t*)p; }
	}
	return (pg_hit_t*)malloc(bytes);
}

// Hit arrays that went out of use on the host side of the boundary (the caller's arrays in parsing order, once their permuted copies are
// installed) are kept here instead of being free()d, and handed out again as the destination of a later genome / a later call: a recycled
// array is already mapped, so the ~430 k first-touch page faults + the munmap of a 1.8 GB result (a third of the download time) go away after
// the first few genomes.  Everything in the pool is plain malloc'd memory, as everything the caller gets (hit.c:62-63 free()s it).
struct SpareHitArrays {
	struct Item { pg_hit_t *p; size_t cap; };
	std::vector<Item> items;
	std::mutex mu;
	static constexpr size_t MAX_ITEMS = 24, MAX_BYTES = (size_t)1 << 30;
	size_t bytes = 0;
	// an array for n records: the smallest pooled one that fits (and is not absurdly larger), else a fresh one; *cap = its capacity in records
	pg_hit_t *take(size_t n, size_t *cap)
	{
		{
			std::lock_guard<std::mutex> lk(mu);
			int best = -1;
			for (int i = 0; i < (int)items.size(); ++i)
				if (items[i].cap >= n && items[i].cap <= n + n / 2 + 4096 && (best < 0 || items[i].cap < items[best].cap)) best = i;
			if (best >= 0) { Item it = items[best]; items.erase(items.begin() + best); bytes -= it.cap * sizeof(pg_hit_t); *cap = it.cap; return it.p; }
		}
		pg_hit_t *p = alloc_hit_array(n);
		*cap = malloc_usable_size(p) / sizeof(pg_hit_t);
		return p;
	}
	void give(pg_hit_t *p)   // (capacity from the allocator, not from the caller's m_hit field)
	{
		if (!p) return;
		const size_t cap = malloc_usable_size(p) / sizeof(pg_hit_t);
		{
			std::lock_guard<std::mutex> lk(mu);
			if (cap > 0 && items.size() < MAX_ITEMS && bytes + cap * sizeof(pg_hit_t) <= MAX_BYTES) { items.push_back(Item{p, cap}); bytes += cap * sizeof(pg_hit_t); return; }
		}
		free(p);
	}
	void trim() { std::lock_guard<std::mutex> lk(mu); for (auto &it : items) free(it.p); items.clear(); bytes = 0; }
	static SpareHitArrays &get() { static SpareHitArrays s; return s; }
};

static inline double wall_s() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }

// device-clock breakdown of a run by phase (PGX_TIMING): events at phase boundaries, summed per label at the end; works at any rank count
struct PhaseTimer {
	bool on = false;
	std::vector<cudaEvent_t> ev; std::vector<const char*> label; size_t used = 0;
	void begin(cudaStream_t st) { on = getenv("PGX_TIMING") != nullptr; used = 0; label.clear(); if (on) mark("(start)", st); }
	void mark(const char *what, cudaStream_t st)
	{
		if (!on) return;
		if (used == ev.size()) { cudaEvent_t e; PGX_CHECK(cudaEventCreate(&e)); ev.push_back(e); }
		PGX_CHECK(cudaEventRecord(ev[used++], st)); label.push_back(what);
	}
	void report(const char *title, int rank, cudaStream_t st)
	{
		if (!on || used < 2) return;
		PGX_CHECK(cudaStreamSynchronize(st));
		std::vector<std::pair<const char*, double>> acc;
		double tot = 0;
		for (size_t i = 1; i < used; ++i) {
			float ms = 0; PGX_CHECK(cudaEventElapsedTime(&ms, ev[i - 1], ev[i]));
			tot += ms;
			size_t k = 0; for (; k < acc.size(); ++k) if (strcmp(acc[k].first, label[i]) == 0) break;
			if (k == acc.size()) acc.push_back({label[i], 0.0});
			acc[k].second += ms;
		}
		fprintf(stderr, "[pgx timing] rank %d %s by phase (device clock, %.2f ms; %llu cudaMalloc calls so far):", rank, title, tot, (unsigned long long)DevCache::get_cache().n_malloc);
		for (auto &a : acc) fprintf(stderr, " %s %.2f |", a.first, a.second);
		fprintf(stderr, "\n");
	}
	~PhaseTimer() { for (auto e : ev) cudaEventDestroy(e); }
};

struct Session {
	PhaseTimer pt;
	cudaStream_t st = nullptr;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	// ---- shapes
	int32_t G = 0, N = 0, n_prot = 0, n_gene = 0, n_ctg = 0, Gtot = 0;
	int64_t E = 0;
	std::vector<int32_t> h_hoff, h_coff, h_prot_gid;
	std::vector<uint32_t> h_eoffg;
	std::vector<uint8_t> h_gene_pref, h_gene_incl;
	const pg_gene_t *h_gene_names = nullptr;
	int32_t coord_bits = 32;             // bits needed for the largest cs/cm of this upload (fewer radix passes)
	std::vector<int32_t> rank_G;         // local genome count of every rank (multi-GPU)
	// ---- host-side results that go back into pg_data_t at download time
	RepResult rep;
	std::vector<int32_t> h_prot_max;
	std::vector<SegHost> h_seg0;
	bool done_post = false, done_graph = false;
	// ---- device: constants and raw input
	DBuf<int32_t> hoff, coff, ctg2g, prot_gid, prot_max, prot_rep, g2s;
	DBuf<RawHit> raw;
	std::vector<const pg_hit_t*> installed;   // hit arrays this session handed to the caller (already in (contig, cs) order)
	DBuf<uint32_t> eoffg, src, counters;
	DBuf<uint8_t> gene_pref;
	DBuf<int2> ex_raw, ex, wex;
	DBuf<unsigned long long> z, cs4, idx;
	// ---- full columns
	DBuf<int32_t> f_ctg, f_cs, f_ce, f_cm, f_pid, f_gidp, f_sadj, f_rank, f_sori, f_sdom, f_pdom0, f_pdom, f_cds, f_pmax;
	DBuf<uint32_t> f_bits, f_bits_alt, f_nex, f_eoff, f_esrc;
	DBuf<uint64_t> f_key, a_key;
	FullCols F;
	// ---- active columns
	int32_t Nw = 0;
	DBuf<int32_t> a_ctg, a_cs, a_ce, a_cm, a_pid, a_gidp, a_sadj, a_rank, a_sori, a_sdom, a_pdom0, a_pdom, a_cds, a_pmax, a_full, goff_act;
	DBuf<uint32_t> a_bits, a_bits_alt, a_nex, a_eoff, by_cm, sv;
	ActCols A;
	int32_t m_sv = 0;
	// ---- graph
	int32_t n_seg = 0, n_arc = 0, n_seg0 = 0;   // host copies; during the rounds the live values are in `dc`
	size_t arc_cap = 0;
	DBuf<DevCounts> dc;
	DBuf<SegDev> seg, seg_alt;
	DBuf<ArcDev> arc, arc_alt;
	DBuf<int32_t> seg_ng, seg_tot, arc_of_rec;
	// ---- overlap.c:108 quirk, moving part (k_first_group / prepare_first_replay / k_set_first)
	DBuf<FirstGroup> first_grp;
	DBuf<int32_t> first_flagged, first_target;
	int32_t n_first = 0, arc_call = 0;
	// ---- branch.c:18-25 ordinals of cs-tied survivors (k_cs_tie_mark / fix_tied_ordinals): the replayed sort chain of a contig, advanced round by round
	struct TieChain { int32_t lo = 0, n = 0, pairs_done = 0; std::vector<int32_t> cs, cm; std::vector<pg128_t> a; };
	std::map<int32_t, TieChain> tie_chains;
	Arena scratch;
	uint32_t *h_pin = nullptr;           // pinned word(s) for small read-backs
	unsigned long long *h_z_pin = nullptr;   // pinned copy of the per-protein sums for the host thread below
	cudaEvent_t ev_z = nullptr;
	std::thread rep_thread; bool rep_pending = false;
	// second stream: the gene-position tables of the next branch round are built beside the arc table of the current pg_gen_arc
	// (both only read the flag words the sweep has just finished with)
	cudaStream_t st2 = nullptr;
	cudaEvent_t ev_swept = nullptr, ev_pos = nullptr;
	bool pos_ready = false;
	enum { C_ERR = 0, C_TIE_CM = 1, C_TIE_CS = 2, C_TIE_DUP = 3, C_TIE_SUB = 4, C_BR_CHANGED = 5, C_FIRST_TIE = 6, C_CMTIE_ROUND = 7, C_TIE_ROUND = 8, C_LONG_RUN = 9, C_ORD_SENS = 10, C_N = 12 };

	Session() { PGX_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); for (auto &e : ev) PGX_CHECK(cudaEventCreate(&e)); PGX_CHECK(cudaEventCreateWithFlags(&ev_z, cudaEventDisableTiming)); PGX_CHECK(cudaMallocHost((void**)&h_pin, 64));
		PGX_CHECK(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking)); PGX_CHECK(cudaEventCreateWithFlags(&ev_swept, cudaEventDisableTiming)); PGX_CHECK(cudaEventCreateWithFlags(&ev_pos, cudaEventDisableTiming)); }
	~Session()
	{
		if (rep_pending) { rep_thread.join(); rep_pending = false; }
		if (st2) { cudaStreamSynchronize(st2); cudaStreamDestroy(st2); }
		if (ev_swept) cudaEventDestroy(ev_swept);
		if (ev_pos) cudaEventDestroy(ev_pos);
		if (st) cudaStreamDestroy(st);
		for (auto &e : ev) if (e) cudaEventDestroy(e);
		if (ev_z) cudaEventDestroy(ev_z);
		if (h_pin) cudaFreeHost(h_pin);
		if (h_z_pin) cudaFreeHost(h_z_pin);
	}

	uint32_t read_u32(const uint32_t *dptr)
	{
		PGX_CHECK(cudaMemcpyAsync(h_pin, dptr, 4, cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		return h_pin[0];
	}
	void check_err(const char *where)
	{
		uint32_t e = read_u32(counters.p + C_ERR);
		if (e & 1) PGX_FATAL("%s: a hit coordinate is outside [0, 2^31) or cs > ce -- this build keeps coordinates in int32", where);
		if (e & 2) PGX_FATAL("%s: contig id out of range", where);
		if (e & 4) PGX_FATAL("%s: protein id out of range", where);
		if (e & 8) PGX_FATAL("%s: hit without exons", where);
		if (e & 16) PGX_FATAL("%s: shadowed hit without a dominating protein (vertex.c:38 assertion)", where);
		if (e & 32) PGX_FATAL("%s: surviving hit without a segment (graph.c:111 assertion)", where);
	}

	// ------------------------------------------------------------------------------------------------
	// defer_exons: the exon lists (half of the bytes) are not copied here but inside the stage_load() that follows, on the second stream, while
	// the device already unpacks, flags and sorts the hits -- none of which reads an exon (pg_post_process() does this; `d` must stay valid until then)
	const pg_data_t *exons_pending = nullptr;
	bool packed_input = false;           // the hits came as device-ready records (pgx_packed.h): there is no pg_hit_t array to materialise results into
	void upload(const pg_data_t *d, bool defer_exons = false, const pgx_packed_t *pk = nullptr)
	{
		if (rep_pending) { rep_thread.join(); rep_pending = false; }
		G = d->n_genome; n_prot = d->n_prot; n_gene = d->n_gene;
		Gtot = g_comm.active()? g_comm.n_genome_total : G;
		h_hoff.assign(G + 1, 0); h_coff.assign(G + 1, 0); h_eoffg.assign(G + 1, 0);
		int64_t nh = 0, nc = 0, ne = 0;
		for (int32_t j = 0; j < G; ++j) {
			h_hoff[j] = (int32_t)nh; h_coff[j] = (int32_t)nc; h_eoffg[j] = (uint32_t)ne;
			nh += pk? pk->n_hit[j] : d->genome[j].n_hit; nc += d->genome[j].n_ctg; ne += d->genome[j].n_exon;
		}
		if (nh >= 0x7fffffffll / AOS_WORDS * 8 || ne >= 0xffffffffll || nc >= 0x7fffffffll) PGX_FATAL("input too large for 32-bit hit/exon indices (%ld hits, %ld exons)", (long)nh, (long)ne);
		h_hoff[G] = (int32_t)nh; h_coff[G] = (int32_t)nc; h_eoffg[G] = (uint32_t)ne;
		N = (int32_t)nh; n_ctg = (int32_t)nc; E = ne;
		h_prot_gid.resize(n_prot); h_gene_pref.resize(n_gene); h_gene_incl.resize(n_gene);
		for (int32_t i = 0; i < n_prot; ++i) h_prot_gid[i] = d->prot[i].gid;
		for (int32_t i = 0; i < n_gene; ++i) h_gene_pref[i] = d->gene[i].preferred, h_gene_incl[i] = d->gene[i].included;
		h_gene_names = d->gene;
		std::vector<int32_t> h_ctg2g(n_ctg > 0? n_ctg : 1);
		for (int32_t j = 0; j < G; ++j) for (int32_t c = h_coff[j]; c < h_coff[j + 1]; ++c) h_ctg2g[c] = j;

		hoff.alloc(G + 1); coff.alloc(G + 1); eoffg.alloc(G + 1); ctg2g.alloc(h_ctg2g.size()); prot_gid.alloc(n_prot); gene_pref.alloc(n_gene);
		PGX_CHECK(cudaMemcpyAsync(hoff.p, h_hoff.data(), (G + 1) * 4, cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaMemcpyAsync(coff.p, h_coff.data(), (G + 1) * 4, cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaMemcpyAsync(eoffg.p, h_eoffg.data(), (G + 1) * 4, cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaMemcpyAsync(ctg2g.p, h_ctg2g.data(), h_ctg2g.size() * 4, cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaMemcpyAsync(prot_gid.p, h_prot_gid.data(), (size_t)n_prot * 4, cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaMemcpyAsync(gene_pref.p, h_gene_pref.data(), (size_t)n_gene, cudaMemcpyHostToDevice, st));
		raw.alloc((size_t)N + 1); ex_raw.alloc((size_t)E + 16);
		packed_input = pk != nullptr;
		if (pk) { if (pk->n_genome != G) PGX_FATAL("upload: %d packed genomes for a pg_data_t of %d", pk->n_genome, G); upload_hits_packed(pk); } else upload_hits(d);
		if (defer_exons && E) exons_pending = d; else upload_exons(d, st);
		g_stats.h2d += (int64_t)N * (int64_t)sizeof(RawHit) + E * 8 + (int64_t)n_prot * 4 + n_gene;
		g_stats.n_hit = N;
		counters.alloc(C_N); counters.zero(st);
		size_t tab = (size_t)(G > 0? G : 1);
		scratch.init((size_t)(N > 0? N : 1) * 96 + tab * n_prot * 16 + tab * n_gene * 48 + ((size_t)256 << 20));
		// full columns
		size_t n1 = (size_t)N + 16;   // slack: the tiled sweep stages 16-byte-granular column windows
		f_key.alloc(n1); f_ctg.alloc(n1); f_cs.alloc(n1); f_ce.alloc(n1); f_cm.alloc(n1); f_pid.alloc(n1); f_gidp.alloc(n1); f_sadj.alloc(n1); f_rank.alloc(n1); f_sori.alloc(n1);
		f_sdom.alloc(n1); f_pdom0.alloc(n1); f_pdom.alloc(n1); f_cds.alloc(n1); f_pmax.alloc(n1); f_bits.alloc(n1); f_bits_alt.alloc(n1); f_nex.alloc(n1); f_eoff.alloc(n1 + 16); f_esrc.alloc(n1);
		src.alloc(n1); ex.alloc((size_t)E + 16);
		set_full();
		if (h_z_pin) { cudaFreeHost(h_z_pin); h_z_pin = nullptr; }
		PGX_CHECK(cudaMallocHost((void**)&h_z_pin, ((size_t)(n_prot > 0? n_prot : 1)) * 8));
		prot_max.alloc(n_prot); prot_rep.alloc(n_prot); z.alloc(n_prot); cs4.alloc((size_t)n_prot * 4); g2s.alloc(n_gene);
		prot_rep.zero(st);
		PGX_CHECK(cudaMemsetAsync(g2s.p, 0xff, (size_t)n_gene * 4, st));
		if (g_comm.active()) {   // every rank learns every rank's local genome and hit counts
			int32_t *tmp = scratch.get<int32_t>(2 * (size_t)g_comm.nranks);
			const int32_t mine[2] = {G, N};
			PGX_CHECK(cudaMemcpyAsync(tmp + 2 * g_comm.rank, mine, 8, cudaMemcpyHostToDevice, st));
			g_comm.allgather(tmp + 2 * g_comm.rank, tmp, 8, st);
			std::vector<int32_t> h2(2 * (size_t)g_comm.nranks);
			PGX_CHECK(cudaMemcpyAsync(h2.data(), tmp, 8 * (size_t)g_comm.nranks, cudaMemcpyDeviceToHost, st));
			PGX_CHECK(cudaStreamSynchronize(st));
			rank_G.resize(g_comm.nranks); rank_N.resize(g_comm.nranks);
			for (int r = 0; r < g_comm.nranks; ++r) { rank_G[r] = h2[2 * r]; rank_N[r] = h2[2 * r + 1]; }
			scratch.reset(0);
		} else { rank_G.assign(1, G); rank_N.assign(1, N); }
		PGX_CHECK(cudaStreamSynchronize(st));
		done_post = done_graph = false; installed.clear();
	}

	// host threads pack 48-byte RawHit records from the caller's pg_hit_t arrays into pinned chunks while the copy engine
	// drains the previous chunk; exon lists follow through the same ring
	void upload_hits(const pg_data_t *d)
	{
		ThreadPool &pool = ThreadPool::get();
		std::atomic<uint32_t> bad{0};
		std::atomic<int64_t> max_coord{0};
		const int64_t BLK = 1 << 14;
		if (N) staged_h2d(raw.p, (size_t)N * sizeof(RawHit), sizeof(RawHit), [&](size_t off, size_t len, char *dst) {
			const int64_t h0 = (int64_t)(off / sizeof(RawHit)), h1 = h0 + (int64_t)(len / sizeof(RawHit));
			pool.parallel_for((h1 - h0 + BLK - 1) / BLK, [&](int64_t b) {
				const int64_t a = h0 + b * BLK, e = std::min(h1, a + BLK);
				int g = (int)(std::upper_bound(h_hoff.begin(), h_hoff.end(), (int32_t)a) - h_hoff.begin()) - 1;
				RawHit *o = (RawHit*)dst + (a - h0);
				uint32_t err = 0;
				int64_t mc = 0;
				for (int64_t t = a; t < e; ++t, ++o) {
					while (t >= h_hoff[g + 1]) ++g;
					const pg_hit_t &h = d->genome[g].hit[t - h_hoff[g]];
					if (h.cs < 0 || h.ce > 0x7fffffffll || h.cm < 0 || h.cm > 0x7fffffffll || h.cs > h.ce) err = 1;
					o->pid = h.pid; o->cid = h.cid; o->rank = h.rank; o->sori = h.score_ori; o->sadj = h.score_adj; o->nexon = h.n_exon; o->offexon = h.off_exon;
					o->rev = h.rev; o->cs = (int32_t)h.cs; o->cm = (int32_t)h.cm; o->ce = (int32_t)h.ce; o->lof = h.lof;
					if (h.cm > mc) mc = h.cm;
					if (h.cs > mc) mc = h.cs;
				}
				if (err) bad.store(1);
				for (int64_t cur = max_coord.load(); mc > cur && !max_coord.compare_exchange_weak(cur, mc);) { }
			});
		}, st);
		if (bad.load()) PGX_FATAL("upload: a hit coordinate is outside [0, 2^31) or cs > ce -- this build keeps coordinates in int32");
		coord_bits = bits_for((uint64_t)max_coord.load());
	}
	// device-ready records (pgx_packed.h) go into the staging ring as they are; the one thing still computed on the way is the largest coordinate
	void upload_hits_packed(const pgx_packed_t *pk)
	{
		static_assert(sizeof(pgx_packed_hit_t) == sizeof(RawHit), "device-ready record == RawHit");
		ThreadPool &pool = ThreadPool::get();
		std::atomic<uint32_t> bad{0};
		std::atomic<int32_t> max_coord{0};
		const int64_t BLK = 1 << 14;
		if (N) staged_h2d(raw.p, (size_t)N * sizeof(RawHit), sizeof(RawHit), [&](size_t off, size_t len, char *dst) {
			const int64_t h0 = (int64_t)(off / sizeof(RawHit)), h1 = h0 + (int64_t)(len / sizeof(RawHit));
			pool.parallel_for((h1 - h0 + BLK - 1) / BLK, [&](int64_t b) {
				int64_t a = h0 + b * BLK;
				const int64_t e = std::min(h1, a + BLK);
				int g = (int)(std::upper_bound(h_hoff.begin(), h_hoff.end(), (int32_t)a) - h_hoff.begin()) - 1;
				int32_t mc = 0; uint32_t err = 0;
				while (a < e) {
					while (a >= h_hoff[g + 1]) ++g;
					const int64_t m = std::min<int64_t>(e, h_hoff[g + 1]) - a;
					const pgx_packed_hit_t *sp = pk->hit[g] + (a - h_hoff[g]);
					RawHit *o = (RawHit*)dst + (a - h0);
					memcpy(o, sp, (size_t)m * sizeof(RawHit));
					for (int64_t t = 0; t < m; ++t) {
						if (sp[t].cs < 0 || sp[t].cm < 0 || sp[t].cs > sp[t].ce) err = 1;
						if (sp[t].cm > mc) mc = sp[t].cm;
						if (sp[t].cs > mc) mc = sp[t].cs;
					}
					a += m;
				}
				if (err) bad.store(1);
				for (int32_t cur = max_coord.load(); mc > cur && !max_coord.compare_exchange_weak(cur, mc);) { }
			});
		}, st);
		if (bad.load()) PGX_FATAL("upload: a packed hit has a negative coordinate or cs > ce");
		coord_bits = bits_for((uint64_t)max_coord.load());
	}
	void upload_exons(const pg_data_t *d, cudaStream_t cs)
	{
		ThreadPool &pool = ThreadPool::get();
		if (E) staged_h2d(ex_raw.p, (size_t)E * 8, 8, [&](size_t off, size_t len, char *dst) {
			const int64_t e0 = (int64_t)(off / 8), e1 = e0 + (int64_t)(len / 8), EB = 1 << 17;
			pool.parallel_for((e1 - e0 + EB - 1) / EB, [&](int64_t b) {
				int64_t a = e0 + b * EB, e = std::min(e1, a + EB);
				int g = (int)(std::upper_bound(h_eoffg.begin(), h_eoffg.end(), (uint32_t)a) - h_eoffg.begin()) - 1;
				while (a < e) {
					while (a >= (int64_t)h_eoffg[g + 1]) ++g;
					int64_t m = std::min<int64_t>(e, h_eoffg[g + 1]) - a;
					memcpy(dst + (a - e0) * 8, d->genome[g].exon + (a - h_eoffg[g]), (size_t)m * 8);
					a += m;
				}
			});
		}, cs);
	}
	void set_full()
	{
		F.ctg = f_ctg.p; F.cs = f_cs.p; F.ce = f_ce.p; F.cm = f_cm.p; F.pid = f_pid.p; F.gidp = f_gidp.p; F.sadj = f_sadj.p; F.rank = f_rank.p; F.sori = f_sori.p;
		F.sdom = f_sdom.p; F.pdom0 = f_pdom0.p; F.pdom = f_pdom.p; F.cds = f_cds.p; F.pmax = f_pmax.p; F.bits = f_bits.p; F.nex = f_nex.p; F.eoff = f_eoff.p; F.esrc = f_esrc.p; F.key = f_key.p;
	}
	void set_active()
	{
		A.ctg = a_ctg.p; A.cs = a_cs.p; A.ce = a_ce.p; A.cm = a_cm.p; A.pid = a_pid.p; A.gidp = a_gidp.p; A.sadj = a_sadj.p; A.rank = a_rank.p; A.sori = a_sori.p;
		A.sdom = a_sdom.p; A.pdom0 = a_pdom0.p; A.pdom = a_pdom.p; A.cds = a_cds.p; A.pmax = a_pmax.p; A.full = a_full.p; A.bits = a_bits.p; A.bits_alt = a_bits_alt.p; A.nex = a_nex.p; A.eoff = a_eoff.p; A.key = a_key.p;
	}
	SweepCols sweep_full() const
	{
		SweepCols c; c.n = N; c.ctg = F.ctg; c.cs = F.cs; c.ce = F.ce; c.pid = F.pid; c.gidp = F.gidp; c.key = F.key; c.rank = F.rank; c.cds = F.cds; c.sori = F.sori; c.pmax = F.pmax;
		c.bits = F.bits; c.eoff = F.eoff; c.ex = ex.p; return c;
	}
	SweepCols sweep_active() const
	{
		SweepCols c; c.n = Nw; c.ctg = A.ctg; c.cs = A.cs; c.ce = A.ce; c.pid = A.pid; c.gidp = A.gidp; c.key = A.key; c.rank = A.rank; c.cds = A.cds; c.sori = A.sori; c.pmax = A.pmax;
		c.bits = A.bits; c.eoff = A.eoff; c.ex = wex.p; return c;
	}

	// one sweep launch: the TMA-tiled kernel; PGX_SWEEP=global selects the plain thread-per-hit kernel (debugging / A-B timing)
	template<int MODE> void launch_sweep(const SweepCols &c, const SweepOpt &so, uint32_t *bits_out, int32_t *pdom_out, int32_t *sdom_out, const uint32_t *dirty = nullptr)
	{
		static const bool plain = []() { const char *e = getenv("PGX_SWEEP"); return e && strcmp(e, "global") == 0; }();
		if (c.n <= 0) return;
		if (plain) { PGX_LAUNCH(k_sweep<MODE>, grid_for(c.n, 128), 128, 0, st, c, so, bits_out, pdom_out, sdom_out, counters.p + C_TIE_DUP); return; }
		// PGX_SWEEP_VARIANT: the A/B switch of k_sweep_tiled (pgx_kernels.cuh): 0 = the kernel as first measured in round 2, 1 (default) = 32-bit cluster
		// maxima + staged result columns, isoform sweep on the window scan, 2 = as 1 with the isoform sweep on the cluster shortcut, 3 = as 1 without
		// the cluster pass in the first sweep
		static const int variant = []() { const char *e = getenv("PGX_SWEEP_VARIANT"); const int v = e? atoi(e) : SWEEP_VARIANT_DEFAULT; return v < 0 || v > 3? SWEEP_VARIANT_DEFAULT : v; }();
		constexpr int V2 = MODE == SW_ISO? 2 : 1, V3 = MODE == SW_LOAD? 3 : 1;   // (variant 2 differs from 1 in the isoform sweep only, variant 3 in the first sweep only)
		static int attr_dev = -1;              // (function attributes are per device: set again if the process moved to another one)
		int cur_dev = 0; PGX_CHECK(cudaGetDevice(&cur_dev));
		if (attr_dev != cur_dev) {
			// five 46 KB CTAs per SM need the large shared-memory carve-out (ncu: the default left room for three)
			auto prep = [](const void *fn) {
				PGX_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWQ_SMEM));
				PGX_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
			};
			prep((const void*)k_sweep_tiled<MODE, 0>);
			if (MODE != SW_SHADOW) { prep((const void*)k_sweep_tiled<MODE, 1>); prep((const void*)k_sweep_tiled<MODE, V2>); prep((const void*)k_sweep_tiled<MODE, V3>); }
			attr_dev = cur_dev;
		}
		const int v = MODE == SW_SHADOW? 0 : variant;
		if (v == 0) PGX_LAUNCH((k_sweep_tiled<MODE, 0>), grid_for(c.n, SWQ_TILE), SWQ_TILE, SWQ_SMEM, st, c, so, bits_out, pdom_out, sdom_out, counters.p + C_TIE_DUP, dirty);
		else if (v == 1) PGX_LAUNCH((k_sweep_tiled<MODE, 1>), grid_for(c.n, SWQ_TILE), SWQ_TILE, SWQ_SMEM, st, c, so, bits_out, pdom_out, sdom_out, counters.p + C_TIE_DUP, dirty);
		else if (v == 2) PGX_LAUNCH((k_sweep_tiled<MODE, V2>), grid_for(c.n, SWQ_TILE), SWQ_TILE, SWQ_SMEM, st, c, so, bits_out, pdom_out, sdom_out, counters.p + C_TIE_DUP, dirty);
		else PGX_LAUNCH((k_sweep_tiled<MODE, V3>), grid_for(c.n, SWQ_TILE), SWQ_TILE, SWQ_SMEM, st, c, so, bits_out, pdom_out, sdom_out, counters.p + C_TIE_DUP, dirty);
	}

	// stable radix sort of n (key, iota) pairs taken from the scratch arena; returns the permutation (device pointer valid until the arena is reset)
	uint32_t *sort_perm(uint64_t *keys, int64_t n, int n_bits)
	{
		RadixSortBuffers b;
		uint32_t *vals = scratch.get<uint32_t>(n + 1);
		b.keys_alt = scratch.get<uint64_t>(n + 1); b.vals_alt = scratch.get<uint32_t>(n + 1);
		size_t he = RadixSortBuffers::hist_elems(n);
		b.hist = scratch.get<uint32_t>(he); b.scan_agg = scratch.get<uint32_t>(scan_scratch_elems((int64_t)he));
		uint64_t *kr; uint32_t *vr;
		radix_sort_pairs(keys, vals, n, n_bits, true, b, &kr, &vr, st);
		if (kr != keys && n > 0) PGX_CHECK(cudaMemcpyAsync(keys, kr, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));   // callers want sorted keys in place
		return vr;
	}
	// exclusive scan of n flags into pos[n+1]; returns the total
	uint32_t scan_flags(const uint32_t *flag, uint32_t *pos, int64_t n)
	{
		uint32_t *agg = scratch.get<uint32_t>(scan_scratch_elems(n));
		exclusive_sum_u32(flag, pos, n, agg, st);
		return read_u32(pos + n);
	}

	// ------------------------------------------------------------------------------------------------
	// stage 1a: the per-genome load-time block (read.c:243-260), all local genomes in single launches
	// ------------------------------------------------------------------------------------------------
	void stage_load(const pg_opt_t *opt)
	{
		if (N == 0) { Nw = 0; return; }
		size_t mk = scratch.mark();
		SweepOpt so; so.min_ov_ratio = opt->min_ov_ratio; so.check_strand = !!(opt->flag & PG_F_CHECK_STRAND);
		// -- raw order: unpack + pseudogene flags (hit.c:66-105)
		int32_t *r_rank = scratch.get<int32_t>(N), *r_gl = scratch.get<int32_t>(N);
		uint8_t *r_pseudo = scratch.get<uint8_t>(N);
		uint64_t *key = scratch.get<uint64_t>(N + 1);
		PGX_LAUNCH(k_unpack, grid_for(N, TB), TB, 0, st, raw.p, N, hoff.p, coff.p, G, r_rank, r_gl, key, n_prot, counters.p + C_ERR);
		{
			size_t tab = (size_t)G * n_prot, m2 = scratch.mark();
			int32_t *t_max = scratch.get<int32_t>(tab), *t_min = scratch.get<int32_t>(tab), *t_r1 = scratch.get<int32_t>(tab);
			PGX_CHECK(cudaMemsetAsync(t_max, 0, tab * 4, st));
			PGX_CHECK(cudaMemsetAsync(t_min, 0x7f, tab * 4, st));     // 0x7f7f7f7f: larger than any exon count
			PGX_LAUNCH(k_fill_i32, grid_for((int64_t)tab, TB), TB, 0, st, t_r1, (int64_t)tab, INT32_MAX);
			PGX_LAUNCH(k_fp_minmax, grid_for(N, TB), TB, 0, st, raw.p, r_gl, N, n_prot, t_max, t_min);
			PGX_LAUNCH(k_fp_flag, grid_for(N, TB), TB, 0, st, raw.p, r_rank, r_gl, N, n_prot, t_max, t_min, t_r1, r_pseudo);
			PGX_LAUNCH(k_fp_rank, grid_for(N, TB), TB, 0, st, raw.p, r_gl, N, n_prot, t_max, t_min, t_r1, r_rank);
			scratch.reset(m2);
		}
		check_err("upload");
		// -- the one physical sort: (genome:contig, cs), stable => ties keep file order
		{
			size_t m2 = scratch.mark();
			PGX_LAUNCH(k_pack_sort_key, grid_for(N, TB), TB, 0, st, key, N, coord_bits);
			uint32_t *perm = sort_perm(key, N, coord_bits + bits_for((uint64_t)(n_ctg > 0? n_ctg - 1 : 0)));
			PGX_CHECK(cudaMemcpyAsync(src.p, perm, (size_t)N * 4, cudaMemcpyDeviceToDevice, st));
			scratch.reset(m2);
		}
		PGX_LAUNCH(k_gather, grid_for(N, TB), TB, 0, st, raw.p, src.p, N, hoff.p, coff.p, eoffg.p, r_rank, r_pseudo, r_gl, prot_gid.p, gene_pref.p, F);
		first_grp.ensure((size_t)G + 1);
		PGX_LAUNCH(k_first_group, grid_for(G, TB), TB, 0, st, hoff.p, G, F.ctg, F.cs, first_grp.p, counters.p + C_FIRST_TIE);
		{
			uint32_t *agg = scratch.get<uint32_t>(scan_scratch_elems(N));
			exclusive_sum_u32(F.nex, F.eoff, N, agg, st);
		}
		if (exons_pending) {   // the deferred half of the upload: host threads stage the exon lists while the launches above run
			upload_exons(exons_pending, st2);
			exons_pending = nullptr;
			PGX_CHECK(cudaEventRecord(ev_pos, st2));
			PGX_CHECK(cudaStreamWaitEvent(st, ev_pos, 0));
		}
		PGX_LAUNCH(k_gather_exons, grid_for(N, 128), 128, 0, st, ex_raw.p, F.esrc, F.eoff, N, ex.p, F.cds);
		{
			SegVal *agg = scratch.get<SegVal>((size_t)(N / SCAN_TILE + 2));
			segmented_running_max(F.ctg, F.ce, F.pmax, N, agg, st);
		}
		if (g_sweep_stat) {   // algorithmic bytes of the launch below (pgx_set_sweep_stat: bench.py turns it on for one warm-up pass)
			unsigned long long *tot = scratch.get<unsigned long long>(1);
			PGX_CHECK(cudaMemsetAsync(tot, 0, 8, st));
			PGX_LAUNCH(k_sweep_bytes, grid_for(N, TB), TB, 0, st, F.bits, F.eoff, N, tot);
			unsigned long long h = 0;
			PGX_CHECK(cudaMemcpyAsync(&h, tot, 8, cudaMemcpyDeviceToHost, st));
			PGX_CHECK(cudaStreamSynchronize(st));
			g_stats.sweep_bytes = (double)h;
		}
		// -- first sweep: pg_shadow(cal_dom_sc=1) + read.c:249-253; this is the launch the roofline figure is quoted on
		PGX_CHECK(cudaEventRecord(ev[0], st));
		launch_sweep<SW_LOAD>(sweep_full(), so, f_bits_alt.p, F.pdom0, F.sdom);
		PGX_CHECK(cudaEventRecord(ev[1], st));
		// (SW_LOAD leaves the flag words unchanged: keep reading F.bits)
		// -- second sweep: pg_flt_ov_isoform
		launch_sweep<SW_ISO>(sweep_full(), so, f_bits_alt.p, nullptr, nullptr);
		f_bits.swap(f_bits_alt); set_full();
		// -- chain shadow + sub-optimal isoform (hit.c:107-146)
		{
			size_t tab = (size_t)G * n_prot, m2 = scratch.mark();
			uint8_t *t_kept = scratch.get<uint8_t>(tab);
			PGX_CHECK(cudaMemsetAsync(t_kept, 0, tab, st));
			PGX_LAUNCH(k_chain_mark, grid_for(N, TB), TB, 0, st, F.ctg, ctg2g.p, F.pid, F.bits, N, n_prot, t_kept);
			PGX_LAUNCH(k_chain_apply, grid_for(N, TB), TB, 0, st, F.ctg, ctg2g.p, F.pdom0, F.bits, N, n_prot, t_kept);
			scratch.reset(m2);
			size_t tg = (size_t)G * n_gene;
			unsigned long long *t_best = scratch.get<unsigned long long>(tg);
			PGX_CHECK(cudaMemsetAsync(t_best, 0, tg * 8, st));
			PGX_LAUNCH(k_subopt_best, grid_for(N, TB), TB, 0, st, F.ctg, ctg2g.p, F.gidp, F.sadj, F.rank, F.bits, N, n_gene, t_best);
			PGX_LAUNCH(k_subopt_apply, grid_for(N, TB), TB, 0, st, F.ctg, ctg2g.p, F.gidp, F.pid, F.bits, N, n_gene, t_best);
			scratch.reset(m2);
		}
		scratch.reset(mk);
	}

	// ------------------------------------------------------------------------------------------------
	// stage 1b: pg_post_process (graph.c:7-32) -- builds the active set on the way
	// ------------------------------------------------------------------------------------------------
	void stage_post(const pg_opt_t *opt)
	{
		SweepOpt so; so.min_ov_ratio = opt->min_ov_ratio; so.check_strand = !!(opt->flag & PG_F_CHECK_STRAND);
		// pg_cap_score_dom (hit.c:227-247): over ALL hits, filtered ones included
		prot_max.zero(st);
		if (N) PGX_LAUNCH(k_cap_max, grid_for(N, TB), TB, 0, st, F.pid, F.sori, N, prot_max.p);
		g_comm.allreduce(prot_max.p, n_prot, Comm::T_INT32, Comm::OP_MAX, st);
		if (N) PGX_LAUNCH(k_cap_apply, grid_for(N, TB), TB, 0, st, F.pid, F.sdom, N, prot_max.p);
		h_prot_max.resize(n_prot);
		PGX_CHECK(cudaMemcpyAsync(h_prot_max.data(), prot_max.p, (size_t)n_prot * 4, cudaMemcpyDeviceToHost, st));
		// per-protein sums (hit.c:196-204, 158-169) over the rank-0 unfiltered hits, one pass over the full columns, one reduction across
		// ranks.  The sums go to the host at once: the representative of every gene needs the exact permutation of the reference's unstable
		// radix sort (hit.c:209), ~2.5 ms on one host thread -- it runs beside the device work below and is joined where its
		// result is first needed (vertex election, materialisation).
		ensure_rep();
		z.zero(st); cs4.zero(st);
		if (N) PGX_LAUNCH(k_prot_stats, grid_for(N, TB), TB, 0, st, F.pid, F.sadj, F.sori, F.rank, F.bits, F.eoff, N, z.p, cs4.p);
		g_comm.allreduce(z.p, n_prot, Comm::T_UINT64, Comm::OP_SUM, st);
		g_comm.allreduce(cs4.p, (size_t)n_prot * 4, Comm::T_UINT64, Comm::OP_SUM, st);
		PGX_CHECK(cudaMemcpyAsync(h_z_pin, z.p, (size_t)n_prot * 8, cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaEventRecord(ev_z, st));
		{
			int dev = 0; PGX_CHECK(cudaGetDevice(&dev));
			rep_thread = std::thread([this, dev]() {
				cudaSetDevice(dev);
				cudaEventSynchronize(ev_z);
				choose_representatives(h_z_pin, h_prot_gid.data(), n_prot, n_gene, rep);  // host: exact tie order
			});
			rep_pending = true;
		}
		build_active();
		if (!(opt->flag & PG_F_NO_JOINT_PSEUDO) && Nw)
			PGX_LAUNCH(k_pseudo_joint, grid_for(Nw, TB), TB, 0, st, A.pid, A.bits, A.eoff, Nw, cs4.p, Gtot, opt->min_vertex_ratio, !!(opt->flag & PG_F_DROP_SGL_EXON));
		if (Nw) {
			launch_sweep<SW_SHADOW>(sweep_active(), so, A.bits_alt, A.pdom, nullptr);
			a_bits.swap(a_bits_alt); set_active();
		}
		done_post = true;
		g_stats.n_active = Nw;
	}

	// joins the host thread that picks the representatives; uploads the per-protein flag the materialisation needs
	void ensure_rep()
	{
		if (!rep_pending) return;
		const double t0 = wall_s();
		rep_thread.join();
		rep_pending = false;
		PGX_CHECK(cudaMemcpyAsync(prot_rep.p, rep.prot_rep.data(), (size_t)n_prot * 4, cudaMemcpyHostToDevice, st));
		if (getenv("PGX_TIMING")) fprintf(stderr, "[pgx timing] waited %.2f ms for the representatives (host thread)\n", (wall_s() - t0) * 1e3);
	}

	void build_active()
	{
		size_t mk = scratch.mark();
		Nw = 0;
		if (N) {
			uint32_t *flag = scratch.get<uint32_t>(N + 1), *pos = scratch.get<uint32_t>((size_t)N + 2);
			PGX_LAUNCH(k_flag_unfiltered, grid_for(N, TB), TB, 0, st, F.bits, N, flag);
			Nw = (int32_t)scan_flags(flag, pos, N);
			size_t n1 = (size_t)Nw + 16;
			a_key.ensure(n1); a_ctg.ensure(n1); a_cs.ensure(n1); a_ce.ensure(n1); a_cm.ensure(n1); a_pid.ensure(n1); a_gidp.ensure(n1); a_sadj.ensure(n1); a_rank.ensure(n1); a_sori.ensure(n1);
			a_sdom.ensure(n1); a_pdom0.ensure(n1); a_pdom.ensure(n1); a_cds.ensure(n1); a_pmax.ensure(n1); a_full.ensure(n1); a_bits.ensure(n1); a_bits_alt.ensure(n1); a_nex.ensure(n1); a_eoff.ensure(n1 + 16);
			by_cm.ensure(n1); sv.ensure(n1);
			set_active();
			PGX_LAUNCH(k_build_active, grid_for(N, TB), TB, 0, st, F, pos, N, A);
			// where each local genome starts inside the active set
			goff_act.ensure(G + 1);
			PGX_LAUNCH(k_gather_u32, grid_for(G + 1, TB), TB, 0, st, pos, (const uint32_t*)hoff.p, G + 1, (uint32_t*)goff_act.p);
		} else { goff_act.ensure(G + 1); goff_act.zero(st); }
		if (Nw) {
			uint32_t *agg = scratch.get<uint32_t>(scan_scratch_elems(Nw));
			exclusive_sum_u32(A.nex, A.eoff, Nw, agg, st);
			uint32_t ne = read_u32(A.eoff + Nw);
			wex.ensure((size_t)ne + 16);
			PGX_LAUNCH(k_copy_exons, grid_for(Nw, 128), 128, 0, st, ex.p, F.eoff, A.full, A.eoff, Nw, wex.p);
			SegVal *sagg = scratch.get<SegVal>((size_t)(Nw / SCAN_TILE + 2));
			segmented_running_max(A.ctg, A.ce, A.pmax, Nw, sagg, st);
			// (contig, cm) order as a permutation; ties keep (contig, cs) order
			uint64_t *key = scratch.get<uint64_t>((size_t)Nw + 1);
			PGX_LAUNCH(k_make_cm_keys, grid_for(Nw, TB), TB, 0, st, A.ctg, A.cm, Nw, key);
			PGX_LAUNCH(k_pack_sort_key, grid_for(Nw, TB), TB, 0, st, key, Nw, coord_bits);
			uint32_t *perm = sort_perm(key, Nw, coord_bits + bits_for((uint64_t)(n_ctg > 0? n_ctg - 1 : 0)));
			PGX_CHECK(cudaMemcpyAsync(by_cm.p, perm, (size_t)Nw * 4, cudaMemcpyDeviceToDevice, st));
		}
		scratch.reset(mk);
	}

	// Shrink the active set to the hits that are still unfiltered.  `flt` only grows and a filtered hit is never touched again
	// (sweeps skip it, weak_br is only raised on survivors), so its state is final: it is written back to the full columns here
	// and dropped.  After the vertex filter ~45 % of the active hits go; every later round (17 sweeps, 34 flag/scan passes,
	// 15 position tables) runs on the smaller set.
	void recompact_active()
	{
		if (Nw == 0) return;
		size_t mk = scratch.mark();
		uint32_t *flag = scratch.get<uint32_t>((size_t)Nw + 1), *pos = scratch.get<uint32_t>((size_t)Nw + 2);
		PGX_LAUNCH(k_flag_unfiltered, grid_for(Nw, TB), TB, 0, st, A.bits, Nw, flag);
		const int32_t n2 = (int32_t)scan_flags(flag, pos, Nw);
		if ((double)n2 > 0.85 * Nw) { scratch.reset(mk); return; }
		PGX_LAUNCH(k_scatter_back, grid_for(Nw, TB), TB, 0, st, A, Nw, F);
		const size_t n1 = (size_t)n2 + 16;
		DBuf<int32_t> b_ctg, b_cs, b_ce, b_cm, b_pid, b_gidp, b_sadj, b_rank, b_sori, b_sdom, b_pdom0, b_pdom, b_cds, b_pmax, b_full, b_goff;
		DBuf<uint32_t> b_bits, b_bits_alt, b_nex, b_eoff, b_by_cm;
		DBuf<uint64_t> b_key;
		DBuf<int2> b_wex;
		b_ctg.alloc(n1); b_cs.alloc(n1); b_ce.alloc(n1); b_cm.alloc(n1); b_pid.alloc(n1); b_gidp.alloc(n1); b_sadj.alloc(n1); b_rank.alloc(n1); b_sori.alloc(n1);
		b_sdom.alloc(n1); b_pdom0.alloc(n1); b_pdom.alloc(n1); b_cds.alloc(n1); b_pmax.alloc(n1); b_full.alloc(n1); b_bits.alloc(n1); b_bits_alt.alloc(n1); b_nex.alloc(n1);
		b_eoff.alloc(n1 + 16); b_by_cm.alloc(n1); b_key.alloc(n1); b_goff.alloc(G + 1);
		ActCols B;
		B.ctg = b_ctg.p; B.cs = b_cs.p; B.ce = b_ce.p; B.cm = b_cm.p; B.pid = b_pid.p; B.gidp = b_gidp.p; B.sadj = b_sadj.p; B.rank = b_rank.p; B.sori = b_sori.p;
		B.sdom = b_sdom.p; B.pdom0 = b_pdom0.p; B.pdom = b_pdom.p; B.cds = b_cds.p; B.pmax = b_pmax.p; B.full = b_full.p; B.bits = b_bits.p; B.bits_alt = b_bits_alt.p; B.nex = b_nex.p; B.eoff = b_eoff.p; B.key = b_key.p;
		int32_t *old_of_new = scratch.get<int32_t>(n1);
		PGX_LAUNCH(k_recompact_active, grid_for(Nw, TB), TB, 0, st, A, pos, Nw, B, old_of_new);
		PGX_LAUNCH(k_gather_u32, grid_for(G + 1, TB), TB, 0, st, pos, (const uint32_t*)goff_act.p, G + 1, (uint32_t*)b_goff.p);
		if (n2) {
			uint32_t *agg = scratch.get<uint32_t>(scan_scratch_elems(n2));
			exclusive_sum_u32(B.nex, B.eoff, n2, agg, st);
			uint32_t ne = read_u32(B.eoff + n2);
			b_wex.alloc((size_t)ne + 16);
			PGX_LAUNCH(k_copy_exons, grid_for(n2, 128), 128, 0, st, wex.p, A.eoff, old_of_new, B.eoff, n2, b_wex.p);
			SegVal *sagg = scratch.get<SegVal>((size_t)(n2 / SCAN_TILE + 2));
			segmented_running_max(B.ctg, B.ce, B.pmax, n2, sagg, st);
			uint32_t *fcm = scratch.get<uint32_t>((size_t)Nw + 1), *pcm = scratch.get<uint32_t>((size_t)Nw + 2);
			PGX_LAUNCH(k_flag_alive_cm, grid_for(Nw, TB), TB, 0, st, pos, by_cm.p, Nw, fcm);
			uint32_t *agg2 = scratch.get<uint32_t>(scan_scratch_elems(Nw));
			exclusive_sum_u32(fcm, pcm, Nw, agg2, st);
			PGX_LAUNCH(k_renumber_cm, grid_for(Nw, TB), TB, 0, st, pos, by_cm.p, pcm, Nw, b_by_cm.p);
		} else b_wex.alloc(16);
		PGX_CHECK(cudaStreamSynchronize(st));   // the old buffers go back to the cache below
		a_ctg.swap(b_ctg); a_cs.swap(b_cs); a_ce.swap(b_ce); a_cm.swap(b_cm); a_pid.swap(b_pid); a_gidp.swap(b_gidp); a_sadj.swap(b_sadj); a_rank.swap(b_rank); a_sori.swap(b_sori);
		a_sdom.swap(b_sdom); a_pdom0.swap(b_pdom0); a_pdom.swap(b_pdom); a_cds.swap(b_cds); a_pmax.swap(b_pmax); a_full.swap(b_full); a_bits.swap(b_bits); a_bits_alt.swap(b_bits_alt);
		a_nex.swap(b_nex); a_eoff.swap(b_eoff); by_cm.swap(b_by_cm); a_key.swap(b_key); wex.swap(b_wex); goff_act.swap(b_goff);
		Nw = n2;
		set_active();
		scratch.reset(mk);
	}

	// ------------------------------------------------------------------------------------------------
	// stage 2: pg_graph_gen (graph.c:280-322)
	// ------------------------------------------------------------------------------------------------
	void set_filter(int mode) { if (Nw) PGX_LAUNCH(k_set_filter, grid_for(Nw, TB), TB, 0, st, A.bits, Nw, mode); }
	void flag_vtx() { if (Nw) PGX_LAUNCH(k_flag_vtx, grid_for(Nw, TB), TB, 0, st, A.gidp, A.bits, Nw, g2s.p, 1); }
	void regen_g2s()
	{
		PGX_CHECK(cudaMemsetAsync(g2s.p, 0xff, (size_t)n_gene * 4, st));
		if (n_seg0) PGX_LAUNCH(k_g2s_fill, GS_GRID, TB, 0, st, seg.p, dc.p, g2s.p);
	}
	uint32_t *dcp(size_t field_off) { return (uint32_t*)((char*)dc.p + field_off); }
	static constexpr int GS_GRID = 148 * 8;      // grid-stride kernels: 8 CTAs of 256 threads per SM
	// exclusive scan of n flags into pos[n+1]; the total goes to a DevCounts field, not to the host
	void scan_to_count(const uint32_t *flag, uint32_t *pos, int64_t n, size_t field_off)
	{
		uint32_t *agg = scratch.get<uint32_t>(scan_scratch_elems(n));
		exclusive_sum_u32(flag, pos, n, agg, st);
		PGX_LAUNCH(k_copy_u32, 1, 1, 0, st, dcp(field_off), pos + n);
	}
	DevCounts read_counts()
	{
		DevCounts h;
		PGX_CHECK(cudaMemcpyAsync(&h, dc.p, sizeof(h), cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		return h;
	}

	void gen_vtx(const pg_opt_t *opt)   // vertex.c:6-100
	{
		size_t mk = scratch.mark();
		size_t tab = (size_t)G * n_gene;
		uint8_t *t_dom = scratch.get<uint8_t>(tab + 1);
		int32_t *t_first = scratch.get<int32_t>(tab + 1);
		int32_t *aux = scratch.get<int32_t>(tab + 1);
		PGX_CHECK(cudaMemsetAsync(t_dom, 0, tab + 1, st));
		PGX_LAUNCH(k_fill_i32, grid_for((int64_t)tab + 1, TB), TB, 0, st, t_first, (int64_t)tab + 1, INT32_MAX);
		if (Nw) PGX_LAUNCH(k_vtx_mark, grid_for(Nw, TB), TB, 0, st, A.ctg, ctg2g.p, A.gidp, A.rank, A.bits, Nw, G, t_dom, t_first);
		if (tab) PGX_LAUNCH(k_vtx_resolve, grid_for((int64_t)tab, TB), TB, 0, st, t_dom, t_first, A.pdom, prot_gid.p, tab, aux, counters.p + C_ERR);
		double tv0 = wall_s();
		// the election's input in compact form: per-gene counts + the shadowed cells (k_vtx_cells); a few hundred KB instead of the whole table
		int32_t *cnt2 = scratch.get<int32_t>(2 * (size_t)n_gene + 2);
		uint32_t *n_cells_d = scratch.get<uint32_t>(g_comm.nranks + 1);
		if (tab >= 0xfffffff0u) PGX_FATAL("vertex selection: the (gene, genome) table has %zu cells; 32-bit cell indices", tab);
		const uint32_t cap0 = (uint32_t)tab + 16;                                              // every cell could be a shadowed one
		VtxCell *cells_d = scratch.get<VtxCell>((size_t)cap0);
		PGX_CHECK(cudaMemsetAsync(cnt2, 0, (2 * (size_t)n_gene + 2) * 4, st));
		PGX_CHECK(cudaMemsetAsync(n_cells_d, 0, ((size_t)g_comm.nranks + 1) * 4, st));
		if (tab) PGX_LAUNCH(k_vtx_cells, grid_for((int64_t)tab, TB), TB, 0, st, aux, tab, G, g_comm.genome_begin, cnt2, cells_d, cap0, n_cells_d + g_comm.rank);
		g_comm.allreduce(cnt2, 2 * (size_t)n_gene, Comm::T_INT32, Comm::OP_SUM, st);
		g_comm.allgather(n_cells_d + g_comm.rank, n_cells_d, 4, st);
		std::vector<int32_t> h_cnt2(2 * (size_t)n_gene + 2);
		std::vector<uint32_t> h_nc(g_comm.nranks);
		PGX_CHECK(cudaMemcpyAsync(h_cnt2.data(), cnt2, 2 * (size_t)n_gene * 4, cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaMemcpyAsync(h_nc.data(), n_cells_d, (size_t)g_comm.nranks * 4, cudaMemcpyDeviceToHost, st));
		check_err("vertex selection");
		uint32_t nc_max = 0; size_t nc_tot = 0;
		for (uint32_t c : h_nc) { nc_max = c > nc_max? c : nc_max; nc_tot += c; }
		std::vector<VtxCellHost> h_cells(nc_tot + 1);
		static_assert(sizeof(VtxCell) == sizeof(VtxCellHost), "cell record");
		if (g_comm.nranks > 1) {
			VtxCell *all = scratch.get<VtxCell>((size_t)nc_max * g_comm.nranks + 1);
			if (nc_max) {
				if (h_nc[g_comm.rank]) PGX_CHECK(cudaMemcpyAsync(all + (size_t)nc_max * g_comm.rank, cells_d, (size_t)h_nc[g_comm.rank] * sizeof(VtxCell), cudaMemcpyDeviceToDevice, st));
				g_comm.allgather(all + (size_t)nc_max * g_comm.rank, all, (size_t)nc_max * sizeof(VtxCell), st);
				size_t o = 0;
				for (int r = 0; r < g_comm.nranks; ++r) { if (h_nc[r]) PGX_CHECK(cudaMemcpyAsync(h_cells.data() + o, all + (size_t)nc_max * r, (size_t)h_nc[r] * sizeof(VtxCell), cudaMemcpyDeviceToHost, st)); o += h_nc[r]; }
			}
		} else if (nc_tot) PGX_CHECK(cudaMemcpyAsync(h_cells.data(), cells_d, nc_tot * sizeof(VtxCell), cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		double tv1 = wall_s();
		ensure_rep();
		elect_vertices_cells(h_cnt2.data(), h_cells.data(), nc_tot, Gtot, n_gene, rep.gene_rep_pid.data(), rep.prot_avg.data(), h_gene_pref.data(), h_gene_incl.data(), opt->min_vertex_ratio,
			!!(opt->flag & PG_F_WRITE_VTX_SEL), h_gene_names, h_seg0);
		if (getenv("PGX_TIMING")) fprintf(stderr, "[pgx timing] gen_vtx: counts + %zu cells D2H %.2f ms, host election %.2f ms\n", nc_tot, (tv1 - tv0) * 1e3, (wall_s() - tv1) * 1e3);
		n_seg = n_seg0 = (int32_t)h_seg0.size();
		std::vector<SegDev> hs(n_seg > 0? n_seg : 1);
		for (int32_t i = 0; i < n_seg; ++i) { memset(&hs[i], 0, sizeof(SegDev)); hs[i].gid = h_seg0[i].gid; hs[i].n_dom = h_seg0[i].n_dom; hs[i].n_sub = h_seg0[i].n_sub; }
		seg.ensure(hs.size()); seg_alt.ensure(hs.size()); seg_ng.ensure(hs.size()); seg_tot.ensure(hs.size()); idx.ensure(hs.size() * 2);
		arc_cap = 2 * (size_t)(Nw > 0? Nw : 1);
		arc.ensure(arc_cap + 1); arc_alt.ensure(arc_cap + 1);
		dc.ensure(1);
		DevCounts h0; memset(&h0, 0, sizeof(h0)); h0.n_seg = (uint32_t)n_seg0;
		PGX_CHECK(cudaMemcpyAsync(dc.p, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaMemcpyAsync(seg.p, hs.data(), hs.size() * sizeof(SegDev), cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		regen_g2s();
		scratch.reset(mk);
		if (pg_verbose >= 3) fprintf(stderr, "[M::pgx::gen_vtx] selected %d vertices out of %d genes\n", n_seg, n_gene);
	}

	// graph.c:87-177.  No host round-trip in the single-GPU path: survivor count, group count and arc count stay on the device.
	void gen_arc(const pg_opt_t *opt)
	{
		struct Acc { double &a; double t0; ~Acc() { a += wall_s() - t0; } } acc_arc{t_arc, wall_s()};
		set_first_for_this_call();
		if (Nw == 0 || n_seg0 == 0) {   // nothing can survive: empty arc set, zero occupancy
			PGX_LAUNCH(k_set_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, n_arc)), 0u);
			PGX_LAUNCH(k_set_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, n_part)), 0u);
			PGX_LAUNCH(k_set_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, m_sv)), 0u);
			m_sv = 0;
			if (n_seg0) {   // n_seg0 is replicated: every rank takes the same branch, so the collectives below stay matched
				size_t mk = scratch.mark();
				PGX_CHECK(cudaMemsetAsync(seg_ng.p, 0, (size_t)n_seg0 * 4, st));
				PGX_CHECK(cudaMemsetAsync(seg_tot.p, 0, (size_t)n_seg0 * 4, st));
				if (g_comm.active()) merge_partials_across_ranks(nullptr);
				g_comm.allreduce(seg_ng.p, n_seg0, Comm::T_INT32, Comm::OP_SUM, st);
				g_comm.allreduce(seg_tot.p, n_seg0, Comm::T_INT32, Comm::OP_SUM, st);
				PGX_LAUNCH(k_seg_store_occ, GS_GRID, TB, 0, st, seg.p, dc.p, seg_ng.p, seg_tot.p);
				scratch.reset(mk);
			}
			return;
		}
		size_t mk = scratch.mark();
		SweepOpt so; so.min_ov_ratio = opt->min_ov_ratio; so.check_strand = !!(opt->flag & PG_F_CHECK_STRAND);
		PGX_CHECK(cudaMemsetAsync(seg_ng.p, 0, (size_t)n_seg0 * 4, st));
		PGX_CHECK(cudaMemsetAsync(seg_tot.p, 0, (size_t)n_seg0 * 4, st));
		launch_sweep<SW_SHADOW>(sweep_active(), so, A.bits_alt, A.pdom, nullptr);
		a_bits.swap(a_bits_alt); set_active();
		{
			uint32_t *flag = scratch.get<uint32_t>((size_t)Nw + 1), *pos = scratch.get<uint32_t>((size_t)Nw + 2);
			PGX_LAUNCH(k_flag_survivor_cm, grid_for(Nw, TB), TB, 0, st, A.bits, by_cm.p, Nw, flag);
			scan_to_count(flag, pos, Nw, offsetof(DevCounts, m_sv));
			PGX_LAUNCH(k_compact_idx, grid_for(Nw, TB), TB, 0, st, pos, by_cm.p, Nw, sv.p);
			m_sv = (int32_t)read_u32(pos + Nw);      // one read-back per arc round: it sizes the record sort (survivors are ~half of the active set)
		}
		{   // cm-tied survivors in the reference's order; pairs completed before this call's cm sort: 0, 1, then 3 + 2 i in branch round i
			const int32_t m = arc_call - 1;
			fix_cm_tied_survivors(m < 2? m : 3 + 2 * (m - 2));
		}
		if (m_sv == 0) {
			PGX_LAUNCH(k_set_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, n_arc)), 0u);
			PGX_LAUNCH(k_set_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, n_part)), 0u);
			if (g_comm.active()) merge_partials_across_ranks(nullptr);
			g_comm.allreduce(seg_ng.p, n_seg0, Comm::T_INT32, Comm::OP_SUM, st);
			g_comm.allreduce(seg_tot.p, n_seg0, Comm::T_INT32, Comm::OP_SUM, st);
			PGX_LAUNCH(k_seg_store_occ, GS_GRID, TB, 0, st, seg.p, dc.p, seg_ng.p, seg_tot.p);
			scratch.reset(mk);
			return;
		}
		const size_t nr = 2 * (size_t)m_sv, bm = ((size_t)G * n_seg0 + 31) / 32 + 1;
		uint64_t *key = scratch.get<uint64_t>(nr + 1);
		int32_t *gen = scratch.get<int32_t>(nr), *dist = scratch.get<int32_t>(nr), *s1 = scratch.get<int32_t>(nr), *s2 = scratch.get<int32_t>(nr);
		uint32_t *seen = scratch.get<uint32_t>(bm);
		PGX_CHECK(cudaMemsetAsync(seen, 0, bm * 4, st));
		ArcEmit e; e.sv = sv.p; e.cap = m_sv; e.ctg = A.ctg; e.cm = A.cm; e.gidp = A.gidp; e.sori = A.sori; e.sdom = A.sdom; e.pdom0 = A.pdom0; e.bits = A.bits;
		e.g2s = g2s.p; e.prot_gid = prot_gid.p; e.ctg2g = ctg2g.p; e.use_ori = !!(opt->flag & PG_F_ORI_FOR_BRANCH); e.n_seg = n_seg0; e.gbits = bits_for((uint64_t)(G > 1? G - 1 : 1));
		PGX_LAUNCH(k_arc_emit, GS_GRID, TB, 0, st, e, dc.p, key, gen, dist, s1, s2, seen, seg_ng.p, seg_tot.p, counters.p + C_TIE_CM, counters.p + C_ERR);
		uint64_t nv = 2ull * n_seg0;
		uint32_t *rec = sort_perm(key, (int64_t)nr, bits_for(nv * nv - 1) + 1);     // +1: the all-ones "no arc" key sorts strictly last
		uint32_t *flag = scratch.get<uint32_t>(nr + 1), *pos = scratch.get<uint32_t>(nr + 2), *start = scratch.get<uint32_t>(nr + 1);
		PGX_LAUNCH(k_arc_heads, grid_for((int64_t)nr, TB), TB, 0, st, key, (int32_t)nr, flag);
		scan_to_count(flag, pos, (int64_t)nr, offsetof(DevCounts, n_part));
		PGX_LAUNCH(k_compact_idx, grid_for((int64_t)nr, TB), TB, 0, st, pos, (const uint32_t*)nullptr, (int32_t)nr, start);
		ArcPartial *part = scratch.get<ArcPartial>(nr + 1);
		arc_of_rec.ensure(nr + 2);
		PGX_LAUNCH(k_arc_reduce, GS_GRID, TB, 0, st, key, rec, gen, dist, s1, s2, (int32_t)nr, n_seg0, start, dc.p, part, arc_of_rec.p);
		if (g_comm.active()) merge_partials_across_ranks(part);
		else {
			PGX_LAUNCH(k_arc_finalize_direct, GS_GRID, TB, 0, st, part, dc.p, arc.p);
			PGX_LAUNCH(k_copy_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, n_arc)), dcp(offsetof(DevCounts, n_part)));
		}
		g_comm.allreduce(seg_ng.p, n_seg0, Comm::T_INT32, Comm::OP_SUM, st);
		g_comm.allreduce(seg_tot.p, n_seg0, Comm::T_INT32, Comm::OP_SUM, st);
		PGX_LAUNCH(k_seg_store_occ, GS_GRID, TB, 0, st, seg.p, dc.p, seg_ng.p, seg_tot.p);
		scratch.reset(mk);
	}

	// multi-GPU: every rank contributes its partial table (padded to the largest one); sort by key, merge equal keys (graph.c:150-175)
	double t_merge = 0, t_branch = 0, t_arc = 0, t_vtx = 0;   // host wall time per piece (PGX_TIMING)
	void merge_partials_across_ranks(const ArcPartial *part)
	{
		double tm0 = wall_s();
		struct Acc { double &a; double t0; ~Acc() { a += wall_s() - t0; } } acc{t_merge, tm0};
		uint32_t n_part = read_counts().n_part;
		uint32_t *cnt = scratch.get<uint32_t>(g_comm.nranks + 1);
		PGX_CHECK(cudaMemcpyAsync(cnt + g_comm.rank, &n_part, 4, cudaMemcpyHostToDevice, st));
		g_comm.allgather(cnt + g_comm.rank, cnt, 4, st);
		std::vector<uint32_t> hc(g_comm.nranks);
		PGX_CHECK(cudaMemcpyAsync(hc.data(), cnt, 4 * g_comm.nranks, cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		uint32_t mx = 0; uint64_t tot = 0;
		for (uint32_t c : hc) { mx = c > mx? c : mx; tot += c; }
		if (tot == 0) { PGX_LAUNCH(k_set_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, n_arc)), 0u); return; }
		if (tot + 1 > arc.n) { arc.alloc(tot + tot / 4 + 16); arc_alt.alloc(tot + tot / 4 + 16); arc_cap = arc.n - 1; }   // the merged table can exceed this rank's own capacity
		ArcPartial *all = scratch.get<ArcPartial>((size_t)mx * g_comm.nranks + 1), *mine = all + (size_t)mx * g_comm.rank;
		if (n_part) PGX_CHECK(cudaMemcpyAsync(mine, part, (size_t)n_part * sizeof(ArcPartial), cudaMemcpyDeviceToDevice, st));
		g_comm.allgather(mine, all, (size_t)mx * sizeof(ArcPartial), st);
		ArcPartial *packed = scratch.get<ArcPartial>(tot + 1), *sorted = scratch.get<ArcPartial>(tot + 1);
		uint64_t o = 0;
		for (int r = 0; r < g_comm.nranks; ++r) { if (hc[r]) PGX_CHECK(cudaMemcpyAsync(packed + o, all + (size_t)mx * r, (size_t)hc[r] * sizeof(ArcPartial), cudaMemcpyDeviceToDevice, st)); o += hc[r]; }
		uint64_t *key = scratch.get<uint64_t>(tot + 1);
		PGX_LAUNCH(k_partial_keys, grid_for((int64_t)tot, TB), TB, 0, st, packed, (int32_t)tot, n_seg0, key);
		uint64_t nv = 2ull * n_seg0;
		uint32_t *perm = sort_perm(key, (int64_t)tot, bits_for(nv * nv - 1));
		PGX_LAUNCH(k_partial_gather, grid_for((int64_t)tot, TB), TB, 0, st, packed, perm, (int32_t)tot, sorted);
		uint32_t *flag = scratch.get<uint32_t>((size_t)tot + 1), *pos = scratch.get<uint32_t>((size_t)tot + 2);
		PGX_LAUNCH(k_partial_heads, grid_for((int64_t)tot, TB), TB, 0, st, sorted, (int32_t)tot, flag);
		scan_to_count(flag, pos, (int64_t)tot, offsetof(DevCounts, n_arc));
		PGX_LAUNCH(k_arc_finalize, grid_for((int64_t)tot, TB), TB, 0, st, sorted, (int32_t)tot, pos, arc.p);
	}

	void arc_index()   // graph.c:202-217
	{
		PGX_CHECK(cudaMemsetAsync(idx.p, 0, (size_t)(n_seg0 > 0? n_seg0 : 1) * 16, st));
		PGX_LAUNCH(k_arc_index, GS_GRID, TB, 0, st, arc.p, dc.p, idx.p);
	}

	void flt_high_occ(int32_t max_avg_occ, int32_t max_degree, int32_t max_dist_loci)   // graph.c:237-263
	{
		if (n_seg0) {
			size_t mk = scratch.mark();
			uint32_t *keep = scratch.get<uint32_t>((size_t)n_seg0 + 1), *pos = scratch.get<uint32_t>((size_t)n_seg0 + 2);
			PGX_LAUNCH(k_seg_mark_del, grid_for(n_seg0, TB), TB, 0, st, seg.p, n_seg0, dc.p, arc.p, max_avg_occ * Gtot, max_degree, max_dist_loci, keep);
			uint32_t *agg = scratch.get<uint32_t>(scan_scratch_elems(n_seg0));
			exclusive_sum_u32(keep, pos, n_seg0, agg, st);
			PGX_LAUNCH(k_seg_compact, GS_GRID, TB, 0, st, seg.p, dc.p, pos, seg_alt.p);
			PGX_LAUNCH(k_copy_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, n_seg)), pos + n_seg0);
			seg.swap(seg_alt);
			scratch.reset(mk);
		}
		regen_g2s();
		flag_vtx();
	}

	// The reference's order of the hits of some contigs after `pairs` (cm, cs) sort pairs of its unstable radix-sort chain:
	// rank_of[k][i] = position of the contig's i-th hit (our (contig, cs) order) in the reference's array, either as the last cs sort left it
	// (by_cm = false: branch.c:18-25 numbers survivors in that order) or after one more sort by cm (by_cm = true: the order pg_gen_arc and
	// pg_mark_branch_flt_hit walk, graph.c:103, branch.c:116).  Chains are fetched once per contig and advanced from call to call.
	void replay_chains(const std::vector<int32_t> &ctgs, int32_t pairs, bool by_cm, std::vector<std::vector<int32_t>> &rank_of)
	{
		std::vector<int32_t> fresh;
		for (int32_t c : ctgs) if (!tie_chains.count(c)) fresh.push_back(c);
		if (!fresh.empty()) {   // fetch (cs, cm, file position) of the contigs seen for the first time
			size_t mk = scratch.mark();
			int32_t *d_ids = scratch.get<int32_t>(fresh.size());
			int2 *d_rng = scratch.get<int2>(fresh.size());
			PGX_CHECK(cudaMemcpyAsync(d_ids, fresh.data(), fresh.size() * 4, cudaMemcpyHostToDevice, st));
			PGX_LAUNCH(k_ctg_ranges, grid_for((int64_t)fresh.size(), TB), TB, 0, st, d_ids, (int32_t)fresh.size(), F.ctg, N, d_rng);
			std::vector<int2> rng(fresh.size());
			PGX_CHECK(cudaMemcpyAsync(rng.data(), d_rng, fresh.size() * sizeof(int2), cudaMemcpyDeviceToHost, st));
			PGX_CHECK(cudaStreamSynchronize(st));
			std::vector<std::vector<uint32_t>> h_src(fresh.size());
			for (size_t k = 0; k < fresh.size(); ++k) {
				TieChain &c = tie_chains[fresh[k]];
				c.lo = rng[k].x; c.n = rng[k].y - rng[k].x; c.pairs_done = 0;
				c.cs.resize(c.n); c.cm.resize(c.n); h_src[k].resize(c.n);
				PGX_CHECK(cudaMemcpyAsync(c.cs.data(), F.cs + c.lo, (size_t)c.n * 4, cudaMemcpyDeviceToHost, st));
				PGX_CHECK(cudaMemcpyAsync(c.cm.data(), F.cm + c.lo, (size_t)c.n * 4, cudaMemcpyDeviceToHost, st));
				PGX_CHECK(cudaMemcpyAsync(h_src[k].data(), src.p + c.lo, (size_t)c.n * 4, cudaMemcpyDeviceToHost, st));
			}
			PGX_CHECK(cudaStreamSynchronize(st));
			ThreadPool::get().parallel_for((int64_t)fresh.size(), [&](int64_t k) {
				TieChain &c = tie_chains.find(fresh[k])->second;
				std::vector<int32_t> ord(c.n);
				for (int32_t i = 0; i < c.n; ++i) ord[i] = i;
				std::sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) { return h_src[k][x] < h_src[k][y]; });   // file order
				c.a.resize(c.n);
				for (int32_t i = 0; i < c.n; ++i) { c.a[i].y = (uint64_t)ord[i]; c.a[i].x = (uint64_t)c.cs[ord[i]]; }
				FlagSortReplay::run(c.a.data(), c.n);                                                                   // read.c:247
			});
			scratch.reset(mk);
		}
		rank_of.assign(ctgs.size(), std::vector<int32_t>());
		ThreadPool::get().parallel_for((int64_t)ctgs.size(), [&](int64_t k) {
			TieChain &c = tie_chains.find(ctgs[k])->second;
			for (; c.pairs_done < pairs; ++c.pairs_done) {
				for (int32_t i = 0; i < c.n; ++i) c.a[i].x = (uint64_t)c.cm[c.a[i].y];
				FlagSortReplay::run(c.a.data(), c.n);                                                                   // graph.c:103, branch.c:116
				for (int32_t i = 0; i < c.n; ++i) c.a[i].x = (uint64_t)c.cs[c.a[i].y];
				FlagSortReplay::run(c.a.data(), c.n);                                                                   // graph.c:123, branch.c:140
			}
			rank_of[k].resize(c.n);
			if (!by_cm) { for (int32_t i = 0; i < c.n; ++i) rank_of[k][c.a[i].y] = i; return; }
			std::vector<pg128_t> b(c.a);                                                                                // one more sort, by cm, on a copy
			for (int32_t i = 0; i < c.n; ++i) b[i].x = (uint64_t)c.cm[b[i].y];
			FlagSortReplay::run(b.data(), c.n);
			for (int32_t i = 0; i < c.n; ++i) rank_of[k][b[i].y] = i;
		});
	}

	// Survivors that share one (contig, cm) with their neighbour in the (contig, cm) list: the reference walks them in the order its sort chain
	// left them in (graph.c:103-122, branch.c:116-136) -- the direction of the arc between two of them and the arcs to their outer neighbours
	// depend on it (DESIGN.md "tie order", tie[0]).  Careful path only: lists the tied positions of `sv`, replays the chains of their contigs
	// on the host and re-orders every tied group.  Returns true if `sv` changed.  pairs: (cm, cs) sort pairs completed before this cm sort.
	bool fix_cm_tied_survivors(int32_t pairs)
	{
		if (Nw == 0) return false;
		size_t mk = scratch.mark();
		uint32_t *list = scratch.get<uint32_t>((size_t)Nw + 1);
		PGX_CHECK(cudaMemsetAsync(counters.p + C_CMTIE_ROUND, 0, 4, st));
		PGX_LAUNCH(k_cm_tie_list, GS_GRID, TB, 0, st, sv.p, dc.p, A.ctg, A.cm, list, counters.p + C_CMTIE_ROUND);
		const uint32_t n_ties = read_u32(counters.p + C_CMTIE_ROUND);
		if (n_ties == 0) { scratch.reset(mk); return false; }
		TieRec *d_rec = scratch.get<TieRec>(n_ties + 1);
		PGX_LAUNCH(k_tie_gather, grid_for(n_ties, TB), TB, 0, st, list, (int32_t)n_ties, sv.p, A.ctg, A.full, d_rec);
		std::vector<TieRec> rec(n_ties);
		PGX_CHECK(cudaMemcpyAsync(rec.data(), d_rec, (size_t)n_ties * sizeof(TieRec), cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		std::sort(rec.begin(), rec.end(), [](const TieRec &x, const TieRec &y) { return x.t < y.t; });
		std::vector<int32_t> ctgs;
		for (const TieRec &r : rec) ctgs.push_back(r.ctg);
		std::sort(ctgs.begin(), ctgs.end()); ctgs.erase(std::unique(ctgs.begin(), ctgs.end()), ctgs.end());
		std::vector<std::vector<int32_t>> rank_of;
		replay_chains(ctgs, pairs, true, rank_of);
		std::vector<int2> patch;   // (position in sv, active index that belongs there)
		for (size_t i0 = 0, i1; i0 < rec.size(); i0 = i1) {
			for (i1 = i0 + 1; i1 < rec.size() && rec[i1].t == rec[i1 - 1].t + 1; ++i1) { }
			const size_t kc = std::lower_bound(ctgs.begin(), ctgs.end(), rec[i0].ctg) - ctgs.begin();
			const TieChain &c = tie_chains.find(rec[i0].ctg)->second;
			struct Mem { int32_t act, rank; };
			std::vector<Mem> mem;
			mem.push_back(Mem{rec[i0].a_prev, rank_of[kc][rec[i0].full_prev - c.lo]});
			for (size_t i = i0; i < i1; ++i) mem.push_back(Mem{rec[i].a, rank_of[kc][rec[i].full - c.lo]});
			std::vector<Mem> srt(mem);
			std::stable_sort(srt.begin(), srt.end(), [](const Mem &x, const Mem &y) { return x.rank < y.rank; });
			for (size_t k = 0; k < mem.size(); ++k) if (srt[k].act != mem[k].act) patch.push_back(make_int2(rec[i0].t - 1 + (int32_t)k, srt[k].act));
		}
		if (!patch.empty()) {
			int2 *d_patch = scratch.get<int2>(patch.size());
			PGX_CHECK(cudaMemcpyAsync(d_patch, patch.data(), patch.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
			PGX_LAUNCH(k_patch_u32, grid_for((int64_t)patch.size(), TB), TB, 0, st, d_patch, (int32_t)patch.size(), sv.p);
			PGX_CHECK(cudaStreamSynchronize(st));   // `patch` goes out of scope
		}
		scratch.reset(mk);
		return !patch.empty();
	}

	// branch.c:18-25 numbers the survivors of a genome in array order; survivors that share one (contig, cs) sit in the order the
	// reference's unstable sort chain left them in after `pairs` (cm, cs) sort pairs.  Replays the chain of the contigs concerned
	// (cached and advanced from round to round) and shifts the ordinals of the tied hits accordingly.
	int32_t branch_round_no = 0;
	void fix_tied_ordinals(int32_t pairs, const uint32_t *tie_list, uint32_t n_ties, const uint32_t *svcs, const int32_t *t_last, GenePos *gpos)
	{
		size_t mk = scratch.mark();
		TieRec *d_rec = scratch.get<TieRec>(n_ties + 1);
		PGX_LAUNCH(k_tie_gather, grid_for(n_ties, TB), TB, 0, st, tie_list, (int32_t)n_ties, svcs, A.ctg, A.full, d_rec);
		std::vector<TieRec> rec(n_ties);
		PGX_CHECK(cudaMemcpyAsync(rec.data(), d_rec, (size_t)n_ties * sizeof(TieRec), cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		std::sort(rec.begin(), rec.end(), [](const TieRec &x, const TieRec &y) { return x.t < y.t; });
		std::vector<int32_t> ctgs;
		for (const TieRec &r : rec) ctgs.push_back(r.ctg);
		std::sort(ctgs.begin(), ctgs.end()); ctgs.erase(std::unique(ctgs.begin(), ctgs.end()), ctgs.end());
		std::vector<std::vector<int32_t>> rank_of(ctgs.size());
		replay_chains(ctgs, pairs, false, rank_of);
		// groups of consecutive tied positions: members t0-1 .. tk hold the ordinals o .. o+m-1 in OUR order; hand them out in the reference's
		std::vector<int2> patch;
		for (size_t i0 = 0, i1; i0 < rec.size(); i0 = i1) {
			for (i1 = i0 + 1; i1 < rec.size() && rec[i1].t == rec[i1 - 1].t + 1; ++i1) { }
			const size_t kc = std::lower_bound(ctgs.begin(), ctgs.end(), rec[i0].ctg) - ctgs.begin();
			const TieChain &c = tie_chains.find(rec[i0].ctg)->second;
			struct Mem { int32_t act, rank, slot; };
			std::vector<Mem> mem;
			mem.push_back(Mem{rec[i0].a_prev, rank_of[kc][rec[i0].full_prev - c.lo], 0});
			for (size_t i = i0; i < i1; ++i) mem.push_back(Mem{rec[i].a, rank_of[kc][rec[i].full - c.lo], (int32_t)(i - i0) + 1});
			std::stable_sort(mem.begin(), mem.end(), [](const Mem &x, const Mem &y) { return x.rank < y.rank; });
			for (size_t k = 0; k < mem.size(); ++k) if ((int32_t)k != mem[k].slot) patch.push_back(make_int2(mem[k].act, (int32_t)k - mem[k].slot));
		}
		if (!patch.empty()) {
			int2 *d_patch = scratch.get<int2>(patch.size());
			PGX_CHECK(cudaMemcpyAsync(d_patch, patch.data(), patch.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
			PGX_LAUNCH(k_patch_ord, grid_for((int64_t)patch.size(), TB), TB, 0, st, d_patch, (int32_t)patch.size(), A.ctg, A.gidp, ctg2g.p, G, t_last, gpos);
			PGX_CHECK(cudaStreamSynchronize(st));   // `patch` goes out of scope
		}
		scratch.reset(mk);
	}

	void branch_round(const pg_opt_t *opt)   // branch.c:48-145 + graph.c:309
	{
		struct Acc { double &a; double t0; ~Acc() { a += wall_s() - t0; } } acc_br{t_branch, wall_s()};
		size_t mk = scratch.mark();
		arc_index();
		if (n_seg0) {   // (a rank without active hits still joins: the arc marks must come out identical everywhere)
			PGX_LAUNCH(k_seg_zero_loci, GS_GRID, TB, 0, st, seg.p, dc.p);
			// representative position of every gene in every local genome (branch.c:6-29)
			size_t tab = (size_t)n_gene * G;
			uint32_t *flag = scratch.get<uint32_t>((size_t)Nw + 1), *spos = scratch.get<uint32_t>((size_t)Nw + 2);
			PGX_LAUNCH(k_flag_survivor_cs, grid_for(Nw, TB), TB, 0, st, A.bits, Nw, flag);
			scan_to_count(flag, spos, Nw, offsetof(DevCounts, m_cs));
			uint32_t *svcs = scratch.get<uint32_t>((size_t)Nw + 1), *tie_list = scratch.get<uint32_t>((size_t)Nw + 1);
			PGX_LAUNCH(k_compact_idx, grid_for(Nw, TB), TB, 0, st, spos, (const uint32_t*)nullptr, Nw, svcs);
			int32_t *t_last = scratch.get<int32_t>(tab + 1);
			GenePos *gpos = scratch.get<GenePos>(tab + 1);
			PGX_CHECK(cudaMemsetAsync(t_last, 0xff, (tab + 1) * 4, st));
			PGX_LAUNCH(k_pos_last, grid_for(Nw, TB), TB, 0, st, A.ctg, ctg2g.p, A.gidp, spos, Nw, G, t_last);
			PGX_LAUNCH(k_pos_fill, grid_for((int64_t)tab, TB), TB, 0, st, t_last, tab, G, A.ctg, A.cm, spos, goff_act.p, gpos);
			// tie detector: two survivors with one (contig, cs) -- counted, listed, their genes marked in the position table
			PGX_CHECK(cudaMemsetAsync(counters.p + C_TIE_ROUND, 0, 12, st));
			PGX_LAUNCH(k_cs_tie_mark, GS_GRID, TB, 0, st, A.ctg, A.cs, A.gidp, ctg2g.p, G, svcs, dc.p, t_last, gpos, counters.p + C_TIE_CS, counters.p + C_TIE_ROUND, tie_list, counters.p + C_LONG_RUN);
			// n_local for every pair of out-arcs of every oriented vertex: local partial sums, one reduction, then the decisions
			const int32_t cap_vtx = 2 * n_seg0;
			uint32_t *cnt = scratch.get<uint32_t>((size_t)cap_vtx + 1), *off = scratch.get<uint32_t>((size_t)cap_vtx + 2);
			PGX_LAUNCH(k_branch_count, grid_for(cap_vtx, TB), TB, 0, st, idx.p, cap_vtx, dc.p, cnt);
			PGX_CHECK(cudaMemcpyAsync(h_pin + 1, counters.p + C_TIE_ROUND, 8, cudaMemcpyDeviceToHost, st));   // rides on the read-back below
			uint32_t n_nl = scan_flags(cnt, off, cap_vtx);       // the one host read-back of a round: it sizes the table below
			const uint32_t round_ties = h_pin[1], long_run = h_pin[2];
			int32_t *nl = scratch.get<int32_t>((size_t)n_nl + 1), *grp = scratch.get<int32_t>(arc_cap + 1);
			BranchOpt bo; bo.local_dist = opt->local_dist; bo.local_count = opt->local_count; bo.frag_mode = !!(opt->flag & PG_F_FRAG_MODE);
			bo.branch_diff = opt->branch_diff; bo.branch_diff_dist = opt->branch_diff_dist; bo.branch_diff_cut = opt->branch_diff_cut;
			if (n_nl) {
				PGX_LAUNCH(k_branch_nlocal, GS_GRID, TB, 0, st, idx.p, dc.p, arc.p, seg.p, gpos, G, off, (int32_t)n_nl, bo, nl, counters.p + C_ORD_SENS);
				// (rare) this round has cs-tied survivors AND one of their ordinals is within reach of the local_count threshold in a
				// comparison it decides: renumber them as the reference does and count again
				if (round_ties && (long_run || read_u32(counters.p + C_ORD_SENS))) {
					fix_tied_ordinals(2 + 2 * branch_round_no, tie_list, round_ties, svcs, t_last, gpos);
					PGX_LAUNCH(k_branch_nlocal, GS_GRID, TB, 0, st, idx.p, dc.p, arc.p, seg.p, gpos, G, off, (int32_t)n_nl, bo, nl, counters.p + C_ORD_SENS);
				}
				g_comm.allreduce(nl, n_nl, Comm::T_INT32, Comm::OP_SUM, st);
			}
			PGX_LAUNCH(k_branch_decide, GS_GRID, 128, 0, st, idx.p, dc.p, arc.p, seg.p, off, nl, bo, grp);
			// hits inherit the weakness of the arcs they sit on; weak_br == 2 filters (branch.c:108-145, graph.c:309)
			const bool resorted = fix_cm_tied_survivors(2 + 2 * branch_round_no);   // branch.c:116: one (cm, cs) pair later than the pg_gen_arc before it
			PGX_LAUNCH(k_branch_hits, GS_GRID, TB, 0, st, sv.p, dc.p, A.ctg, A.gidp, A.bits, g2s.p, arc.p, (const int32_t*)((g_comm.active() || resorted)? nullptr : arc_of_rec.p), counters.p + C_BR_CHANGED);
		}
		scratch.reset(mk);
	}

	// The hit at the front of the reference's array is never un-shadowed by pg_shadow (overlap.c:108).  Where a genome's first
	// (contig, cs) is shared by several hits, the reference's sort chain -- one unstable radix sort by cs at load time, then a
	// (cm, cs) pair per pg_gen_arc and per pg_mark_branch_flt_hit -- decides which of them is in front at each sweep.  The chain
	// depends on the contig's cs/cm sequences only, so it is replayed here, once per run, for the first contig of the flagged
	// genomes: target[m][k] = F index of the front hit of flagged genome k at the sweep of the m-th pg_gen_arc call.
	void prepare_first_replay(const pg_opt_t *opt)
	{
		n_first = 0; arc_call = 0;
		if (G == 0 || N == 0) return;
		if (read_u32(counters.p + C_FIRST_TIE) == 0) return;
		std::vector<FirstGroup> fg(G);
		PGX_CHECK(cudaMemcpyAsync(fg.data(), first_grp.p, (size_t)G * sizeof(FirstGroup), cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		std::vector<int32_t> flagged;
		for (int32_t g = 0; g < G; ++g) if (fg[g].m >= 2) flagged.push_back(g);
		if (flagged.empty()) return;
		const int32_t nf = (int32_t)flagged.size(), n_calls = 2 + (opt->n_branch_flt > 0? opt->n_branch_flt : 0);
		std::vector<std::vector<int32_t>> h_cs(nf), h_cm(nf);
		std::vector<std::vector<uint32_t>> h_src(nf);
		for (int32_t k = 0; k < nf; ++k) {
			const FirstGroup &r = fg[flagged[k]];
			h_cs[k].resize(r.len); h_cm[k].resize(r.len); h_src[k].resize(r.len);
			PGX_CHECK(cudaMemcpyAsync(h_cs[k].data(), F.cs + r.i0, (size_t)r.len * 4, cudaMemcpyDeviceToHost, st));
			PGX_CHECK(cudaMemcpyAsync(h_cm[k].data(), F.cm + r.i0, (size_t)r.len * 4, cudaMemcpyDeviceToHost, st));
			PGX_CHECK(cudaMemcpyAsync(h_src[k].data(), src.p + r.i0, (size_t)r.len * 4, cudaMemcpyDeviceToHost, st));
		}
		PGX_CHECK(cudaStreamSynchronize(st));
		// (cm, cs) sort pairs the reference has done when the m-th pg_gen_arc runs its sweep: 0, 1, then 3 + 2i in branch round i
		std::vector<int32_t> pairs_at(n_calls);
		for (int32_t m = 0; m < n_calls; ++m) pairs_at[m] = m < 2? m : 3 + 2 * (m - 2);
		std::vector<int32_t> target((size_t)n_calls * nf);
		ThreadPool::get().parallel_for(nf, [&](int64_t k) {
			const FirstGroup &r = fg[flagged[k]];
			const int32_t n = r.len;
			std::vector<pg128_t> a(n);
			std::vector<int32_t> ord(n);
			for (int32_t i = 0; i < n; ++i) ord[i] = i;
			std::sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) { return h_src[k][x] < h_src[k][y]; });   // the contig's hits in file order
			for (int32_t i = 0; i < n; ++i) { a[i].y = (uint64_t)ord[i]; a[i].x = (uint64_t)h_cs[k][ord[i]]; }
			FlagSortReplay::run(a.data(), n);                                                                       // read.c:247
			int32_t m = 0;
			for (int32_t p = 0;; ++p) {
				while (m < n_calls && pairs_at[m] == p) target[(size_t)m++ * nf + k] = r.i0 + (int32_t)a[0].y;
				if (m == n_calls) break;
				for (int32_t i = 0; i < n; ++i) a[i].x = (uint64_t)h_cm[k][a[i].y];
				FlagSortReplay::run(a.data(), n);                                                                   // graph.c:103, branch.c:116
				for (int32_t i = 0; i < n; ++i) a[i].x = (uint64_t)h_cs[k][a[i].y];
				FlagSortReplay::run(a.data(), n);                                                                   // graph.c:123, branch.c:140
			}
		});
		first_flagged.alloc(nf); first_target.alloc(target.size());
		PGX_CHECK(cudaMemcpyAsync(first_flagged.p, flagged.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaMemcpyAsync(first_target.p, target.data(), target.size() * 4, cudaMemcpyHostToDevice, st));
		PGX_CHECK(cudaStreamSynchronize(st));   // `flagged` and `target` go out of scope
		n_first = nf;
		if (pg_verbose >= 3) fprintf(stderr, "[M::pgx::first_replay] %d genome(s) start with hits sharing one (contig, cs): sort chain replayed for their first contig\n", nf);
	}
	// called at the top of every gen_arc: moves HB_FIRST for the sweep of this call
	void set_first_for_this_call()
	{
		const int32_t m = arc_call++;
		if (n_first == 0 || Nw == 0) return;
		PGX_LAUNCH(k_set_first, grid_for(n_first, TB), TB, 0, st, first_flagged.p, first_target.p + (size_t)m * n_first, n_first, first_grp.p, goff_act.p, A.full, A.bits);
	}


	// ------------------------------------------------------------------------------------------------
	// the sync-free rounds (pgx_round.cuh): buffers allocated once per run, pointers stable across rounds
	// ------------------------------------------------------------------------------------------------
	struct Fast {
		bool on = false, use_dirty = false;
		int32_t NV = 0, tiles_hit = 0, tiles_vtx = 0, tiles_seg = 0;
		uint32_t H = 0, H2 = 0, arc_cap = 0, nl_cap = 0, part_cap = 0;
		size_t bm_words = 0, lb_bytes = 0, seen_words = 0, dirty_words = 0;
		DBuf<unsigned long long> lkey, gkey, l2key, l2dsum, lb, nl_total;
		DBuf<ArcAcc> lacc, gacc;
		DBuf<uint32_t> llist, glist, bm, l2n, l2list, degcur, off, nl_off, seg_seen, spos, svcs, tie_list, dirty, rec, pos_of, seg_pos;
		DBuf<int32_t> l2m, segocc, t_last, grp;
		DBuf<SvInfo> info; DBuf<uint8_t> nl; DBuf<GenePos> gpos; DBuf<ArcDev> arc_tmp; DBuf<ArcWire> all; DBuf<int2> reach;
		ArcTable loc, glob; OvfTable l2;
		unsigned long long *st_cm = nullptr, *st_vs = nullptr, *st_cs = nullptr, *st_sg = nullptr; uint32_t *tk_cm = nullptr, *tk_vs = nullptr, *tk_cs = nullptr, *tk_sg = nullptr, *n_keep = nullptr;
	} fast;
	bool force_careful = false;
	std::vector<int64_t> rank_N;          // hits of every rank (multi-GPU)

	void clear_launch(const ClearList &c) { if (c.n) PGX_LAUNCH(k_round_clear, 148 * 4, 256, 0, st, c); }
	DirtyCols dirty_cols() const { DirtyCols D; D.reach = fast.reach.p; D.tiles = fast.use_dirty? fast.dirty.p : nullptr; return D; }
	ArcSrc arc_src(const pg_opt_t *opt) const
	{
		ArcSrc e; e.sv = sv.p; e.ctg = A.ctg; e.cm = A.cm; e.gidp = A.gidp; e.sori = A.sori; e.sdom = A.sdom; e.pdom0 = A.pdom0; e.bits = A.bits;
		e.g2s = g2s.p; e.prot_gid = prot_gid.p; e.ctg2g = ctg2g.p; e.use_ori = !!(opt->flag & PG_F_ORI_FOR_BRANCH); e.n_seg0 = n_seg0; e.Gl = G > 0? G : 1;
		return e;
	}

	// Allocates the tables of the fast rounds before the first pg_gen_arc, sized from the vertex count (the arc count is not known yet;
	// a graph with more than ~14 arcs per vertex side overflows the table, raises the fallback and is redone on the careful path).
	// Every decision here depends on replicated values only, so all ranks decide alike.
	bool setup_fast_tables(const pg_opt_t *opt)
	{
		fast.on = false;
		static const bool env_careful = []() { const char *e = getenv("PGX_ROUNDS"); return e && strcmp(e, "careful") == 0; }();
		if (env_careful || force_careful || n_seg0 == 0) return false;
		// the packed sums of ArcAcc: counts below 2^26, score sums below 2^38
		int64_t n_hits_all = 0, s_max = 0;
		for (int64_t x : rank_N) n_hits_all += x;
		for (int32_t x : h_prot_max) s_max = x > s_max? x : s_max;
		if (n_hits_all >= (1ll << 26) || Gtot >= (1 << 26) || (double)s_max * (double)Gtot >= (double)(1ll << ACC_SHIFT)) return false;
		Fast &f = fast;
		f.NV = 2 * n_seg0;
		uint64_t want = 28ull * n_seg0; if (want < (1u << 16)) want = 1u << 16;          // canonical keys at load factor <= 1/2: room for 14 arcs per oriented vertex
		uint32_t H = 1u << 16; while (H < want && H < (1u << 29)) H <<= 1;
		int Gmax = 0; for (int g : rank_G) Gmax = g > Gmax? g : Gmax;
		if ((uint64_t)H * (uint64_t)(Gmax > 0? Gmax : 1) / 8 > ((uint64_t)1 << 30)) return false;   // (key, genome) marks beyond 1 GB: stay on the sort path
		f.H = H; f.H2 = 1u << 20;
		f.bm_words = ((size_t)H * (size_t)(G > 0? G : 1) + 31) / 32;
		f.bm_words = (f.bm_words + 3) & ~(size_t)3;
		const bool multi = g_comm.active();
		f.lkey.ensure(H); f.lacc.ensure(H); f.llist.ensure(H); f.bm.ensure(2 * f.bm_words); f.pos_of.ensure(2 * (size_t)H);
		f.loc.key = f.lkey.p; f.loc.acc = f.lacc.p; f.loc.list = f.llist.p; f.loc.bm1 = f.bm.p; f.loc.bm2 = f.bm.p + f.bm_words; f.loc.H = H;
		f.glob = f.loc;
		if (multi) {
			f.gkey.ensure(H); f.gacc.ensure(H); f.glist.ensure(H);
			f.glob.key = f.gkey.p; f.glob.acc = f.gacc.p; f.glob.list = f.glist.p; f.glob.bm1 = f.glob.bm2 = nullptr;
			f.part_cap = H / 2;                                                            // rounds 1-2: whatever the table can hold; shrunk to the measured size afterwards
			f.all.ensure((size_t)g_comm.nranks * (f.part_cap + 1));
		}
		f.l2key.ensure(f.H2); f.l2dsum.ensure(f.H2); f.l2n.ensure(f.H2); f.l2list.ensure(f.H2); f.l2m.ensure(2 * (size_t)f.H2);
		f.l2.key = f.l2key.p; f.l2.dsum = f.l2dsum.p; f.l2.n = f.l2n.p; f.l2.list = f.l2list.p; f.l2.m1 = f.l2m.p; f.l2.m2 = f.l2m.p + f.H2; f.l2.H = f.H2;
		PGX_CHECK(cudaMemsetAsync(f.l2key.p, 0xff, (size_t)f.H2 * 8, st)); PGX_CHECK(cudaMemsetAsync(f.l2dsum.p, 0, (size_t)f.H2 * 8, st));
		PGX_CHECK(cudaMemsetAsync(f.l2n.p, 0, (size_t)f.H2 * 4, st)); PGX_CHECK(cudaMemsetAsync(f.l2m.p, 0, (size_t)f.H2 * 8, st));
		const size_t nw1 = (size_t)Nw + 16, tab = (size_t)n_gene * (size_t)(G > 0? G : 1);
		f.rec.ensure(nw1); f.info.ensure(nw1); f.spos.ensure(nw1 + 2); f.svcs.ensure(nw1); f.tie_list.ensure(nw1); f.reach.ensure(nw1);
		f.degcur.ensure(2 * (size_t)f.NV + 8); f.off.ensure((size_t)f.NV + 8); f.nl_off.ensure((size_t)f.NV + 8); f.nl_total.ensure(2); f.seg_pos.ensure((size_t)n_seg0 + 8);
		f.seen_words = (((size_t)(G > 0? G : 1) * n_seg0 + 31) / 32 + 4) & ~(size_t)3;
		f.seg_seen.ensure(f.seen_words); f.segocc.ensure(2 * (size_t)n_seg0 + 4);
		f.t_last.ensure(tab + 4); f.gpos.ensure(tab + 1);
		f.dirty_words = ((size_t)Nw / SWQ_TILE / 32 + 8) & ~(size_t)3;
		f.dirty.ensure(f.dirty_words);
		f.tiles_hit = (int32_t)(((size_t)Nw + 1 + SCAN_TILE - 1) / SCAN_TILE); f.tiles_vtx = (int32_t)(((size_t)f.NV + 1 + SCAN_TILE - 1) / SCAN_TILE);
		f.tiles_seg = (int32_t)(((size_t)n_seg0 + 1 + SCAN_TILE - 1) / SCAN_TILE);
		const size_t lb_words = 2 * (size_t)f.tiles_hit + 2 * (size_t)f.tiles_vtx + (size_t)f.tiles_seg + 4;   // (sized for the active set as it is now: it only shrinks)
		f.lb.ensure(lb_words + 2); f.lb_bytes = ((lb_words * 8) + 15) & ~(size_t)15;
		f.st_cm = f.lb.p; f.st_cs = f.st_cm + f.tiles_hit; f.st_vs = f.st_cs + f.tiles_hit; f.st_sg = f.st_vs + 2 * (size_t)f.tiles_vtx;
		f.tk_cm = (uint32_t*)(f.st_sg + f.tiles_seg); f.tk_vs = f.tk_cm + 1; f.tk_cs = f.tk_cm + 2; f.tk_sg = f.tk_cm + 3; f.n_keep = f.tk_cm + 4;
		if (arc.n < (size_t)H + 1) { arc.alloc((size_t)H + 1); arc_alt.alloc((size_t)H + 1); arc_cap = arc.n - 1; }   // up to H / 2 canonical keys -> H arcs
		f.arc_cap = (uint32_t)(arc.n - 1);
		f.arc_tmp.ensure((size_t)f.arc_cap + 1); f.grp.ensure((size_t)f.arc_cap + 1);
		f.use_dirty = false;                                                              // the first sweeps are complete ones; setup_fast_pairs() turns the incremental mode on
		f.nl_cap = 0;
		PGX_CHECK(cudaMemsetAsync(f.lb.p, 0, f.lb_bytes, st));
		PGX_LAUNCH(k_set_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, fallback)), 0u);
		fast.on = true;
		return true;
	}
	// After the round-2 graph: the size of the locality table of the branch rounds (one read-back per run), the exchange buffer of a
	// sharded run cut down to the measured graph, the reach of every active hit for the incremental sweeps.
	bool setup_fast_pairs(const pg_opt_t *opt)
	{
		Fast &f = fast;
		if (opt->n_branch_flt <= 0) return true;
		struct { unsigned long long nl; DevCounts c; } h;
		PGX_CHECK(cudaMemcpyAsync(&h.nl, f.nl_total.p, 8, cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaMemcpyAsync(&h.c, dc.p, sizeof(DevCounts), cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		if (g_comm.active()) {   // (the fallback word may differ between ranks at this point; the arc counts are replicated)
			g_comm.allreduce(dcp(offsetof(DevCounts, fallback)), 1, Comm::T_UINT32, Comm::OP_MAX, st);
			h.c.fallback = read_counts().fallback;
		}
		if (h.c.fallback || h.nl > (1ull << 29)) return false;                           // overflow already, or half a billion arc pairs: not the regime this path is for
		f.nl_cap = (uint32_t)(2 * h.nl + (1u << 20));
		f.nl.ensure((size_t)f.nl_cap + 16);
		PGX_CHECK(cudaMemsetAsync(f.nl.p, 0, f.nl.n, st));
		if (g_comm.active()) {   // the exchange buffer: the largest per-rank table of round 2 plus a quarter (the graph only loses arcs from here on)
			uint32_t *w = (uint32_t*)(f.nl_total.p + 1);
			PGX_LAUNCH(k_copy_u32, 1, 1, 0, st, w, dcp(offsetof(DevCounts, n_part)));
			g_comm.allreduce(w, 1, Comm::T_UINT32, Comm::OP_MAX, st);
			const uint32_t mx = read_u32(w), c = mx + mx / 4 + 1024;
			if (c < f.part_cap) f.part_cap = c;
		}
		f.use_dirty = n_first == 0 && !(getenv("PGX_SWEEP") && strcmp(getenv("PGX_SWEEP"), "global") == 0) && !getenv("PGX_NO_DIRTY");
		PGX_CHECK(cudaMemsetAsync(f.dirty.p, 0, f.dirty_words * 4, st));
		if (f.use_dirty && Nw) PGX_LAUNCH(k_reach, GS_GRID, TB, 0, st, A.ctg, A.cs, A.ce, A.pmax, Nw, (int32_t)SWQ_TILE, f.reach.p);
		return true;
	}

	// graph.c:87-177 without sorting the records and without a host round trip (pgx_round.cuh)
	void gen_arc_fast(const pg_opt_t *opt, bool branch_round_follows = false)
	{
		Fast &f = fast;
		set_first_for_this_call();
		const bool multi = g_comm.active();
		uint32_t *deg = f.degcur.p, *cur = f.degcur.p + f.NV;
		ClearList cl; cl.n = 0;
		clear_add(cl, f.lb.p, f.lb_bytes, 0u);
		clear_add(cl, f.segocc.p, 2 * (size_t)n_seg0 * 4, 0u);
		clear_add(cl, f.seg_seen.p, f.seen_words * 4, 0u);
		clear_add(cl, f.lkey.p, (size_t)f.H * 8, 0xffffffffu);
		clear_add(cl, f.lacc.p, (size_t)f.H * sizeof(ArcAcc), 0u);
		clear_add(cl, f.bm.p, 2 * f.bm_words * 4, 0u);
		clear_add(cl, f.degcur.p, 2 * (size_t)f.NV * 4, 0u);
		clear_add(cl, idx.p, (size_t)n_seg0 * 16, 0u);
		clear_add(cl, f.nl_total.p, 16, 0u);
		if (multi) { clear_add(cl, f.gkey.p, (size_t)f.H * 8, 0xffffffffu); clear_add(cl, f.gacc.p, (size_t)f.H * sizeof(ArcAcc), 0u); }
		clear_launch(cl);
		pt.mark("arc:clear", st);
		SweepOpt so; so.min_ov_ratio = opt->min_ov_ratio; so.check_strand = !!(opt->flag & PG_F_CHECK_STRAND);
		launch_sweep<SW_SHADOW>(sweep_active(), so, A.bits, A.pdom, nullptr, f.use_dirty? f.dirty.p : nullptr);   // in place: a sweep reads flt / weak_br / rev of its partners, writes its own shadow bit
		pt.mark("arc:sweep", st);
		if (branch_round_follows) {   // fork: the position tables of that round, on the second stream
			PGX_CHECK(cudaEventRecord(ev_swept, st));
			PGX_CHECK(cudaStreamWaitEvent(st2, ev_swept, 0));
			branch_positions(st2);
			PGX_CHECK(cudaEventRecord(ev_pos, st2));
			pos_ready = true;
		}
		const int32_t tiles_now = (int32_t)(((size_t)Nw + 1 + SCAN_TILE - 1) / SCAN_TILE);   // <= f.tiles_hit: the active set only shrinks
		PGX_LAUNCH(k_compact_survivors, tiles_now, SCAN_B, 0, st, A.bits, by_cm.p, Nw, (uint32_t*)nullptr, sv.p, dcp(offsetof(DevCounts, m_sv)), f.st_cm, f.tk_cm, dc.p);
		const ArcSrc e = arc_src(opt);
		int32_t *seg_ng_f = f.segocc.p, *seg_tot_f = f.segocc.p + n_seg0;
		PGX_LAUNCH(k_survivor_info, GS_GRID, TB, 0, st, e, dc.p, f.info.p, f.seg_seen.p, seg_ng_f, seg_tot_f, counters.p + C_ERR);
		PGX_LAUNCH(k_arc_insert, GS_GRID, TB, 0, st, f.info.p, ctg2g.p, e.Gl, f.loc, dc.p, f.rec.p, multi? (uint32_t*)nullptr : deg, counters.p + C_TIE_CM);
		PGX_LAUNCH(k_arc_fold, GS_GRID, TB, 0, st, f.info.p, f.rec.p, ctg2g.p, f.loc, dc.p, e.Gl, f.l2);
		PGX_LAUNCH(k_arc_ovf_fold, 148, TB, 0, st, f.l2, f.loc, dc.p);
		pt.mark("arc:survivors+table", st);
		if (multi) {
			ArcWire *mine = f.all.p + (size_t)g_comm.rank * (f.part_cap + 1);
			PGX_LAUNCH(k_arc_pack, GS_GRID, TB, 0, st, f.loc, dc.p, mine, f.part_cap);
			g_comm.group_begin();
			g_comm.allgather(mine, f.all.p, (size_t)(f.part_cap + 1) * sizeof(ArcWire), st);
			g_comm.allreduce(f.segocc.p, 2 * (size_t)n_seg0, Comm::T_INT32, Comm::OP_SUM, st);
			g_comm.group_end();
			PGX_LAUNCH(k_arc_merge, GS_GRID, TB, 0, st, f.all.p, g_comm.nranks, f.part_cap, f.glob, dc.p, deg);
			pt.mark("arc:exchange+merge", st);
		}
		PGX_LAUNCH(k_vertex_scan, f.tiles_vtx, SCAN_B, 0, st, deg, f.NV, f.off.p, f.nl_off.p, dc.p, f.arc_cap, f.nl_total.p, f.st_vs, f.tk_vs);
		PGX_LAUNCH(k_arc_place, GS_GRID, TB, 0, st, f.glob, dcp(multi? offsetof(DevCounts, n_gpart) : offsetof(DevCounts, n_part)), f.off.p, cur, dc.p, f.arc_tmp.p);
		PGX_LAUNCH(k_arc_rank, GS_GRID, TB, 0, st, f.arc_tmp.p, f.off.p, dc.p, arc.p, idx.p, f.pos_of.p, seg.p, seg_ng_f, seg_tot_f);   // (+ graph.c:125-126: the occupancy of this build)
		pt.mark("arc:scan+place+rank", st);
		++g_stats.fast_rounds;
	}

	// branch.c:6-29: the representative position of every gene in every local genome, from the flag words as the last sweep left them;
	// plus the detector for survivors that share one (contig, cs) (DESIGN.md "tie order").  Clears the dirty-tile marks: the sweep that
	// consumed them is done, the next ones come from k_branch_hits_fast.
	void branch_positions(cudaStream_t s)
	{
		Fast &f = fast;
		const size_t tab = (size_t)n_gene * G;
		ClearList cl; cl.n = 0;
		clear_add(cl, f.t_last.p, (tab + 1) * 4, 0xffffffffu);
		clear_add(cl, counters.p + C_TIE_ROUND, 16, 0u);
		clear_add(cl, f.dirty.p, f.dirty_words * 4, 0u);
		PGX_LAUNCH(k_round_clear, 148 * 4, 256, 0, s, cl);
		const int32_t tiles_now = (int32_t)(((size_t)Nw + 1 + SCAN_TILE - 1) / SCAN_TILE);
		PGX_LAUNCH(k_compact_survivors, tiles_now, SCAN_B, 0, s, A.bits, (const uint32_t*)nullptr, Nw, f.spos.p, f.svcs.p, dcp(offsetof(DevCounts, m_cs)), f.st_cs, f.tk_cs, (DevCounts*)nullptr);
		if (Nw) PGX_LAUNCH(k_pos_last, grid_for(Nw, TB), TB, 0, s, A.ctg, ctg2g.p, A.gidp, f.spos.p, Nw, G, f.t_last.p);
		if (tab) PGX_LAUNCH(k_pos_fill, grid_for((int64_t)tab, TB), TB, 0, s, f.t_last.p, tab, G, A.ctg, A.cm, f.spos.p, goff_act.p, f.gpos.p);
		PGX_LAUNCH(k_cs_tie_mark, GS_GRID, TB, 0, s, A.ctg, A.cs, A.gidp, ctg2g.p, G, f.svcs.p, dc.p, f.t_last.p, f.gpos.p, counters.p + C_TIE_CS, counters.p + C_TIE_ROUND, f.tie_list.p, counters.p + C_LONG_RUN);
	}

	// branch.c:48-145 + graph.c:309-313, same rules as branch_round() below; idx / nl_off / the key table come from the preceding arc build
	void branch_round_fast(const pg_opt_t *opt, int32_t round, int32_t max_avg_occ, int32_t max_degree, int32_t max_dist_loci)
	{
		Fast &f = fast;
		if (pos_ready) { PGX_CHECK(cudaStreamWaitEvent(st, ev_pos, 0)); pos_ready = false; }   // join: built beside the preceding pg_gen_arc
		else branch_positions(st);
		pt.mark("br:positions", st);
		BranchOpt bo; bo.local_dist = opt->local_dist; bo.local_count = opt->local_count; bo.frag_mode = !!(opt->flag & PG_F_FRAG_MODE);
		bo.branch_diff = opt->branch_diff; bo.branch_diff_dist = opt->branch_diff_dist; bo.branch_diff_cut = opt->branch_diff_cut;
		PGX_LAUNCH(k_branch_local, GS_GRID, TB, 0, st, idx.p, dc.p, arc.p, seg.p, f.gpos.p, G, f.nl_off.p, f.NV, f.nl_cap, f.nl_total.p, bo, f.nl.p, counters.p + C_ORD_SENS);
		pt.mark("br:locality", st);
		g_comm.allreduce(f.nl.p, f.nl_cap, Comm::T_UINT8, Comm::OP_SUM, st);
		if (g_comm.active()) pt.mark("br:allreduce", st);
		PGX_LAUNCH(k_branch_decide_fast, GS_GRID, 128, 0, st, idx.p, dc.p, arc.p, seg.p, f.nl_off.p, f.nl.p, f.nl_cap, f.nl_total.p, bo, f.grp.p, counters.p + C_TIE_ROUND);
		PGX_LAUNCH(k_branch_hits_fast, GS_GRID, TB, 0, st, sv.p, f.info.p, dc.p, A.bits, arc.p, f.glob, f.pos_of.p, g_comm.active()? (const uint32_t*)nullptr : (const uint32_t*)f.rec.p, dirty_cols(), counters.p + C_BR_CHANGED);
		pt.mark("br:decide+hits", st);
		if (round > 0) flt_high_occ_fast(max_avg_occ, max_degree, max_dist_loci);          // graph.c:310-313
	}
	// graph.c:237-263 + PG_SET_FILTER(vtx == 0); needs idx[] of the current arcs and a zeroed look-back block (every gen_arc_fast leaves both)
	void flt_high_occ_fast(int32_t max_avg_occ, int32_t max_degree, int32_t max_dist_loci)
	{
		Fast &f = fast;
		PGX_CHECK(cudaMemsetAsync(g2s.p, 0xff, (size_t)n_gene * 4, st));
		PGX_LAUNCH(k_seg_keep_scan, f.tiles_seg, SCAN_B, 0, st, seg.p, dc.p, idx.p, max_avg_occ * Gtot, max_degree, max_dist_loci, f.seg_pos.p, f.n_keep, f.st_sg, f.tk_sg);
		PGX_LAUNCH(k_seg_compact_fast, 148, TB, 0, st, seg.p, seg_alt.p, dc.p, f.seg_pos.p, f.n_keep, g2s.p);
		seg.swap(seg_alt);
		if (Nw) PGX_LAUNCH(k_flag_vtx_fast, grid_for(Nw, TB), TB, 0, st, A.gidp, A.bits, Nw, g2s.p, dirty_cols(), dc.p, f.n_keep);   // (also stores the new segment count)
		else PGX_LAUNCH(k_copy_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, n_seg)), f.n_keep);
		pt.mark("br:segfilter+vtx", st);
	}

	void stage_graph(const pg_opt_t *opt)
	{
		if (!done_post) PGX_FATAL("pg_graph_gen() called before pg_post_process()");
		if (stage_graph_impl(opt, true)) return;
		// a fixed-capacity table of the fast rounds overflowed, or a round needs the reference's order of cs-tied survivors: the same run again
		// from the raw hits, every round on the careful path (host-sized sorts, per-round read-backs, chain replay)
		if (pg_verbose >= 3) fprintf(stderr, "[M::pgx::graph_gen] redoing the run on the careful path\n");
		++g_stats.redo;
		PGX_CHECK(cudaMemsetAsync(counters.p, 0, C_N * 4, st));
		stage_load(opt);
		stage_post(opt);
		if (!stage_graph_impl(opt, false)) PGX_FATAL("careful path asked for a redo");
	}
	bool stage_graph_impl(const pg_opt_t *opt, bool allow_fast)
	{
		pt.begin(st);
		PGX_CHECK(cudaStreamSynchronize(st2)); pos_ready = false;
		set_filter(SF_PSEUDO);                                                           // graph.c:285
		gen_vtx(opt);                                                                    // graph.c:286
		pt.mark("vertex election", st);
		prepare_first_replay(opt);
		tie_chains.clear();
		flag_vtx();                                                                      // graph.c:287-288
		recompact_active();
		pt.mark("recompact", st);
		force_careful = !allow_fast;
		bool use_fast = setup_fast_tables(opt);
		pt.mark("fast-round setup", st);
		if (use_fast) {
			gen_arc_fast(opt);                                                             // graph.c:289
			flt_high_occ_fast(opt->max_avg_occ * 2, opt->max_degree * 2, opt->max_dist_loci);   // graph.c:294-295
			recompact_active();
			pt.mark("recompact", st);
			gen_arc_fast(opt, opt->n_branch_flt > 0);                                      // graph.c:296
			if (!setup_fast_pairs(opt)) { fast.on = false; PGX_CHECK(cudaStreamSynchronize(st2)); pos_ready = false; return false; }
			pt.mark("fast-round setup", st);
		} else {
			gen_arc(opt);                                                                  // graph.c:289
			pt.mark("arcs round 1-2 (sort path)", st);
			flt_high_occ(opt->max_avg_occ * 2, opt->max_degree * 2, opt->max_dist_loci);   // graph.c:294-295
			recompact_active();
			pt.mark("recompact", st);
			gen_arc(opt);                                                                  // graph.c:296
			pt.mark("arcs round 1-2 (sort path)", st);
		}
		for (int32_t i = 0; i < opt->n_branch_flt; ++i) {                                // graph.c:301-315
			double r = 1.0 + (double)(opt->n_branch_flt - 1 - i) / opt->n_branch_flt;
			int32_t max_avg_occ = (int32_t)(opt->max_avg_occ * r + .499);
			int32_t max_degree = (int32_t)(opt->max_degree * r + .499);
			int32_t max_dist_loci = (int32_t)(opt->max_dist_loci * r + .499);
			static const bool trace = getenv("PGX_TRACE") != nullptr;   // (per-round read-backs: they stall the pipeline, keep them out of timing runs)
			uint32_t seg_before = 0;
			if (trace) { seg_before = read_counts().n_seg; PGX_CHECK(cudaMemsetAsync(counters.p + C_BR_CHANGED, 0, 4, st)); }
			branch_round_no = i;
			if (use_fast) branch_round_fast(opt, i, max_avg_occ, max_degree, max_dist_loci);
			else {
				branch_round(opt);
				if (i > 0) flt_high_occ(max_avg_occ, max_degree, max_dist_loci);
			}
			if (trace) {   // how much does a round change?
				const uint32_t chg = read_u32(counters.p + C_BR_CHANGED);
				DevCounts hc = read_counts();
				fprintf(stderr, "[pgx trace] branch round %d: %u hit words changed, segments %u -> %u, arcs %u, survivors %u, arc pairs %u, overflow records %u, fallback %u\n", i, chg, seg_before, hc.n_seg, hc.n_arc, hc.m_sv, hc.n_nl, hc.n_ovf, hc.fallback);
			}
			if (use_fast) gen_arc_fast(opt, i + 1 < opt->n_branch_flt); else { gen_arc(opt); pt.mark("careful rounds", st); }
		}
		if (use_fast) {
			if (g_comm.active()) g_comm.allreduce(dcp(offsetof(DevCounts, fallback)), 1, Comm::T_UINT32, Comm::OP_MAX, st);
			if (read_counts().fallback) { fast.on = false; PGX_CHECK(cudaStreamSynchronize(st2)); pos_ready = false; return false; }
		}
		set_filter(SF_SHADOW);                                                           // graph.c:316
		DevCounts hc = read_counts();
		n_seg = (int32_t)hc.n_seg; n_arc = (int32_t)hc.n_arc;
		if (opt->min_arc_cnt > 1 && n_arc) {                                             // graph.c:317-318
			size_t mk = scratch.mark();
			uint32_t *keep = scratch.get<uint32_t>((size_t)n_arc + 1), *pos = scratch.get<uint32_t>((size_t)n_arc + 2);
			PGX_LAUNCH(k_arc_keep_min_cnt, grid_for(n_arc, TB), TB, 0, st, arc.p, n_arc, opt->min_arc_cnt, keep);
			uint32_t k = scan_flags(keep, pos, n_arc);
			PGX_LAUNCH(k_arc_compact, grid_for(n_arc, TB), TB, 0, st, arc.p, n_arc, pos, arc_alt.p);
			arc.swap(arc_alt);
			n_arc = (int32_t)k;
			PGX_LAUNCH(k_set_u32, 1, 1, 0, st, dcp(offsetof(DevCounts, n_arc)), (uint32_t)n_arc);
			scratch.reset(mk);
		}
		arc_index();                                                                     // graph.c:319
		check_err("graph generation");
		if (pg_verbose >= 3) fprintf(stderr, "[M::pgx::graph_gen] final graph: %d genes and %d arcs\n", n_seg, n_arc);
		done_graph = true;
		g_stats.n_seg = n_seg; g_stats.n_arc = n_arc;
		pt.mark("tail", st);
		pt.report("stage 2", g_comm.rank, st);
		if (getenv("PGX_TIMING")) fprintf(stderr, "[pgx timing] rank %d stage 2 host wall: arc rounds %.2f ms (of which cross-rank merge %.2f), branch rounds %.2f ms\n", g_comm.rank, t_arc * 1e3, t_merge * 1e3, t_branch * 1e3);
		t_arc = t_merge = t_branch = 0;
		return true;
	}

	void collect_stats()
	{
		uint32_t h[C_N];
		PGX_CHECK(cudaMemcpyAsync(h, counters.p, sizeof(h), cudaMemcpyDeviceToHost, st));
		PGX_CHECK(cudaStreamSynchronize(st));
		g_stats.tie[0] = h[C_TIE_CM]; g_stats.tie[1] = h[C_TIE_CS]; g_stats.tie[2] = h[C_TIE_DUP]; g_stats.tie[3] = h[C_TIE_SUB];
		float ms = 0;
		if (N) { PGX_CHECK(cudaEventElapsedTime(&ms, ev[0], ev[1])); }
		g_stats.sweep_ms = ms;
	}

	// ------------------------------------------------------------------------------------------------
	// results back into the caller's structs (malloc'd, as the reference's destroy functions expect)
	// ------------------------------------------------------------------------------------------------
	void download(pg_data_t *d, pg_graph_t *q, bool with_hits = true)
	{
		if (d->n_genome != G) PGX_FATAL("download: pg_data_t does not match the uploaded one");
		ensure_rep();
		if (with_hits && packed_input) PGX_FATAL("the hits of this run came as device-ready records (pgx_post_process_packed): there are no pg_hit_t records to rebuild -- use pgx_walk_tokens()");
		if (Nw && with_hits) PGX_LAUNCH(k_scatter_back, grid_for(Nw, TB), TB, 0, st, A, Nw, F);
		if (N && with_hits) {
			DBuf<OutHit> out;
			out.alloc((size_t)N + 1);
			PGX_LAUNCH(k_pack_out, grid_for(N, TB), TB, 0, st, src.p, N, F, g2s.p, prot_rep.p, done_graph? 1 : 0, out.p);
			ThreadPool &pool = ThreadPool::get();
			SpareHitArrays &spare = SpareHitArrays::get();
			std::vector<pg_hit_t*> fresh(G, nullptr);
			std::vector<size_t> fresh_cap(G, 0);
			// second download of a strict-mode run (after pg_graph_gen): the caller still holds the arrays we returned after
			// pg_post_process, already permuted -> patch them in place.  Anything else is rebuilt from raw order.
			bool host_sorted = (int32_t)installed.size() == G && G > 0;
			for (int32_t j = 0; j < G && host_sorted; ++j) host_sorted = d->genome[j].hit == installed[j];
			for (int32_t j = 0; j < G; ++j)
				if (h_hoff[j + 1] - h_hoff[j] != d->genome[j].n_hit) PGX_FATAL("download: genome %d changed size since upload", j);
			const int64_t BLK = 1 << 14;
			const bool in_place = host_sorted;
			int32_t g_next = 0;                       // genomes below g_next have their destination array
			int32_t g_done = 0;                       // genomes below g_done are complete: new array installed, old one handed to the pool
			// only the 24 bytes per hit that changed cross the bus; host threads rebuild the 88-byte records from the caller's own copy.
			// Destination arrays are taken as the chunks reach their genomes, and a genome's old array goes to the pool as soon as its last
			// chunk is done -- from the second chunk on the destinations are those recycled (mapped) arrays.
			staged_d2h(out.p, (size_t)N * sizeof(OutHit), sizeof(OutHit), [&](size_t off, size_t len, const char *srcbuf) {
				const int64_t h0 = (int64_t)(off / sizeof(OutHit)), h1 = h0 + (int64_t)(len / sizeof(OutHit));
				if (!in_place) for (; g_next < G && h_hoff[g_next] < h1; ++g_next) fresh[g_next] = spare.take((size_t)d->genome[g_next].n_hit, &fresh_cap[g_next]);
				pool.parallel_for((h1 - h0 + BLK - 1) / BLK, [&](int64_t b) {
					const int64_t a = h0 + b * BLK, e = std::min(h1, a + BLK);
					int g = (int)(std::upper_bound(h_hoff.begin(), h_hoff.end(), (int32_t)a) - h_hoff.begin()) - 1;
					const OutHit *o = (const OutHit*)srcbuf + (a - h0);
					for (int64_t t = a; t < e; ++t, ++o) {
						while (t >= h_hoff[g + 1]) ++g;
						const int64_t base = h_hoff[g];
						pg_hit_t *dst = in_place? &d->genome[g].hit[t - base] : &fresh[g][t - base];
						if (!in_place && t + 12 < e) {   // the source records are visited in random order: ask for the one needed 12 hits from now
							const char *pf = (const char*)&d->genome[g].hit[(int64_t)o[12].src - base];   // (past a genome boundary: a harmless wrong guess)
							__builtin_prefetch(pf); __builtin_prefetch(pf + 64);
						}
						if (!in_place) *dst = d->genome[g].hit[(int64_t)o->src - base];
						dst->rank = o->rank; dst->score_dom = o->sdom; dst->pid_dom = o->pdom; dst->pid_dom0 = o->pdom0;
						memcpy((char*)dst + PG_HIT_BITS_OFFSET, &o->bits, 4);
					}
				});
				if (!in_place) for (; g_done < G && h_hoff[g_done + 1] <= h1; ++g_done) {   // hit.c:62-63: the array in the old order goes, the new one takes its place
					pg_genome_t *gn = &d->genome[g_done];
					spare.give(gn->hit);
					gn->hit = fresh[g_done]; gn->m_hit = (int32_t)std::min<size_t>(fresh_cap[g_done], 0x7fffffff);
				}
			}, st);
			if (!host_sorted) {
				for (; g_done < G; ++g_done) {            // (empty genomes at the end of the list)
					pg_genome_t *gn = &d->genome[g_done];
					if (!fresh[g_done]) fresh[g_done] = spare.take((size_t)gn->n_hit, &fresh_cap[g_done]);
					spare.give(gn->hit);
					gn->hit = fresh[g_done]; gn->m_hit = (int32_t)std::min<size_t>(fresh_cap[g_done], 0x7fffffff);
				}
				installed.assign(fresh.begin(), fresh.end());
			}
			g_stats.d2h += (int64_t)N * (int64_t)sizeof(OutHit);
		}
		for (int32_t i = 0; i < n_prot; ++i) {   // hit.c:206-207,215,230-237
			d->prot[i].n = rep.prot_n[i]; d->prot[i].avg_score_adj = rep.prot_avg[i]; d->prot[i].rep = rep.prot_rep[i]; d->prot[i].max_score_ori = h_prot_max[i];
		}
		for (int32_t i = 0; i < n_gene; ++i) d->gene[i].rep_pid = rep.gene_rep_pid[i];
		if (q && done_graph) {
			free(q->seg); free(q->arc); free(q->g2s); free(q->idx);
			q->n_seg = q->m_seg = n_seg; q->n_arc = q->m_arc = n_arc;
			q->seg = (pg_seg_t*)calloc(n_seg > 0? n_seg : 1, sizeof(pg_seg_t));
			q->arc = (pg_arc_t*)calloc(n_arc > 0? n_arc : 1, sizeof(pg_arc_t));
			q->g2s = (int32_t*)malloc(sizeof(int32_t) * (n_gene > 0? n_gene : 1));
			q->idx = (uint64_t*)calloc((size_t)(n_seg > 0? n_seg : 1) * 2, sizeof(uint64_t));
			static_assert(sizeof(SegDev) == sizeof(pg_seg_t) && sizeof(ArcDev) == sizeof(pg_arc_t), "device mirrors of the graph records");
			if (n_seg) PGX_CHECK(cudaMemcpyAsync(q->seg, seg.p, (size_t)n_seg * sizeof(pg_seg_t), cudaMemcpyDeviceToHost, st));
			if (n_arc) PGX_CHECK(cudaMemcpyAsync(q->arc, arc.p, (size_t)n_arc * sizeof(pg_arc_t), cudaMemcpyDeviceToHost, st));
			PGX_CHECK(cudaMemcpyAsync(q->g2s, g2s.p, (size_t)n_gene * 4, cudaMemcpyDeviceToHost, st));
			if (n_seg) PGX_CHECK(cudaMemcpyAsync(q->idx, idx.p, (size_t)n_seg * 16, cudaMemcpyDeviceToHost, st));
			PGX_CHECK(cudaStreamSynchronize(st));
			g_stats.d2h += (int64_t)n_seg * 48 + (int64_t)n_arc * 32 + (int64_t)n_gene * 4;
		}
	}

	// ------------------------------------------------------------------------------------------------
	// what the W lines (format.c:183-225) and the presence/absence matrix (pangene.js:1168-1246) need, read off the device-resident
	// final state -- 12 bytes per printed hit instead of the 88-byte records of every hit.  tokens: malloc'd, genome-major, inside a genome contig-major,
	// inside a contig by cm (ties in cs order, as pgio's writer orders them); genome_off[G + 1].  Token contig ids are genome-local (pg_hit_t.cid).
	// ------------------------------------------------------------------------------------------------
	int64_t walk_tokens(WalkToken **tokens, int64_t *genome_off)
	{
		if (!done_graph) PGX_FATAL("walk tokens are read off the final state: call pg_graph_gen() first");
		*tokens = nullptr;
		for (int32_t j = 0; j <= G; ++j) genome_off[j] = 0;
		if (Nw == 0) { *tokens = (WalkToken*)malloc(sizeof(WalkToken)); return 0; }
		size_t mk = scratch.mark();
		uint32_t *flag = scratch.get<uint32_t>((size_t)Nw + 1), *pos = scratch.get<uint32_t>((size_t)Nw + 2);
		PGX_LAUNCH(k_flag_walk_cm, grid_for(Nw, TB), TB, 0, st, A.bits, by_cm.p, Nw, flag);
		const int64_t n_tok = (int64_t)scan_flags(flag, pos, Nw);
		WalkToken *h = (WalkToken*)malloc(sizeof(WalkToken) * (size_t)(n_tok > 0? n_tok : 1));
		if (n_tok) {
			WalkToken *dtok = scratch.get<WalkToken>((size_t)n_tok);
			PGX_LAUNCH(k_walk_tokens, grid_for(Nw, TB), TB, 0, st, pos, by_cm.p, Nw, A, src.p, raw.p, dtok);
			staged_d2h(dtok, (size_t)n_tok * sizeof(WalkToken), sizeof(WalkToken), [&](size_t off, size_t len, const char *srcbuf) { memcpy((char*)h + off, srcbuf, len); }, st);
			g_stats.d2h += n_tok * (int64_t)sizeof(WalkToken);
			// tokens are sorted by contig key: genome boundaries by binary search, then keys -> genome-local contig ids
			for (int32_t j = 0; j <= G; ++j)
				genome_off[j] = std::lower_bound(h, h + n_tok, h_coff[j], [](const WalkToken &t, int32_t key) { return t.ctg < key; }) - h;
			ThreadPool::get().parallel_for(G, [&](int64_t j) { for (int64_t t = genome_off[j]; t < genome_off[j + 1]; ++t) h[t].ctg -= h_coff[j]; });
		}
		scratch.reset(mk);
		*tokens = h;
		return n_tok;
	}
};

} // namespace pgx
