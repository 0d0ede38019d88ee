// pgx_api.cu -- the C-ABI surface of libpgx.so (include/pgx.h) over the device engine (pgx_engine.cuh).
// Built for sm_100a only; there is no CPU path behind any of these entry points.
#include <map>
#include <time.h>
#include <mutex>
#include "pgx_engine.cuh"

namespace pgx {
LaunchStats g_launch;
Stats g_stats;
Comm g_comm;
int g_sweep_stat = 0;
static int g_device = -1, g_lazy = 0;
static std::map<const pg_data_t*, Session*> g_sessions;   // pg_post_process() -> pg_graph_gen() hand-over
static std::mutex g_sessions_mu;

static double wall() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }
static bool timing_on() { static const bool on = getenv("PGX_TIMING") != nullptr; return on; }

static void need_device()
{
	static int checked_dev = -2;          // the property query is slow (tens of ms with several GPUs on the box): once per device
	if (g_device >= 0) PGX_CHECK(cudaSetDevice(g_device));
	{ int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess && cur == checked_dev) return; }
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n < 1) PGX_FATAL("no CUDA device available (%s). libpgx has no CPU fallback.", cudaGetErrorString(e));
	if (g_device >= 0) PGX_CHECK(cudaSetDevice(g_device));
	cudaDeviceProp p;
	int dev = 0;
	PGX_CHECK(cudaGetDevice(&dev));
	PGX_CHECK(cudaGetDeviceProperties(&p, dev));
	if (p.major < 10) PGX_FATAL("device %d (%s, sm_%d%d) is not a Blackwell part; libpgx is built for sm_100a only", dev, p.name, p.major, p.minor);
	checked_dev = dev;
}

static void timed_run(Session *s, const pg_opt_t *opt, int stage)
{
	cudaEvent_t e0 = s->ev[2], e1 = s->ev[3];
	float ms;
	if (stage & 1) {
		PGX_CHECK(cudaEventRecord(e0, s->st));
		PGX_CHECK(cudaMemsetAsync(s->counters.p, 0, Session::C_N * 4, s->st));
		s->stage_load(opt);
		s->stage_post(opt);
		PGX_CHECK(cudaEventRecord(e1, s->st));
		PGX_CHECK(cudaEventSynchronize(e1));
		PGX_CHECK(cudaEventElapsedTime(&ms, e0, e1));
		g_stats.run_ms[0] = ms;
	}
	if (stage & 2) {
		PGX_CHECK(cudaEventRecord(e0, s->st));
		s->stage_graph(opt);
		PGX_CHECK(cudaEventRecord(e1, s->st));
		PGX_CHECK(cudaEventSynchronize(e1));
		PGX_CHECK(cudaEventElapsedTime(&ms, e0, e1));
		g_stats.run_ms[1] = ms;
	}
	g_stats.run_ms[2] = ((stage & 1)? g_stats.run_ms[0] : 0) + ((stage & 2)? g_stats.run_ms[1] : 0);
	s->collect_stats();
}
} // namespace pgx

using namespace pgx;

extern "C" {

int pg_verbose = 3;

static void post_process_impl(const pg_opt_t *opt, pg_data_t *d, const pgx_packed_t *pk, const char *who);

void pg_post_process(const pg_opt_t *opt, pg_data_t *d) { post_process_impl(opt, d, nullptr, __func__); }

void pgx_post_process_packed(const pg_opt_t *opt, pg_data_t *d, const pgx_packed_t *pk)
{
	if (g_lazy != 2) PGX_FATAL("pgx_post_process_packed() needs pgx_set_lazy(2): packed input has no pg_hit_t records for results to come back in");
	post_process_impl(opt, d, pk, __func__);
}

static void post_process_impl(const pg_opt_t *opt, pg_data_t *d, const pgx_packed_t *pk, const char *who)
{
	const double t_enter = wall();
	need_device();
	if (pg_verbose >= 3) fprintf(stderr, "[M::%s] %d genes and %d proteins\n", who, d->n_gene, d->n_prot);
	{
		std::lock_guard<std::mutex> lk(g_sessions_mu);
		auto it = g_sessions.find(d);
		if (it != g_sessions.end()) { delete it->second; g_sessions.erase(it); }
	}
	double t0 = wall();
	Session *s = new Session();
	s->upload(d, true, pk);
	double t1 = wall();
	timed_run(s, opt, 1);
	double t2 = wall();
	if (timing_on()) fprintf(stderr, "[pgx timing] device check %.1f ms, session + upload of the hits %.1f ms, stage 1 beside the upload of the exons %.1f ms\n", (t0 - t_enter) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3);
	if (!g_lazy) s->download(d, nullptr);
	else {   // the protein/gene tables are cheap and always current
		s->ensure_rep();
		for (int32_t i = 0; i < d->n_prot; ++i) { d->prot[i].n = s->rep.prot_n[i]; d->prot[i].avg_score_adj = s->rep.prot_avg[i]; d->prot[i].rep = s->rep.prot_rep[i]; d->prot[i].max_score_ori = s->h_prot_max[i]; }
		for (int32_t i = 0; i < d->n_gene; ++i) d->gene[i].rep_pid = s->rep.gene_rep_pid[i];
	}
	{ std::lock_guard<std::mutex> lk(g_sessions_mu); g_sessions[d] = s; }
	int32_t n_pref = 0;
	for (int32_t i = 0; i < d->n_gene; ++i) n_pref += d->gene[i].preferred;
	if (pg_verbose >= 3) {
		fprintf(stderr, "[M::%s] %ld hits in %d genomes; %ld remain after the per-genome filters (device %.2f ms)\n", who, (long)s->N, s->G, (long)s->Nw, g_stats.run_ms[0]);
		fprintf(stderr, "[M::%s] there are %d preferred genes\n", who, n_pref);   // graph.c:31
	}
}

pg_graph_t *pg_graph_init(pg_data_t *d)   // graph.c:34-41
{
	pg_graph_t *g = (pg_graph_t*)calloc(1, sizeof(pg_graph_t));
	g->d = d;
	g->seg = (pg_seg_t*)calloc(d->n_gene > 0? d->n_gene : 1, sizeof(pg_seg_t));
	return g;
}

void pg_graph_gen(const pg_opt_t *opt, pg_graph_t *q)
{
	need_device();
	Session *s = nullptr;
	{
		std::lock_guard<std::mutex> lk(g_sessions_mu);
		auto it = g_sessions.find(q->d);
		if (it == g_sessions.end()) PGX_FATAL("pg_graph_gen(): no device state for this pg_data_t -- call pg_post_process() of libpgx first");
		s = it->second;
		if (g_lazy != 2) g_sessions.erase(it);
	}
	// the device state is keyed by the pg_data_t pointer: make sure it still describes THIS data (a freed pg_data_t whose address was reused)
	if (s->G != q->d->n_genome || s->n_prot != q->d->n_prot || s->n_gene != q->d->n_gene) PGX_FATAL("pg_graph_gen(): the pg_data_t changed since pg_post_process() (%d/%d genomes)", q->d->n_genome, s->G);
	for (int32_t j = 0; j < s->G && !s->packed_input; ++j) if (s->h_hoff[j + 1] - s->h_hoff[j] != q->d->genome[j].n_hit) PGX_FATAL("pg_graph_gen(): genome %d changed size since pg_post_process()", j);
	double t0 = wall();
	timed_run(s, opt, 2);
	double t1 = wall();
	s->download(q->d, q, g_lazy != 2);       // lazy 2: the graph only; the hits stay on the device (pgx_walk_tokens / pgx_fetch_hits / pgx_data_release)
	double t2 = wall();
	if (g_lazy != 2) delete s;
	if (timing_on()) fprintf(stderr, "[pgx timing] stage 2 %.1f ms, download %.1f ms, free %.1f ms\n", (t1 - t0) * 1e3, (t2 - t1) * 1e3, (wall() - t2) * 1e3);
}

void pg_graph_destroy(pg_graph_t *q)   // graph.c:43-47
{
	if (q == 0) return;
	free(q->g2s); free(q->seg); free(q->arc); free(q->idx);
	free(q);
}

// A pg_post_process() that is not followed by pg_graph_gen() (the CLI's --bed=raw) leaves its device state behind: hand it back.
void pgx_data_release(const pg_data_t *d)
{
	std::lock_guard<std::mutex> lk(g_sessions_mu);
	auto it = g_sessions.find(d);
	if (it != g_sessions.end()) { delete it->second; g_sessions.erase(it); }
}

static Session *session_of(const pg_data_t *d, const char *who)
{
	std::lock_guard<std::mutex> lk(g_sessions_mu);
	auto it = g_sessions.find(d);
	if (it == g_sessions.end()) PGX_FATAL("%s(): no device state for this pg_data_t (lazy mode 2 keeps it after pg_graph_gen; pgx_data_release frees it)", who);
	return it->second;
}

int64_t pgx_walk_tokens(const pg_data_t *d, pgx_walk_token_t **tokens, int64_t *genome_off)
{
	need_device();
	static_assert(sizeof(pgx_walk_token_t) == sizeof(WalkToken), "walk token layout");
	return session_of(d, __func__)->walk_tokens((WalkToken**)tokens, genome_off);
}

void pgx_fetch_hits(pg_data_t *d)
{
	need_device();
	session_of(d, __func__)->download(d, nullptr);
}

int pgx_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
void pgx_set_device(int dev)
{
	// the buffer cache and the pinned ring belong to the device they were created on: switching devices inside one process is only allowed
	// before anything was allocated (one process drives one GPU; the sharded path is one process per GPU)
	if (g_device >= 0 && g_device != dev && DevCache::get_cache().n_malloc > 0) PGX_FATAL("pgx_set_device(%d): this process already holds device memory on device %d", dev, g_device);
	g_device = dev; PGX_CHECK(cudaSetDevice(dev));
}
void pgx_set_lazy(int lazy) { g_lazy = lazy; }

// Everything the first pg_post_process() would otherwise pay for on the caller's critical path -- CUDA context, the pinned
// staging ring, the host thread pool -- done ahead of time; the CLI calls this on a side thread while the PAF files are parsed.
// Never fatal: without a device it returns 0 and pg_post_process() reports the problem in its own words.
int pgx_warmup(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) { cudaGetLastError(); return 0; }
	if (g_device >= 0 && cudaSetDevice(g_device) != cudaSuccess) { cudaGetLastError(); return 0; }
	if (cudaFree(0) != cudaSuccess) { cudaGetLastError(); return 0; }
	PinnedRing::get();
	ThreadPool::get();
	return 1;
}

int pgx_comm_unique_id(void *out128)
{
	if (!g_comm.load()) PGX_FATAL("cannot load libnccl.so.2");
	Comm::UniqueId id;
	g_comm.check(g_comm.GetUniqueId(&id), "ncclGetUniqueId");
	memcpy(out128, &id, 128);
	return 0;
}

int pgx_comm_init(int rank, int nranks, const void *id128, int32_t n_genome_total, int32_t genome_begin)
{
	need_device();
	g_comm.rank = rank; g_comm.nranks = nranks; g_comm.n_genome_total = n_genome_total; g_comm.genome_begin = genome_begin;
	if (nranks > 1) {
		if (!g_comm.load()) PGX_FATAL("cannot load libnccl.so.2");
		Comm::UniqueId id;
		memcpy(&id, id128, 128);
		g_comm.check(g_comm.CommInitRank(&g_comm.comm, nranks, id, rank), "ncclCommInitRank");
	}
	return 0;
}

int pgx_comm_init_loopback(int rank, int nranks, const char *segment_name, int32_t n_genome_total, int32_t genome_begin)
{
	need_device();
	g_comm.rank = rank; g_comm.nranks = nranks; g_comm.n_genome_total = n_genome_total; g_comm.genome_begin = genome_begin;
	if (nranks > 1) g_comm.shm_open_segment(segment_name, rank, nranks, (size_t)64 << 20);
	return 0;
}

void pgx_comm_destroy(void)
{
	g_comm.shm_close();
	if (g_comm.comm) { g_comm.CommDestroy(g_comm.comm); g_comm.comm = nullptr; }
	g_comm.rank = 0; g_comm.nranks = 1; g_comm.n_genome_total = -1; g_comm.genome_begin = 0;
}

struct pgx_session_s { Session s; };

pgx_session_t *pgx_upload(const pg_data_t *d) { need_device(); pgx_session_t *h = new pgx_session_s(); h->s.upload(d); return h; }
void pgx_run(pgx_session_t *h, const pg_opt_t *opt, int stage) { timed_run(&h->s, opt, stage); }
void pgx_download(pgx_session_t *h, pg_data_t *d, pg_graph_t *q) { h->s.download(d, q); }
void pgx_session_free(pgx_session_t *h) { delete h; }

void pgx_get_stats(pgx_stats_t *o)
{
	o->launches = g_launch.launches; o->sweep_ms = g_stats.sweep_ms; o->sweep_bytes = g_stats.sweep_bytes;
	for (int i = 0; i < 3; ++i) o->run_ms[i] = g_stats.run_ms[i];
	o->n_hit = g_stats.n_hit; o->n_active = g_stats.n_active;
	for (int i = 0; i < 4; ++i) o->tie[i] = g_stats.tie[i];
	o->n_seg = g_stats.n_seg; o->n_arc = g_stats.n_arc; o->h2d_bytes = g_stats.h2d; o->d2h_bytes = g_stats.d2h;
	o->fast_rounds = g_stats.fast_rounds; o->redo = g_stats.redo;
}
void pgx_reset_stats(void) { g_launch.launches = 0; g_stats = Stats(); }
void pgx_set_sweep_stat(int on) { g_sweep_stat = on; }
void pgx_trim(void) { DevCache::get_cache().trim(); SpareHitArrays::get().trim(); }

// ---- test hooks -----------------------------------------------------------------------------------
void pgx_test_radix_sort(uint64_t *keys, uint32_t *vals, int64_t n, int n_bits)
{
	need_device();
	DBuf<uint64_t> k, ka; DBuf<uint32_t> v, va, hist, agg;
	k.alloc(n + 1); ka.alloc(n + 1); v.alloc(n + 1); va.alloc(n + 1);
	size_t he = RadixSortBuffers::hist_elems(n);
	hist.alloc(he); agg.alloc(scan_scratch_elems((int64_t)he));
	PGX_CHECK(cudaMemcpy(k.p, keys, n * 8, cudaMemcpyHostToDevice));
	RadixSortBuffers b; b.keys_alt = ka.p; b.vals_alt = va.p; b.hist = hist.p; b.scan_agg = agg.p;
	uint64_t *kr; uint32_t *vr;
	radix_sort_pairs(k.p, v.p, n, n_bits, true, b, &kr, &vr, 0);
	PGX_CHECK(cudaDeviceSynchronize());
	PGX_CHECK(cudaMemcpy(keys, kr, n * 8, cudaMemcpyDeviceToHost));
	PGX_CHECK(cudaMemcpy(vals, vr, n * 4, cudaMemcpyDeviceToHost));
}

void pgx_test_exclusive_sum(const uint32_t *in, uint32_t *out, int64_t n)
{
	need_device();
	DBuf<uint32_t> a, o, agg;
	a.alloc(n + 1); o.alloc(n + 2); agg.alloc(scan_scratch_elems(n));
	PGX_CHECK(cudaMemcpy(a.p, in, n * 4, cudaMemcpyHostToDevice));
	exclusive_sum_u32(a.p, o.p, n, agg.p, 0);
	PGX_CHECK(cudaDeviceSynchronize());
	PGX_CHECK(cudaMemcpy(out, o.p, (n + 1) * 4, cudaMemcpyDeviceToHost));
}

void pgx_test_segmented_max(const int32_t *seg, const int32_t *val, int32_t *out, int64_t n)
{
	need_device();
	DBuf<int32_t> s, v, o; DBuf<SegVal> agg;
	s.alloc(n + 1); v.alloc(n + 1); o.alloc(n + 1); agg.alloc((size_t)(n / SCAN_TILE + 2));
	PGX_CHECK(cudaMemcpy(s.p, seg, n * 4, cudaMemcpyHostToDevice));
	PGX_CHECK(cudaMemcpy(v.p, val, n * 4, cudaMemcpyHostToDevice));
	segmented_running_max(s.p, v.p, o.p, n, agg.p, 0);
	PGX_CHECK(cudaDeviceSynchronize());
	PGX_CHECK(cudaMemcpy(out, o.p, n * 4, cudaMemcpyDeviceToHost));
}

void pgx_test_flag_sort_replay(pg128_t *a, int64_t n) { FlagSortReplay::run(a, n); }

// hit.c:205-217 on the host (no device needed): gene_rep_pid[n_gene] from the per-protein sums; returns 1 if the exact sort order had to be replayed
int pgx_test_choose_representatives(const unsigned long long *z, const int32_t *prot_gid, int32_t n_prot, int32_t n_gene, int32_t force_replay, int32_t *gene_rep_pid)
{
	RepResult r;
	choose_representatives(z, prot_gid, n_prot, n_gene, r, force_replay != 0);
	memcpy(gene_rep_pid, r.gene_rep_pid.data(), (size_t)n_gene * 4);
	return r.replayed? 1 : 0;
}

} // extern "C"
