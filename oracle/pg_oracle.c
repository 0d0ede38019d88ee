/*
 * pg_oracle.c -- TEST INFRASTRUCTURE ONLY.  A CPU restatement, in plain C, of pangene's hot path
 * (parsed PAF hits -> GFA arcs).  It is NOT the product and nothing in the product links or calls it:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load libpgoracle.so, and
 * there only as the checker.  Parity status: PINNED -- tests/test_oracle.py checks this file against the
 * compiled, unmodified reference (oracle/_ref, built from /root/reference by oracle/Makefile) on test/C4
 * (GFA/BED md5s of SURVEY.md §4) and on randomized synthetic inputs across option combinations, and
 * against the committed golden files in tests/golden/.
 *
 * The restatement deliberately has the SHAPE of the CUDA design, not of the reference code:
 *   - hits are sorted ONCE per genome, stably, by (contig, cs); the (contig, cm) order is an index
 *     permutation computed once -- the reference re-sorts the 88-byte structs 67 times per genome with
 *     an unstable radix sort (hit.c:29-64);
 *   - the interval sweeps (overlap.c:58-178) are a set-of-pairs computation evaluated from BOTH members
 *     of each pair so that every per-hit result is a local reduction (max key, smallest partner index)
 *     -- no order dependence, no scatter;
 *   - per-(genome,protein) and per-(genome,gene) group logic (hit.c:66-146) uses dense tables and
 *     commutative min/max/sum updates instead of a sort + sequential group walk;
 *   - arcs (graph.c:87-177) are emitted as records and collapsed by a (key, genome) sort.
 * Where the reference's result depends on the exact permutation of its unstable radix sort and that is
 * observable in the GFA (hit.c:209, vertex.c:54: SURVEY.md H1 b,c), the permutation is replayed exactly
 * (pgo_radix_replay128).  For coincident coordinates (H1 a) the reference's order follows from its whole chain of
 * hit sorts; the chain is replayed per contig (replay_contig_order) in the two places where randomized campaigns
 * against the compiled reference showed the order to be observable -- the never-reset hit at the front of the
 * array (overlap.c:108: ref_first_pos) and the ordinals of cs-tied survivors (branch.c:18-25: mark_branch_arcs);
 * the remaining coincidences are counted by a detector (pgo_tie_report).
 *
 * Every function cites the reference lines it restates.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdint.h>
#include <limits.h>
#include "pg_oracle.h"

#define NEW(type, n)  ((type*)malloc(sizeof(type) * ((int64_t)(n) > 0? (size_t)(n) : 1)))
#define NEW0(type, n) ((type*)calloc(((int64_t)(n) > 0? (size_t)(n) : 1), sizeof(type)))

static int64_t g_tie[4];

void pgo_tie_report(int64_t out[4]) { memcpy(out, g_tie, sizeof(g_tie)); }

/* ------------------------------------------------------------------------------------------------ */
/* exact replay of the reference radix sort's permutation (ksort.h:37-87, 8-bit digits, 8-byte key)  */
/* ------------------------------------------------------------------------------------------------ */
static void replay_insertion(pg128_t *a, int64_t n)   /* ksort.h:41-51: stable */
{
	int64_t i, j;
	for (i = 1; i < n; ++i) {
		if (a[i].x < a[i - 1].x) {
			pg128_t t = a[i];
			for (j = i; j > 0 && t.x < a[j - 1].x; --j) a[j] = a[j - 1];
			a[j] = t;
		}
	}
}

static void replay_level(pg128_t *a, int64_t n, int shift)   /* ksort.h:52-82 */
{
	int64_t head[256], tail[256], start[256], c;
	int k;
	memset(tail, 0, sizeof(tail));
	for (c = 0; c < n; ++c) ++tail[a[c].x >> shift & 0xff];
	for (k = 0, c = 0; k < 256; ++k) { head[k] = start[k] = c; c += tail[k]; tail[k] = c; }
	/* in-place permutation by cycle leaders, buckets taken in ascending order; elements already in their
	 * bucket are skipped one at a time, a displaced element is carried until one for the current bucket
	 * turns up (ksort.h:62-74).  The order in which equal-digit elements land is what makes the sort
	 * unstable -- and is exactly what is being reproduced here. */
	for (k = 0; k < 256;) {
		if (head[k] == tail[k]) { ++k; continue; }
		int l = (int)(a[head[k]].x >> shift & 0xff);
		if (l == k) { ++head[k]; continue; }
		pg128_t carry = a[head[k]];
		do {
			pg128_t t = a[head[l]];
			a[head[l]++] = carry;
			carry = t;
			l = (int)(carry.x >> shift & 0xff);
		} while (l != k);
		a[head[k]++] = carry;
	}
	if (shift == 0) return;
	shift = shift > 8? shift - 8 : 0;
	for (k = 0; k < 256; ++k) {
		int64_t m = tail[k] - start[k];
		if (m > 64) replay_level(a + start[k], m, shift);
		else if (m > 1) replay_insertion(a + start[k], m);
	}
}

void pgo_radix_replay128(pg128_t *a, int64_t n)   /* ksort.h:83-87 */
{
	if (n <= 64) replay_insertion(a, n);
	else replay_level(a, n, 56);
}

/* ------------------------------------------------------------------------------------------------ */
/* small helpers                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
static inline uint32_t mix32(uint32_t k)   /* pgpriv.h:88-97: an invertible 32-bit mixer, used to break score ties by protein */
{
	k += ~(k << 15); k ^= k >> 10; k += k << 3; k ^= k >> 6; k += ~(k << 11); k ^= k >> 16;
	return k;
}

/* per-genome working view; requires g->hit sorted by (contig, cs) */
typedef struct {
	pg_genome_t *g;
	int32_t n;
	int32_t *gid;      /* gene of each hit */
	int32_t *cds;      /* CDS length, overlap.c:45-51 */
	uint64_t *key;     /* score_adj<<33 | preferred<<32 | mix32(pid), overlap.c:82,137 */
	int64_t *pmax;     /* running max of ce inside the contig: bounds the backward scan */
	int32_t *by_cm;    /* permutation: contig-major, then cm; ties keep (contig, cs) order */
	int32_t j;         /* genome index */
} view_t;

struct sort_ctx { const pg_hit_t *h; int by_cm; };

static int cmp_hit_idx(const void *pa, const void *pb, void *pc)
{
	const struct sort_ctx *c = (const struct sort_ctx*)pc;
	int32_t a = *(const int32_t*)pa, b = *(const int32_t*)pb;
	const pg_hit_t *x = &c->h[a], *y = &c->h[b];
	int64_t kx = c->by_cm? x->cm : x->cs, ky = c->by_cm? y->cm : y->cs;
	if (x->cid != y->cid) return x->cid < y->cid? -1 : 1;
	if (kx != ky) return kx < ky? -1 : 1;
	return a < b? -1 : a > b;
}

/* File position of every hit of every genome, in the (contig, cs) order the hits are kept in: the input the reference's
 * sort chain starts from (replay_contig_order below).  Filled by pgo_post_process(), used by the pgo_graph_gen() that follows
 * on the same pg_data_t (test infrastructure: one data set at a time). */
static int32_t **g_file_pos, g_file_pos_n;

static void file_pos_reset(int32_t n_genome)
{
	int32_t j;
	for (j = 0; j < g_file_pos_n; ++j) free(g_file_pos[j]);
	free(g_file_pos);
	g_file_pos = NEW0(int32_t*, n_genome); g_file_pos_n = n_genome;
}

/* stable (contig, cs) order of one genome; stands in for the first pg_hit_sort(g,0) of read.c:247.
 * The reference's own tie order is history-dependent; ties here keep file order, and the places where the
 * reference's order is observable replay its sort chain (replay_contig_order). */
static void sort_by_cs(pg_genome_t *g, int32_t **file_pos)
{
	int32_t i, *idx = NEW(int32_t, g->n_hit);
	struct sort_ctx c = { g->hit, 0 };
	pg_hit_t *a;
	for (i = 0; i < g->n_hit; ++i) idx[i] = i;
	qsort_r(idx, g->n_hit, sizeof(int32_t), cmp_hit_idx, &c);
	a = NEW(pg_hit_t, g->n_hit);
	for (i = 0; i < g->n_hit; ++i) a[i] = g->hit[idx[i]];
	if (file_pos) *file_pos = idx; else free(idx);
	free(g->hit);
	g->hit = a; g->m_hit = g->n_hit;
}

static int cmp_i64pair(const void *a, const void *b) { int64_t x = *(const int64_t*)a, y = *(const int64_t*)b; return x < y? -1 : x > y; }

/* The reference re-sorts every genome's hit array with an UNSTABLE radix sort, alternating keys: once by cs when the file is
 * loaded (read.c:247), then one (cm, cs) pair per pg_gen_arc (graph.c:103,123) and per pg_mark_branch_flt_hit
 * (branch.c:116,140).  pg_hit_sort (hit.c:29-64) first buckets by contig, keeping the array order, then radix-sorts each
 * contig's (key, index) pairs (radix_sort_pg128x) -- so where hits with equal keys end up depends only on the contig's key
 * sequences and on the number of sorts so far, never on flags.  This replays that chain for the hits [lo, hi) of one contig
 * (positions in OUR (contig, cs, file order) array) and returns rank[p - lo] = position of hit p inside the contig in the
 * reference's array after `n_pairs` (cm, cs) pairs (by_cm = 0) or after one more sort by cm (by_cm = 1). */
static void replay_contig_order(const pg_genome_t *g, const int32_t *file_pos, int32_t lo, int32_t hi, int32_t n_pairs, int by_cm, int32_t *rank)
{
	int32_t n = hi - lo, i, p;
	int64_t *fo = NEW(int64_t, 2 * (size_t)n);
	pg128_t *a = NEW(pg128_t, n);
	for (i = 0; i < n; ++i) fo[2 * i] = file_pos[lo + i], fo[2 * i + 1] = lo + i;
	qsort(fo, n, 16, cmp_i64pair);                                  /* the contig's hits in file order */
	for (i = 0; i < n; ++i) a[i].y = (uint64_t)fo[2 * i + 1];
	for (i = 0; i < n; ++i) a[i].x = (uint64_t)g->hit[a[i].y].cs;
	pgo_radix_replay128(a, n);                                      /* read.c:247 */
	for (p = 0; p < n_pairs + (by_cm? 1 : 0); ++p) {
		for (i = 0; i < n; ++i) a[i].x = (uint64_t)g->hit[a[i].y].cm;
		pgo_radix_replay128(a, n);
		if (p == n_pairs) break;
		for (i = 0; i < n; ++i) a[i].x = (uint64_t)g->hit[a[i].y].cs;
		pgo_radix_replay128(a, n);
	}
	for (i = 0; i < n; ++i) rank[a[i].y - lo] = i;
	free(a); free(fo);
}

/* test hook (tests/test_oracle.py): the reference's array order of one genome, given in FILE order, after read.c:247 and
 * `n_pairs` further (cm, cs) sort pairs (+ one sort by cm if by_cm): order_out[k] = file index of the hit at array position k.
 * Checked against the compiled reference's own pg_hit_sort() chain. */
void pgo_test_chain_order(const pg_genome_t *g0, int32_t n_pairs, int32_t by_cm, int32_t *order_out)
{
	pg_genome_t g = *g0;
	int32_t *file_pos = 0, lo, hi, i, *rank;
	g.hit = NEW(pg_hit_t, g0->n_hit);
	memcpy(g.hit, g0->hit, sizeof(pg_hit_t) * g0->n_hit);
	sort_by_cs(&g, &file_pos);
	rank = NEW(int32_t, g.n_hit);
	for (lo = 0; lo < g.n_hit; lo = hi) {
		for (hi = lo; hi < g.n_hit && g.hit[hi].cid == g.hit[lo].cid; ++hi) {}
		replay_contig_order(&g, file_pos, lo, hi, n_pairs, by_cm, rank + lo);
		for (i = lo; i < hi; ++i) order_out[lo + rank[i]] = file_pos[i];
	}
	free(rank); free(file_pos); free(g.hit);
}

static int32_t g_sort_pairs;   /* (cm, cs) sort pairs the reference has executed so far in pg_graph_gen */

/* Which hit sits at position 0 of the reference's array right now?  pg_shadow (overlap.c:108) starts its loop at i = 1, so
 * that hit's shadow bit is never reset.  It is a hit with the smallest cs of the first contig; when several share that cs, the
 * reference's sort chain decides which of them is in front at this point -- replay it. */
static int32_t ref_first_pos(const pg_genome_t *g, int32_t j)
{
	int32_t hi, i, first = 0, *rank;
	if (g->n_hit < 2 || g->hit[1].cid != g->hit[0].cid || g->hit[1].cs != g->hit[0].cs) return 0;
	if (g_file_pos == 0 || j >= g_file_pos_n || g_file_pos[j] == 0) return 0;
	for (hi = 1; hi < g->n_hit && g->hit[hi].cid == g->hit[0].cid; ++hi) {}
	rank = NEW(int32_t, hi);
	replay_contig_order(g, g_file_pos[j], 0, hi, g_sort_pairs, 0, rank);
	for (i = 0; i < hi; ++i) if (rank[i] == 0) first = i;
	free(rank);
	return first;
}

static void view_init(view_t *v, const pg_data_t *d, int32_t j)
{
	pg_genome_t *g = &d->genome[j];
	int32_t i, k, n = g->n_hit;
	struct sort_ctx c = { g->hit, 1 };
	v->g = g; v->n = n; v->j = j;
	v->gid = NEW(int32_t, n); v->cds = NEW(int32_t, n); v->key = NEW(uint64_t, n);
	v->pmax = NEW(int64_t, n); v->by_cm = NEW(int32_t, n);
	for (i = 0; i < n; ++i) {
		const pg_hit_t *a = &g->hit[i];
		int32_t len = 0, gid = d->prot[a->pid].gid;
		for (k = 0; k < a->n_exon; ++k) len += g->exon[a->off_exon + k].oe - g->exon[a->off_exon + k].os;
		v->gid[i] = gid; v->cds[i] = len;
		v->key[i] = (uint64_t)a->score_adj << 33 | (uint64_t)d->gene[gid].preferred << 32 | mix32(a->pid);
		v->pmax[i] = (i > 0 && g->hit[i - 1].cid == a->cid && v->pmax[i - 1] > a->ce)? v->pmax[i - 1] : a->ce;
		v->by_cm[i] = i;
	}
	qsort_r(v->by_cm, n, sizeof(int32_t), cmp_hit_idx, &c);
}

static void view_free(view_t *v) { free(v->gid); free(v->cds); free(v->key); free(v->pmax); free(v->by_cm); }

/* |CDS(lo) & CDS(hi)| for two hits on one contig (overlap.c:6-42; the union it also computes is unused).
 * `lo` is the hit that comes first in (contig, cs) order, as in the reference's call pg_hit_overlap(g, aj, ai). */
static int32_t cds_overlap(const pg_genome_t *g, const pg_hit_t *lo, const pg_hit_t *hi)
{
	const pg_exon_t *e0 = &g->exon[lo->off_exon], *e1 = &g->exon[hi->off_exon];
	int32_t i0 = 0, i1 = 0, inter = 0;
	if (lo->cid != hi->cid || !(lo->cs < hi->ce && lo->ce > hi->cs)) return 0;
	while (i0 < lo->n_exon && i1 < hi->n_exon) {
		int64_t s0 = lo->cs + e0[i0].os, t0 = lo->cs + e0[i0].oe, s1 = hi->cs + e1[i1].os, t1 = hi->cs + e1[i1].oe;
		if (s0 < s1) {             /* exon of `lo` starts first */
			if (t0 < t1) { int64_t o = t0 - s1; inter += o > 0? o : 0; ++i0; }
			else { inter += e1[i1].oe - e1[i1].os; ++i1; }      /* hi's exon lies inside lo's */
		} else {                   /* exon of `hi` starts first (or both start together) */
			if (t1 < t0) { int64_t o = t1 - s0; inter += o > 0? o : 0; ++i1; }
			else { inter += e0[i0].oe - e0[i0].os; ++i0; }
		}
	}
	return inter;
}

/* ------------------------------------------------------------------------------------------------ */
/* the interval sweep, as a set of pairs                                                             */
/* ------------------------------------------------------------------------------------------------ */
enum { SWEEP_SHADOW = 0, SWEEP_ISOFORM = 1 };

/* Decide pair (lo < hi in (contig,cs) order, both unfiltered, ce_lo > cs_hi).  Returns -1 if the pair is
 * not in relation, 0 if `hi` loses, 1 if `lo` loses; *ov = CDS overlap.
 * SWEEP_SHADOW: overlap.c:121-146.  SWEEP_ISOFORM: overlap.c:72-86. */
static int pair_decide(const pg_opt_t *opt, const view_t *v, int32_t lo, int32_t hi, int mode, int check_strand, int32_t *ov_)
{
	const pg_hit_t *a = &v->g->hit[hi], *b = &v->g->hit[lo];   /* a = "i", b = "j" of the reference loops */
	int32_t ov, same_gene = v->gid[hi] == v->gid[lo];
	uint64_t si = v->key[hi], sj = v->key[lo];
	if (mode == SWEEP_ISOFORM && !same_gene) return -1;
	if (check_strand && a->rev != b->rev) return -1;
	ov = cds_overlap(v->g, b, a);
	if (ov == 0) return -1;
	*ov_ = ov;
	if (mode == SWEEP_SHADOW && !same_gene) {
		int32_t shorter = v->cds[hi] < v->cds[lo]? v->cds[hi] : v->cds[lo];
		double cov_short = (double)(uint64_t)ov / shorter;
		if (cov_short < opt->min_ov_ratio) return -1;
	}
	if (mode == SWEEP_ISOFORM || same_gene || a->weak_br == b->weak_br)
		return (si < sj || (si == sj && a->rank > b->rank))? 0 : 1;
	return a->weak_br > b->weak_br? 0 : 1;
}

typedef struct { uint64_t key; int32_t partner, ov; } dom_t;

/* One sweep over one genome. SWEEP_SHADOW == pg_shadow (overlap.c:101-178); SWEEP_ISOFORM ==
 * pg_flt_ov_isoform (overlap.c:58-93). Returns the count the reference returns. */
static int32_t sweep(const pg_opt_t *opt, view_t *v, int mode, int cal_dom_sc, int check_strand)
{
	pg_hit_t *h = v->g->hit;
	int32_t n = v->n, i, p, cnt = 0, first = mode == SWEEP_SHADOW? ref_first_pos(v->g, v->j) : 0;
	uint8_t *lose = NEW0(uint8_t, n);
	dom_t *best = NEW0(dom_t, n);
	for (i = 0; i < n; ++i) {
		int32_t side;
		if (h[i].flt) continue;
		best[i].partner = INT32_MAX;
		for (side = 0; side < 2; ++side) {
			/* side 0: earlier hits still open at cs_i; side 1: later hits starting before ce_i */
			for (p = side == 0? i - 1 : i + 1; p >= 0 && p < n; p += side == 0? -1 : 1) {
				int32_t lo, hi, ov = 0, r;
				if (h[p].cid != h[i].cid) break;
				if (side == 0) { if (v->pmax[p] <= h[i].cs) break; if (h[p].ce <= h[i].cs) continue; }
				else if (h[p].cs >= h[i].ce) break;
				if (h[p].flt) continue;
				lo = side == 0? p : i, hi = side == 0? i : p;
				r = pair_decide(opt, v, lo, hi, mode, check_strand, &ov);
				if (r < 0 || (r == 0? hi : lo) != i) continue;
				lose[i] = 1;   /* i loses to p */
				if (v->key[p] > best[i].key || (v->key[p] == best[i].key && v->key[p] > 0 && p < best[i].partner)) {
					if (cal_dom_sc && v->key[p] == best[i].key && h[p].cs == h[best[i].partner].cs) ++g_tie[2];
					best[i].key = v->key[p], best[i].partner = p, best[i].ov = ov;
				} else if (cal_dom_sc && v->key[p] == best[i].key && v->key[p] > 0 && h[p].cs == h[best[i].partner].cs) ++g_tie[2];
			}
		}
	}
	if (mode == SWEEP_ISOFORM) {
		for (i = 0; i < n; ++i) if (lose[i]) h[i].flt_iso_ov = 1;
		for (i = 0; i < n; ++i) if (h[i].flt_iso_ov) h[i].flt = 1, ++cnt;   /* overlap.c:89-91 */
	} else {
		for (i = 0; i < n; ++i) {
			pg_hit_t *a = &h[i];
			if (a->flt) continue;
			/* overlap.c:108,113: the sweep starts at i = 1, so the hit at the front of the reference's array is never reset */
			a->shadow = i == first? (a->shadow | lose[i]) : lose[i];
			a->pid_dom = -1;
			if (cal_dom_sc) a->score_dom = -1;
			if (best[i].key > 0) {   /* overlap.c:162 */
				const pg_hit_t *w = &h[best[i].partner];
				a->pid_dom = w->pid;
				if (cal_dom_sc)      /* overlap.c:170, IEEE double, evaluated left to right, no contraction */
					a->score_dom = (int32_t)(a->score_ori * (1.0 - (double)best[i].ov / v->cds[i]) + w->score_ori * ((double)best[i].ov / v->cds[best[i].partner]) + .499);
			}
			if (a->shadow) ++cnt;
		}
	}
	free(lose); free(best);
	return cnt;
}

/* ------------------------------------------------------------------------------------------------ */
/* the per-genome load-time block (read.c:243-260)                                                   */
/* ------------------------------------------------------------------------------------------------ */

/* hit.c:66-105 with a dense per-protein table instead of a sort: per protein of this genome, the exon-count
 * range decides which hits look like processed pseudogenes; the lowest-rank survivor is promoted to rank 0. */
static void flag_pseudo(const pg_data_t *d, pg_genome_t *g)
{
	int32_t i, np = d->n_prot;
	int32_t *mx = NEW0(int32_t, np), *mn = NEW(int32_t, np), *r1 = NEW(int32_t, np);
	for (i = 0; i < np; ++i) mn[i] = INT32_MAX, r1[i] = INT32_MAX;
	for (i = 0; i < g->n_hit; ++i) {
		const pg_hit_t *a = &g->hit[i];
		if (a->n_exon > mx[a->pid]) mx[a->pid] = a->n_exon;
		if (a->n_exon < mn[a->pid]) mn[a->pid] = a->n_exon;
	}
	for (i = 0; i < g->n_hit; ++i) {
		pg_hit_t *a = &g->hit[i];
		int32_t M = mx[a->pid], m = mn[a->pid];
		if (!(M > 1 && (m == 1 || m * 2 <= M))) continue;
		if (a->n_exon == 1 || a->n_exon * 2 <= M) a->pseudo = 1;
		else if (a->rank < r1[a->pid]) r1[a->pid] = a->rank;
	}
	for (i = 0; i < g->n_hit; ++i) {
		pg_hit_t *a = &g->hit[i];
		int32_t M = mx[a->pid], m = mn[a->pid], r = r1[a->pid];
		if (!(M > 1 && (m == 1 || m * 2 <= M)) || r <= 0 || r == INT32_MAX) continue;
		if (a->rank < r) a->rank++;
		else if (a->rank == r) a->rank = 0;
	}
	free(mx); free(mn); free(r1);
}

/* hit.c:130-146 */
static void flt_chain_shadow(const pg_data_t *d, pg_genome_t *g)
{
	int32_t i;
	uint8_t *some_kept = NEW0(uint8_t, d->n_prot);   /* protein has a hit in this genome that is not flt_iso_ov */
	for (i = 0; i < g->n_hit; ++i) if (!g->hit[i].flt_iso_ov) some_kept[g->hit[i].pid] = 1;
	for (i = 0; i < g->n_hit; ++i) {
		pg_hit_t *a = &g->hit[i];
		if (a->pid_dom0 >= 0 && !some_kept[a->pid_dom0]) a->flt = a->flt_chain = 1;
	}
	free(some_kept);
}

/* hit.c:107-128: first-wins, strict, signed-vs-unsigned compare reproduced by the same conversions */
static void flt_subopt_isoform(const pg_data_t *d, pg_genome_t *g)
{
	int32_t i;
	uint64_t *best = NEW0(uint64_t, d->n_gene);
	int32_t *best_i = NEW(int32_t, d->n_gene);
	for (i = 0; i < d->n_gene; ++i) best_i[i] = -1;
	for (i = 0; i < g->n_hit; ++i) {
		const pg_hit_t *a = &g->hit[i];
		int32_t gid;
		if (a->flt || a->rank > 0) continue;
		gid = d->prot[a->pid].gid;
		if ((uint64_t)(int64_t)a->score_adj > best[gid] >> 32) best[gid] = (uint64_t)(int64_t)a->score_adj << 32 | (uint32_t)a->pid, best_i[gid] = i;
		else if (best_i[gid] >= 0 && (uint64_t)(int64_t)a->score_adj == best[gid] >> 32 && a->pid != (int32_t)best[gid]
				 && a->cid == g->hit[best_i[gid]].cid && a->cs == g->hit[best_i[gid]].cs) ++g_tie[3];
	}
	for (i = 0; i < g->n_hit; ++i) {
		pg_hit_t *a = &g->hit[i];
		if (a->flt) continue;
		if (a->pid != (int32_t)best[d->prot[a->pid].gid]) a->flt = a->flt_iso_sub_self = 1;
	}
	free(best); free(best_i);
}

static void load_time_block(const pg_opt_t *opt, pg_data_t *d, int32_t j)
{
	pg_genome_t *g = &d->genome[j];
	int32_t i, check_strand = !!(opt->flag & PG_F_CHECK_STRAND);
	view_t v;
	flag_pseudo(d, g);                                                    /* read.c:245 */
	for (i = 0; i < g->n_hit; ++i) if (g->hit[i].pseudo) g->hit[i].flt = 1; /* read.c:246 */
	sort_by_cs(g, &g_file_pos[j]);                                        /* read.c:247 */
	if (g->n_hit > 1 && g->hit[0].cid == g->hit[1].cid && g->hit[0].cs == g->hit[1].cs && !g->hit[0].flt && !g->hit[1].flt) ++g_tie[1];
	view_init(&v, d, j);
	sweep(opt, &v, SWEEP_SHADOW, 1, check_strand);                        /* read.c:248 */
	for (i = 0; i < g->n_hit; ++i) {                                      /* read.c:249-253 */
		pg_hit_t *a = &g->hit[i];
		a->pid_dom0 = a->pid_dom; a->pid_dom = -1; a->shadow = 0;
	}
	sweep(opt, &v, SWEEP_ISOFORM, 0, check_strand);                       /* read.c:254 */
	flt_chain_shadow(d, g);                                               /* read.c:255 */
	flt_subopt_isoform(d, g);                                             /* read.c:256 */
	view_free(&v);
}

/* ------------------------------------------------------------------------------------------------ */
/* pg_post_process (graph.c:7-32)                                                                    */
/* ------------------------------------------------------------------------------------------------ */
static void cap_score_dom(pg_data_t *d)   /* hit.c:227-247 */
{
	int32_t i, j;
	for (i = 0; i < d->n_prot; ++i) d->prot[i].max_score_ori = 0;
	for (j = 0; j < d->n_genome; ++j)
		for (i = 0; i < d->genome[j].n_hit; ++i) {
			const pg_hit_t *a = &d->genome[j].hit[i];
			if (a->score_ori > d->prot[a->pid].max_score_ori) d->prot[a->pid].max_score_ori = a->score_ori;
		}
	for (j = 0; j < d->n_genome; ++j)
		for (i = 0; i < d->genome[j].n_hit; ++i) {
			pg_hit_t *a = &d->genome[j].hit[i];
			if (a->score_dom > d->prot[a->pid].max_score_ori) a->score_dom = d->prot[a->pid].max_score_ori;
		}
}

static void flag_representative(pg_data_t *d)   /* hit.c:189-225 */
{
	int32_t i, j;
	pg128_t *z = NEW0(pg128_t, d->n_prot);
	for (i = 0; i < d->n_gene; ++i) d->gene[i].rep_pid = -1;
	for (i = 0; i < d->n_prot; ++i) z[i].y = i, d->prot[i].rep = 0;
	for (j = 0; j < d->n_genome; ++j)
		for (i = 0; i < d->genome[j].n_hit; ++i) {
			pg_hit_t *a = &d->genome[j].hit[i];
			if (a->rank == 0 && !a->flt) z[a->pid].x += (uint64_t)(int64_t)a->score_adj << 32 | 1;
			a->rep = 0;
		}
	for (i = 0; i < d->n_prot; ++i) {
		d->prot[i].n = (int32_t)(uint32_t)z[i].x;
		d->prot[i].avg_score_adj = d->prot[i].n? (int32_t)((double)(z[i].x >> 32) / d->prot[i].n + .499) : 0;
	}
	pgo_radix_replay128(z, d->n_prot);   /* tie order among equal sums decides rep_pid */
	for (i = d->n_prot - 1; i >= 0; --i) {
		int32_t pid = (int32_t)z[i].y, gid = d->prot[pid].gid;
		if (d->gene[gid].rep_pid < 0) d->gene[gid].rep_pid = pid, d->prot[pid].rep = 1;
	}
	free(z);
	for (j = 0; j < d->n_genome; ++j)
		for (i = 0; i < d->genome[j].n_hit; ++i)
			if (d->prot[d->genome[j].hit[i].pid].rep) d->genome[j].hit[i].rep = 1;
}

static void flag_pseudo_joint(const pg_opt_t *opt, pg_data_t *d)   /* hit.c:153-187 */
{
	int32_t i, j;
	int32_t *c = NEW0(int32_t, 2 * (size_t)d->n_prot);
	int64_t *s = NEW0(int64_t, 2 * (size_t)d->n_prot);
	for (j = 0; j < d->n_genome; ++j)
		for (i = 0; i < d->genome[j].n_hit; ++i) {
			const pg_hit_t *a = &d->genome[j].hit[i];
			if (a->flt || a->rank != 0) continue;
			c[2 * a->pid + (a->n_exon != 1)]++;
			s[2 * a->pid + (a->n_exon != 1)] += a->score_ori;
		}
	for (j = 0; j < d->n_genome; ++j)
		for (i = 0; i < d->genome[j].n_hit; ++i) {
			pg_hit_t *a = &d->genome[j].hit[i];
			const int32_t *pc = &c[2 * a->pid];
			const int64_t *ps = &s[2 * a->pid];
			if (a->flt || a->pseudo || a->n_exon != 1) continue;
			if (pc[1] > 0 && pc[1] >= d->n_genome * opt->min_vertex_ratio && ((double)ps[1] / pc[1]) / ((double)ps[0] / pc[0]) >= 0.99) a->pseudo = 1;
			else if ((pc[1] == 0 || pc[1] <= d->n_genome * opt->min_vertex_ratio) && (opt->flag & PG_F_DROP_SGL_EXON)) a->pseudo = 1;
		}
	free(c); free(s);
}

void pgo_post_process(const pg_opt_t *opt, pg_data_t *d)
{
	int32_t j, check_strand = !!(opt->flag & PG_F_CHECK_STRAND);
	memset(g_tie, 0, sizeof(g_tie));
	g_sort_pairs = 0;
	file_pos_reset(d->n_genome);
	for (j = 0; j < d->n_genome; ++j) load_time_block(opt, d, j);
	cap_score_dom(d);
	flag_representative(d);
	if (!(opt->flag & PG_F_NO_JOINT_PSEUDO)) flag_pseudo_joint(opt, d);
	for (j = 0; j < d->n_genome; ++j) {   /* graph.c:20-28 */
		view_t v;
		view_init(&v, d, j);
		sweep(opt, &v, SWEEP_SHADOW, 0, check_strand);
		view_free(&v);
	}
}

/* ------------------------------------------------------------------------------------------------ */
/* graph construction                                                                                */
/* ------------------------------------------------------------------------------------------------ */
pg_graph_t *pgo_graph_init(pg_data_t *d)   /* graph.c:34-41 */
{
	pg_graph_t *q = NEW0(pg_graph_t, 1);
	q->d = d;
	q->seg = NEW0(pg_seg_t, d->n_gene);
	return q;
}

void pgo_graph_destroy(pg_graph_t *q) { free(q->g2s); free(q->seg); free(q->arc); free(q->idx); free(q); }

#define SET_FILTER(d, cond) do { int32_t j_, i_; for (j_ = 0; j_ < (d)->n_genome; ++j_) for (i_ = 0; i_ < (d)->genome[j_].n_hit; ++i_) { pg_hit_t *a = &(d)->genome[j_].hit[i_]; if (cond) a->flt = 1; } } while (0)   /* pgpriv.h:109-116 */

static void gen_g2s(pg_graph_t *q)   /* graph.c:49-59 */
{
	int32_t i;
	free(q->g2s);
	q->g2s = NEW(int32_t, q->d->n_gene);
	for (i = 0; i < q->d->n_gene; ++i) q->g2s[i] = -1;
	for (i = 0; i < q->n_seg; ++i) q->g2s[q->seg[i].gid] = i;
}

static void flag_vtx(pg_graph_t *q)   /* graph.c:61-69 */
{
	int32_t j, i;
	for (j = 0; j < q->d->n_genome; ++j)
		for (i = 0; i < q->d->genome[j].n_hit; ++i) {
			pg_hit_t *a = &q->d->genome[j].hit[i];
			a->vtx = q->g2s[q->d->prot[a->pid].gid] >= 0;
		}
}

static int cmp_seg_gid(const void *a, const void *b) { int32_t x = ((const pg_seg_t*)a)->gid, y = ((const pg_seg_t*)b)->gid; return x < y? -1 : x > y; }

/* vertex.c:6-100. dom[j*n_gene+g]: NONE, DOM, or the gene that shadows g's only rank-0 hit in genome j; the
 * "already claimed" mark of the greedy pass is kept in a separate byte table. */
static void gen_vtx(const pg_opt_t *opt, pg_graph_t *q)
{
	const pg_data_t *d = q->d;
	const int32_t NONE = -1, DOM = -2;
	int32_t i, j, ng = d->n_gene;
	int32_t *dom = NEW(int32_t, (size_t)d->n_genome * ng);
	uint8_t *claimed = NEW0(uint8_t, (size_t)d->n_genome * ng);
	pg128_t *cnt = NEW(pg128_t, ng);
	pg_seg_t *seg = 0;
	int32_t n_seg = 0, m_seg = 0;
	for (i = 0; i < ng; ++i) cnt[i].x = (uint64_t)(int64_t)d->prot[d->gene[i].rep_pid].avg_score_adj, cnt[i].y = i;
	for (j = 0; j < d->n_genome; ++j) {
		const pg_genome_t *g = &d->genome[j];
		int32_t *row = &dom[(size_t)j * ng];
		uint8_t *sub = NEW0(uint8_t, ng);
		for (i = 0; i < ng; ++i) row[i] = NONE;
		for (i = 0; i < g->n_hit; ++i) {
			const pg_hit_t *a = &g->hit[i];
			int32_t gid;
			if (a->rank != 0 || a->flt) continue;
			gid = d->prot[a->pid].gid;
			if (!a->shadow) row[gid] = DOM;
			else { sub[gid] = 1; if (row[gid] == NONE) row[gid] = d->prot[a->pid_dom].gid; }
		}
		for (i = 0; i < ng; ++i) {
			if (row[i] == DOM) cnt[i].x += 1ULL << 32;
			else if (sub[i]) cnt[i].y += 1ULL << 32;
		}
		free(sub);
	}
	for (i = 0; i < ng; ++i) if (d->gene[i].preferred) cnt[i].x |= 1ULL << 63;
	pgo_radix_replay128(cnt, ng);   /* the greedy pass below is order-sensitive */
	for (i = ng - 1; i >= 0; --i) {
		int32_t n_dom = (int32_t)(cnt[i].x << 1 >> 33), n_sub = (int32_t)(cnt[i].y >> 32), gid = (int32_t)cnt[i].y, x = 0, y = 0;
		for (j = 0; j < d->n_genome; ++j)
			if (dom[(size_t)j * ng + gid] == DOM) ++x, y += claimed[(size_t)j * ng + gid];
		if (opt->flag & PG_F_WRITE_VTX_SEL)
			printf("g\t%s\t%d\t%d\t%d\t%d\t%c\t%c\n", d->gene[gid].name, (int32_t)cnt[i].x, x, y, n_sub, "NY"[d->gene[gid].included], "NY"[d->gene[gid].preferred]);
		if (d->gene[gid].included || (n_dom >= d->n_genome * opt->min_vertex_ratio && y < x)) {
			pg_seg_t *p;
			if (n_seg == m_seg) { m_seg = m_seg? m_seg * 2 : 256; seg = (pg_seg_t*)realloc(seg, sizeof(pg_seg_t) * m_seg); }
			p = &seg[n_seg++];
			memset(p, 0, sizeof(*p));
			p->gid = gid, p->n_dom = n_dom, p->n_sub = n_sub;
			if (x > 0)
				for (j = 0; j < d->n_genome; ++j) {
					int32_t t = dom[(size_t)j * ng + gid];
					if (t >= 0) claimed[(size_t)j * ng + t] = 1;
				}
		}
	}
	qsort(seg, n_seg, sizeof(pg_seg_t), cmp_seg_gid);   /* vertex.c:85-94: unique keys */
	free(q->seg);
	q->seg = seg? seg : NEW0(pg_seg_t, 1); q->n_seg = q->m_seg = n_seg;
	free(dom); free(claimed); free(cnt);
	gen_g2s(q);
}


/* The surviving hits of one genome in the reference's (contig, cm) array order (graph.c:103, branch.c:116).  Ours is the stable
 * one (ties keep (contig, cs) order); survivors that share one (contig, cm) are put in the order the reference's sort chain
 * leaves them in at this point (replay_contig_order, by_cm = 1).  In every campaign case with such ties (tools/fuzz_oracle.py)
 * the two orders were the same, so unlike the two replays above this one has never been observed to change a result. */
static int32_t *survivors_in_ref_cm_order(const view_t *v, int32_t *n_out)
{
	const pg_hit_t *h = v->g->hit;
	int32_t *sl = NEW(int32_t, v->n), n = 0, k, i0, i1, *rank = 0, rank_lo = -1, rank_hi = -1;
	for (k = 0; k < v->n; ++k) if (!h[v->by_cm[k]].flt && !h[v->by_cm[k]].shadow) sl[n++] = v->by_cm[k];
	for (i0 = 0; i0 < n && g_file_pos && v->j < g_file_pos_n && g_file_pos[v->j]; i0 = i1) {
		int32_t t, u;
		for (i1 = i0 + 1; i1 < n && h[sl[i1]].cid == h[sl[i0]].cid && h[sl[i1]].cm == h[sl[i0]].cm; ++i1) {}
		if (i1 - i0 < 2) continue;
		if (!(sl[i0] >= rank_lo && sl[i0] < rank_hi)) {
			int32_t lo, hi;
			for (lo = sl[i0]; lo > 0 && h[lo - 1].cid == h[sl[i0]].cid; --lo) {}
			for (hi = sl[i0]; hi < v->n && h[hi].cid == h[sl[i0]].cid; ++hi) {}
			free(rank); rank = NEW(int32_t, hi - lo); rank_lo = lo, rank_hi = hi;
			replay_contig_order(v->g, g_file_pos[v->j], lo, hi, g_sort_pairs, 1, rank);
		}
		for (t = i0 + 1; t < i1; ++t) {   /* insertion sort of the group by the replayed position */
			int32_t x = sl[t];
			for (u = t; u > i0 && rank[sl[u - 1] - rank_lo] > rank[x - rank_lo]; --u) sl[u] = sl[u - 1];
			sl[u] = x;
		}
	}
	free(rank);
	*n_out = n;
	return sl;
}

typedef struct { uint64_t x; int32_t genome, dist, s1, s2; } arcrec_t;

static int cmp_arcrec(const void *pa, const void *pb)
{
	const arcrec_t *a = (const arcrec_t*)pa, *b = (const arcrec_t*)pb;
	if (a->x != b->x) return a->x < b->x? -1 : 1;
	return a->genome < b->genome? -1 : a->genome > b->genome;
}

static inline int32_t arc_score(const pg_graph_t *q, const pg_hit_t *a, int ori)   /* graph.c:82-85 */
{
	return (ori || a->score_ori > a->score_dom || a->pid_dom0 < 0 || q->g2s[q->d->prot[a->pid_dom0].gid] >= 0)? a->score_ori : a->score_dom;
}

/* graph.c:87-177 */
static void gen_arc(const pg_opt_t *opt, pg_graph_t *q, view_t *views)
{
	pg_data_t *d = q->d;
	int32_t j, i, k, use_ori = !!(opt->flag & PG_F_ORI_FOR_BRANCH), check_strand = !!(opt->flag & PG_F_CHECK_STRAND);
	int64_t n_rec = 0, m_rec = 0, a0, a1;
	arcrec_t *rec = 0;
	int32_t *seg_cnt = NEW(int32_t, q->n_seg), *sl, n_sl;
	for (i = 0; i < q->n_seg; ++i) q->seg[i].n_genome = q->seg[i].tot_cnt = 0;
	for (j = 0; j < d->n_genome; ++j) {
		view_t *v = &views[j];
		const pg_hit_t *h = v->g->hit, *prev = 0;
		uint32_t pv = 0;
		sweep(opt, v, SWEEP_SHADOW, 0, check_strand);
		memset(seg_cnt, 0, sizeof(int32_t) * q->n_seg);
		sl = survivors_in_ref_cm_order(v, &n_sl);
		for (k = 0; k < n_sl; ++k) {   /* survivors in (contig, cm) order */
			const pg_hit_t *a = &h[sl[k]];
			uint32_t w;
			int32_t sid;
			sid = q->g2s[v->gid[sl[k]]];
			if (sid < 0) { fprintf(stderr, "[pg_oracle] surviving hit without a segment\n"); abort(); }   /* graph.c:111 */
			w = (uint32_t)sid << 1 | a->rev;
			++seg_cnt[sid];
			if (prev && prev->cid == a->cid) {
				int32_t sp = arc_score(q, prev, use_ori), sa = arc_score(q, a, use_ori);
				if (prev->cm == a->cm) ++g_tie[0];
				if (n_rec + 2 > m_rec) { m_rec = m_rec? m_rec * 2 : 1 << 16; rec = (arcrec_t*)realloc(rec, sizeof(arcrec_t) * m_rec); }
				rec[n_rec].x = (uint64_t)pv << 32 | w, rec[n_rec].genome = j, rec[n_rec].dist = (int32_t)(a->cm - prev->cm), rec[n_rec].s1 = sp, rec[n_rec].s2 = sa; ++n_rec;
				rec[n_rec].x = (uint64_t)(w ^ 1) << 32 | (pv ^ 1), rec[n_rec].genome = j, rec[n_rec].dist = (int32_t)(a->cm - prev->cm), rec[n_rec].s1 = sa, rec[n_rec].s2 = sp; ++n_rec;
			}
			prev = a, pv = w;
		}
		free(sl);
		for (i = 0; i < q->n_seg; ++i) q->seg[i].n_genome += seg_cnt[i] > 0, q->seg[i].tot_cnt += seg_cnt[i];
	}
	free(seg_cnt);
	qsort(rec, n_rec, sizeof(arcrec_t), cmp_arcrec);
	q->n_arc = 0;
	for (a0 = 0; a0 < n_rec; a0 = a1) {   /* one output arc per distinct key */
		int64_t b0, b1, sum_dist = 0, sum_s1 = 0, sum_s2 = 0;
		int32_t n_genome = 0, tot = 0;
		pg_arc_t *p;
		for (a1 = a0; a1 < n_rec && rec[a1].x == rec[a0].x; ++a1) {}
		for (b0 = a0; b0 < a1; b0 = b1) {   /* per genome: count, mean distance, max scores (graph.c:128-145) */
			uint64_t dsum = 0;
			int32_t m1 = 0, m2 = 0, cnt, dist;
			for (b1 = b0; b1 < a1 && rec[b1].genome == rec[b0].genome; ++b1) {
				dsum += (uint64_t)(int64_t)rec[b1].dist;
				m1 = m1 > rec[b1].s1? m1 : rec[b1].s1;
				m2 = m2 > rec[b1].s2? m2 : rec[b1].s2;
			}
			cnt = (int32_t)(b1 - b0);
			dist = (int32_t)((double)dsum / cnt + .499);
			++n_genome, tot += cnt;
			sum_dist += (int64_t)((uint64_t)(int64_t)dist * (uint64_t)(int64_t)cnt);
			sum_s1 += m1, sum_s2 += m2;
		}
		if (q->n_arc == q->m_arc) { q->m_arc = q->m_arc? q->m_arc * 2 : 1024; q->arc = (pg_arc_t*)realloc(q->arc, sizeof(pg_arc_t) * q->m_arc); }
		p = &q->arc[q->n_arc++];
		memset(p, 0, sizeof(*p));   /* graph.c:153-174 */
		p->x = rec[a0].x, p->n_genome = n_genome, p->tot_cnt = tot;
		p->avg_dist = (int32_t)(int64_t)((double)sum_dist / tot + .499);
		p->s1 = (int32_t)((double)sum_s1 / n_genome + .499);
		p->s2 = (int32_t)((double)sum_s2 / n_genome + .499);
	}
	free(rec);
}

static void arc_index(pg_graph_t *q)   /* graph.c:202-217 */
{
	int32_t i0, i;
	free(q->idx);
	q->idx = NEW0(uint64_t, (size_t)q->n_seg * 2);
	for (i0 = 0, i = 1; i <= q->n_arc; ++i)
		if (i == q->n_arc || q->arc[i].x >> 32 != q->arc[i0].x >> 32)
			q->idx[q->arc[i0].x >> 32] = (uint64_t)i0 << 32 | (uint32_t)(i - i0), i0 = i;
}

/* graph.c:237-263 (+ pg_hard_delete graph.c:219-235) */
static void flt_high_occ(int32_t max_avg_occ, int32_t max_degree, int32_t max_dist_loci, pg_graph_t *q)
{
	int32_t i, i0, k;
	for (i = 0; i < q->n_seg; ++i)
		if (q->seg[i].tot_cnt > max_avg_occ * q->d->n_genome) q->seg[i].del = 1;
	for (i0 = 0, i = 1; i <= q->n_arc; ++i)
		if (i == q->n_arc || q->arc[i].x >> 32 != q->arc[i0].x >> 32) {
			int32_t sid = (int32_t)(q->arc[i0].x >> 33);
			if (i - i0 > max_degree) q->seg[sid].del = 1;
			i0 = i;
		}
	for (i = 0; i < q->n_seg; ++i) {
		const pg_seg_t *s = &q->seg[i];
		int32_t m = s->n_dist_loci[0] > s->n_dist_loci[1]? s->n_dist_loci[0] : s->n_dist_loci[1];
		if (m > max_dist_loci) q->seg[i].del = 1;
	}
	for (i = k = 0; i < q->n_seg; ++i) if (!q->seg[i].del) q->seg[k++] = q->seg[i];
	q->n_seg = k;
	gen_g2s(q);
	flag_vtx(q);
}

typedef struct { int32_t cid, ord; int64_t cm; } genepos_t;   /* cid < 0: gene absent from this genome */

/* branch.c:31-46 */
static int32_t n_local(const pg_opt_t *opt, const pg_data_t *d, const genepos_t *pos, int32_t g1, int32_t g2)
{
	int32_t j, n = 0, frag = !!(opt->flag & PG_F_FRAG_MODE);
	for (j = 0; j < d->n_genome; ++j) {
		const genepos_t *a = &pos[(size_t)j * d->n_gene + g1], *b = &pos[(size_t)j * d->n_gene + g2];
		int64_t dd;
		int32_t c;
		if (a->cid < 0 || b->cid < 0) continue;
		if (!frag && a->cid != b->cid) continue;
		dd = a->cm - b->cm, c = a->ord - b->ord;
		if ((dd >= -opt->local_dist && dd <= opt->local_dist) || (c >= -opt->local_count && c <= opt->local_count)) ++n;
	}
	return n;
}

/* branch.c:6-29 + branch.c:48-106 */
static void mark_branch_arcs(const pg_opt_t *opt, pg_graph_t *q)
{
	const pg_data_t *d = q->d;
	int32_t j, i, k, ng = d->n_gene;
	uint32_t v;
	genepos_t *pos = NEW(genepos_t, (size_t)d->n_genome * ng);
	int32_t *top = NEW(int32_t, q->n_arc), *grp = NEW(int32_t, q->n_arc), *owner = NEW(int32_t, ng);
	for (j = 0; j < d->n_genome; ++j) {
		const pg_genome_t *g = &d->genome[j];
		genepos_t *row = &pos[(size_t)j * ng];
		int32_t r = 0, last = -1, n_tied = 0;
		for (i = 0; i < ng; ++i) row[i].cid = -1;
		for (i = 0; i < g->n_hit; ++i) {   /* (contig, cs) order: the LAST survivor of a gene wins */
			const pg_hit_t *b = &g->hit[i];
			if (b->shadow || b->flt) continue;
			if (last >= 0 && g->hit[last].cid == b->cid && g->hit[last].cs == b->cs) ++g_tie[1], ++n_tied;
			row[d->prot[b->pid].gid].cid = b->cid, row[d->prot[b->pid].gid].ord = r++, row[d->prot[b->pid].gid].cm = b->cm;
			owner[d->prot[b->pid].gid] = i;
			last = i;
		}
		/* Survivors with one (contig, cs): their ordinals follow the reference's array order at this point of its sort chain
		 * (branch.c:18-25 walks the array as left by the last pg_hit_sort(g, 0)): replay the chain for those contigs. */
		if (n_tied > 0 && g_file_pos && j < g_file_pos_n && g_file_pos[j]) {
			int32_t i0, i1, lo, hi, *rank = 0, rank_lo = -1, rank_hi = -1;
			for (i0 = 0, r = 0; i0 < g->n_hit; i0 = i1) {   /* groups of hits with one (contig, cs); r = survivors before the group */
				int32_t n_sv = 0, t, u;
				for (i1 = i0; i1 < g->n_hit && g->hit[i1].cid == g->hit[i0].cid && g->hit[i1].cs == g->hit[i0].cs; ++i1)
					n_sv += !g->hit[i1].shadow && !g->hit[i1].flt;
				if (n_sv >= 2) {
					int32_t *sv = NEW(int32_t, n_sv);
					if (!(i0 >= rank_lo && i0 < rank_hi)) {   /* a new contig: replay it */
						for (lo = i0; lo > 0 && g->hit[lo - 1].cid == g->hit[i0].cid; --lo) {}
						for (hi = i0; hi < g->n_hit && g->hit[hi].cid == g->hit[i0].cid; ++hi) {}
						free(rank); rank = NEW(int32_t, hi - lo); rank_lo = lo, rank_hi = hi;
						replay_contig_order(g, g_file_pos[j], lo, hi, g_sort_pairs, 0, rank);
					}
					for (t = i0, u = 0; t < i1; ++t) if (!g->hit[t].shadow && !g->hit[t].flt) sv[u++] = t;
					for (t = 1; t < n_sv; ++t) {   /* insertion sort by the replayed position */
						int32_t x = sv[t];
						for (u = t; u > 0 && rank[sv[u - 1] - rank_lo] > rank[x - rank_lo]; --u) sv[u] = sv[u - 1];
						sv[u] = x;
					}
					for (t = 0; t < n_sv; ++t)   /* the ordinals of the group, in the reference's order (a gene keeps its LAST survivor: only its owner is renumbered) */
						if (owner[d->prot[g->hit[sv[t]].pid].gid] == sv[t]) row[d->prot[g->hit[sv[t]].pid].gid].ord = r + t;
					free(sv);
				}
				r += n_sv;
			}
			free(rank);
		}
	}
	for (i = 0; i < q->n_seg; ++i) q->seg[i].n_dist_loci[0] = q->seg[i].n_dist_loci[1] = 0;
	for (v = 0; v < (uint32_t)q->n_seg * 2; ++v) {
		int32_t n = (int32_t)q->idx[v], n_top = 0, best = 0, n_grp = 0;
		pg_arc_t *a = &q->arc[q->idx[v] >> 32];
		if (n < 2) continue;
		for (i = 0; i < n; ++i) best = best > a[i].s1? best : a[i].s1;
		for (i = 0; i < n; ++i) if (a[i].s1 == best) top[n_top++] = q->seg[(uint32_t)a[i].x >> 1].gid;
		for (i = 0; i < n; ++i) {
			double r = 1.0 - (double)a[i].s1 / best;
			if (r > opt->branch_diff) {
				int32_t nl = 0, gid = q->seg[(uint32_t)a[i].x >> 1].gid;
				for (k = 0; k < n_top; ++k) nl += n_local(opt, d, pos, top[k], gid);
				a[i].weak_br = ((nl == 0 && r > opt->branch_diff_dist) || r > opt->branch_diff_cut)? 2 : 1;
			}
		}
		memset(grp, 0, sizeof(int32_t) * n);   /* branch.c:82-90: greedy grouping of neighbours by locality */
		for (i = 0; i < n; ++i) {
			int32_t gi = q->seg[(uint32_t)a[i].x >> 1].gid;
			if (grp[i] == 0) grp[i] = ++n_grp;
			for (k = i + 1; k < n; ++k)
				if (grp[k] == 0 && n_local(opt, d, pos, gi, q->seg[(uint32_t)a[k].x >> 1].gid) > 0) grp[k] = grp[i];
		}
		q->seg[v >> 1].n_dist_loci[v & 1] = n_grp;
	}
	free(pos); free(top); free(grp); free(owner);
}

static const pg_arc_t *find_arc(const pg_graph_t *q, uint32_t v, uint32_t w)   /* pgpriv.h:99-107 */
{
	int32_t i, n = (int32_t)q->idx[v];
	const pg_arc_t *a = &q->arc[q->idx[v] >> 32];
	for (i = 0; i < n; ++i) if ((uint32_t)a[i].x == w) return &a[i];
	return 0;
}

/* branch.c:108-145 */
static void mark_branch_hits(pg_graph_t *q, view_t *views)
{
	int32_t j, k;
	for (j = 0; j < q->d->n_genome; ++j) {
		view_t *v = &views[j];
		pg_hit_t *h = v->g->hit, *prev = 0;
		uint32_t pv = 0;
		int32_t n_sl, *sl = survivors_in_ref_cm_order(v, &n_sl);
		for (k = 0; k < n_sl; ++k) {
			pg_hit_t *a = &h[sl[k]];
			const pg_arc_t *e;
			uint32_t w;
			w = (uint32_t)q->g2s[v->gid[sl[k]]] << 1 | a->rev;
			if (prev && prev->cid == a->cid) {
				if ((e = find_arc(q, pv, w)) != 0 && e->weak_br > prev->weak_br) prev->weak_br = e->weak_br;
				if ((e = find_arc(q, w ^ 1, pv ^ 1)) != 0 && e->weak_br > a->weak_br) a->weak_br = e->weak_br;
			}
			prev = a, pv = w;
		}
		free(sl);
	}
}

void pgo_graph_gen(const pg_opt_t *opt, pg_graph_t *q)   /* graph.c:280-322 */
{
	pg_data_t *d = q->d;
	int32_t i, j, k;
	view_t *views = NEW(view_t, d->n_genome);
	g_sort_pairs = 0;
	SET_FILTER(d, a->pseudo);
	gen_vtx(opt, q);
	flag_vtx(q);
	SET_FILTER(d, !a->vtx);
	for (j = 0; j < d->n_genome; ++j) view_init(&views[j], d, j);
	gen_arc(opt, q, views); ++g_sort_pairs;                   /* graph.c:103,123 */
	flt_high_occ(opt->max_avg_occ * 2, opt->max_degree * 2, opt->max_dist_loci, q);
	SET_FILTER(d, !a->vtx);
	gen_arc(opt, q, views); ++g_sort_pairs;
	for (i = 0; i < opt->n_branch_flt; ++i) {
		double r = 1.0 + (double)(opt->n_branch_flt - 1 - i) / opt->n_branch_flt;
		int32_t max_avg_occ = (int32_t)(opt->max_avg_occ * r + .499);
		int32_t max_degree = (int32_t)(opt->max_degree * r + .499);
		int32_t max_dist_loci = (int32_t)(opt->max_dist_loci * r + .499);
		arc_index(q);
		mark_branch_arcs(opt, q);
		mark_branch_hits(q, views); ++g_sort_pairs;               /* branch.c:116,140 */
		SET_FILTER(d, a->weak_br == 2);
		if (i > 0) {
			flt_high_occ(max_avg_occ, max_degree, max_dist_loci, q);
			SET_FILTER(d, !a->vtx);
		}
		gen_arc(opt, q, views); ++g_sort_pairs;
	}
	SET_FILTER(d, a->shadow);
	if (opt->min_arc_cnt > 1) {   /* graph.c:179-200 */
		for (i = k = 0; i < q->n_arc; ++i)
			if (q->arc[i].n_genome >= opt->min_arc_cnt) q->arc[k++] = q->arc[i];
		q->n_arc = k;
	}
	arc_index(q);
	for (j = 0; j < d->n_genome; ++j) view_free(&views[j]);
	free(views);
}
