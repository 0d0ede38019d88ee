// pgx_round.cuh -- the round body of pg_graph_gen (graph.c:301-315) without host round trips.
//
// Round 1 showed stage 2 (17 x pg_gen_arc + 15 branch rounds) to be latency-bound: ~45 launches and two host read-backs per round for
// <1 ms of HBM work, the largest piece a 4-pass radix sort of ~3.7 M arc records per round.  Here every size that changes from round
// to round stays in DevCounts, every launch is capacity-sized or grid-stride, and the arc multiset is built WITHOUT sorting the records:
//
//   pg_gen_arc (graph.c:87-177) =  k_compact_survivors   survivors in (contig, cm) order, one pass (flag + decoupled look-back scan + compaction)
//                                  k_arc_insert          adjacent survivor pairs -> the two arc keys (graph.c:115-120) -> open-addressing table of
//                                                        the DISTINCT keys (64-bit CAS); per (key, genome) a two-bit "seen once / seen again" mark
//                                  k_arc_fold            a (key, genome) group with ONE record -- almost all of them -- is its own per-genome
//                                                        collapse (graph.c:128-145: n = 1, dist = d, max scores = the scores): integer atomics
//                                                        into the key's sums; records of larger groups are grouped by (key, genome) in a second
//                                                        small table (sum, count, max)
//                                  k_arc_ovf_fold        each such group is collapsed with the reference's rounded mean and added to the key's sums
//                                  [N ranks: k_arc_pack -> one allgather -> k_arc_merge: the per-rank sums are integers, adding them is exact]
//                                  k_vertex_scan         out-degree -> arc offsets (and the offsets of the n_local pair table), one pass
//                                  k_arc_place/k_arc_rank counting sort by source vertex, rank by target inside a vertex: arc[] sorted by key
//                                                        (graph.c:151) and idx[] (graph.c:202-211) at once; the three double divisions last
//   branch round (branch.c)     =  k_compact_survivors (cs order: ordinals) k_pos_last k_pos_fill k_cs_tie_mark k_branch_local [allreduce]
//                                  k_branch_decide_fast k_branch_hits_fast k_seg_filter k_flag_vtx_fast
//
// Anything the fixed-capacity tables cannot hold, and the (rare) rounds in which the reference's history-dependent order of cs-tied
// survivors decides a locality test (DESIGN.md "tie order"), raise DevCounts.fallback: the engine then redoes the run on the
// round-1 path (Session::stage_graph, careful mode), which handles both.  The result is bit-identical either way.
#pragma once
#include "pgx_util.cuh"
#include "pgx_scan.cuh"
#include "pgx_graphtypes.cuh"
#include "pangene_abi.h"

namespace pgx {

// [emu-begin round]
constexpr uint32_t ARC_NONE = 0xffffffffu;
constexpr unsigned long long KEY_EMPTY = ~0ull;

// ---- small building blocks ---------------------------------------------------------------------------------------------
// exclusive scan of one value per thread over a CTA of B threads (B <= 1024); wsum: 33 words of shared memory
template<int B> __device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t &total, uint32_t *wsum)
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
	if (lane == 31) wsum[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		const uint32_t w = lane < B / 32? wsum[lane] : 0u;
		uint32_t ws = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, ws, d); if (lane >= d) ws += t; }
		if (lane < B / 32) wsum[lane] = ws - w;
		if (lane == 31) wsum[32] = ws;
	}
	__syncthreads();
	total = wsum[32];
	const uint32_t r = wsum[warp] + inc - v;
	__syncthreads();                                  // wsum is free again when this returns
	return r;
}

// decoupled look-back (same protocol as k_scan_sum_lookback): called by the 32 lanes of warp 0; returns the sum of all tiles before `tile`
__device__ __forceinline__ uint32_t lookback_prefix(unsigned long long *status, uint32_t tile, uint32_t tile_sum)
{
	const int lane = threadIdx.x & 31;
	if (lane == 0) { __threadfence(); atomicExch(&status[tile], (tile == 0? LB_PREFIX : LB_AGG) | tile_sum); }
	uint32_t prefix = 0;
	if (tile > 0) {
		int64_t look = (int64_t)tile - 1;
		for (;;) {
			const int64_t p = look - lane;
			unsigned long long w = LB_PREFIX;
			if (p >= 0) { do { w = *(volatile unsigned long long*)&status[p]; } while ((w & LB_MASK) == 0); }
			const uint32_t closed = __ballot_sync(0xffffffffu, (w & LB_MASK) == LB_PREFIX);
			const int first = closed? __ffs(closed) - 1 : 32;
			uint32_t add = lane <= first? (uint32_t)(w & 0xffffffffull) : 0u;
#pragma unroll
			for (int sft = 16; sft > 0; sft >>= 1) add += __shfl_xor_sync(0xffffffffu, add, sft);
			prefix += add;
			if (closed) break;
			look -= 32;
		}
		if (lane == 0) { __threadfence(); atomicExch(&status[tile], LB_PREFIX | (unsigned long long)(prefix + tile_sum)); }
	}
	return prefix;
}

// one launch zeroes / fills everything a round needs fresh (instead of a dozen memset nodes)
struct ClearList { void *p[16]; unsigned long long n16[16]; uint32_t fill[16]; int n; };   // regions of n16 16-byte words
static inline void clear_add(ClearList &c, void *p, size_t bytes, uint32_t fill)
{
	if (c.n >= 16 || ((uintptr_t)p & 15)) { fprintf(stderr, "[pgx] clear list: bad region\n"); abort(); }
	c.p[c.n] = p; c.n16[c.n] = (bytes + 15) / 16; c.fill[c.n] = fill; ++c.n;
}
__global__ void k_round_clear(ClearList c)
{
	const unsigned long long g0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, gs = (unsigned long long)gridDim.x * blockDim.x;
	for (int r = 0; r < c.n; ++r) {
		uint4 v; v.x = v.y = v.z = v.w = c.fill[r];
		uint4 *q = (uint4*)c.p[r];
		for (unsigned long long i = g0; i < c.n16[r]; i += gs) q[i] = v;
	}
}

// ---- survivors: flag -> exclusive scan -> compaction in ONE pass ------------------------------------------------------------
// element k of the scan is hit perm[k] (perm == nullptr: hit k); it survives if it is neither filtered nor shadowed (graph.c:109,
// branch.c:20).  pos (optional, n + 1 entries) receives the exclusive scan, list the surviving hit indices in scan order, *count_out
// the total.  status / ticket: zeroed before the launch.  reset (optional): DevCounts fields that the kernels after this one bump.
__global__ void __launch_bounds__(SCAN_B) k_compact_survivors(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ perm, int32_t n,
	uint32_t *pos, uint32_t *list, uint32_t *count_out, unsigned long long *status, uint32_t *ticket, DevCounts *reset)
{
	__shared__ uint32_t wsum[33];
	__shared__ uint32_t s_tile, s_prefix;
	const int tid = threadIdx.x;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	if (reset && tile == 0 && tid == 0) { reset->n_part = 0; reset->n_ovf = 0; reset->n_l2 = 0; reset->n_gpart = 0; }
	const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)tid * SCAN_I;
	uint32_t idx[SCAN_I], fl = 0, sum = 0;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) {
		const int64_t i = base + k;
		idx[k] = 0;
		if (i < n) {
			idx[k] = perm? perm[i] : (uint32_t)i;
			if (!(bits[idx[k]] & (PG_HB_FLT | PG_HB_SHADOW))) { fl |= 1u << k; ++sum; }
		}
	}
	uint32_t tile_sum;
	const uint32_t excl = block_exclusive_scan<SCAN_B>(sum, tile_sum, wsum);
	if (tid < 32) { const uint32_t p = lookback_prefix(status, tile, tile_sum); if (tid == 0) s_prefix = p; }
	__syncthreads();
	uint32_t run = s_prefix + excl;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) {
		const int64_t i = base + k;
		if (pos && i <= n) pos[i] = run;
		if (i == n) *count_out = run;
		if (fl >> k & 1u) list[run++] = idx[k];
	}
}

// ---- the arc table ------------------------------------------------------------------------------------------------------
// Every adjacency yields the arc (v -> w) and its mirror (w^1 -> v^1) with the two scores swapped (graph.c:115-120), always together, so the
// whole multiset is mirror-symmetric: only the CANONICAL key of a pair -- the smaller of the two -- is stored and summed; the mirror arc
// is written out from the same sums when the table is turned into arc[] (same n_genome, tot_cnt, avg_dist; s1 and s2 exchanged).  An arc
// that is its own mirror (w == v^1) receives both records of its pair: a (key, genome) group of two, handled by the overflow path.
constexpr int ACC_SHIFT = 38;                                    // packed sums: [count : 26 bits][score sum : 38 bits] (the engine checks the bounds)
constexpr unsigned long long ACC_MASK = (1ull << ACC_SHIFT) - 1;
struct ArcAcc { unsigned long long a, b, dist; };              // a = n_genome << 38 | sum_s1, b = tot_cnt << 38 | sum_s2, dist = sum(dist * n); zero at the start of a round
struct ArcTable {
	unsigned long long *key;     // [H] canonical arc key x = v << 32 | w, KEY_EMPTY when free
	ArcAcc *acc;                 // [H]
	uint32_t *list;              // [H] slots in creation order
	uint32_t *bm1, *bm2;         // [H * Gl bits] (key, local genome): seen once / seen more than once
	uint32_t H;                  // power of two
};
struct SvInfo { uint32_t w; int32_t sc, cm, ctg; };             // a survivor as the arc walk sees it: oriented vertex (graph.c:112), score (graph.c:82-85), position, contig
constexpr uint32_t REC_SWAP = 0x40000000u, REC_SELF = 0x80000000u, REC_SLOT = 0x3fffffffu;   // pair record: slot of the canonical key | canonical is the mirror | arc is its own mirror
struct OvfTable { unsigned long long *key; uint32_t *n; unsigned long long *dsum; int32_t *m1, *m2; uint32_t *list; uint32_t H; };   // groups (slot, genome) of the overflow records

__device__ __forceinline__ uint32_t hash_u64(unsigned long long k)
{
	k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
	return (uint32_t)k;
}
__device__ __forceinline__ unsigned long long arc_mirror(unsigned long long x) { return (unsigned long long)((uint32_t)x ^ 1u) << 32 | ((uint32_t)(x >> 32) ^ 1u); }

// slot of canonical key k, created if new.  n_created: running count of distinct keys (list position); deg (optional): out-degrees, bumped
// for the key's source vertex and for its mirror's
__device__ __forceinline__ uint32_t arc_slot(const ArcTable &T, unsigned long long k, uint32_t *n_created, uint32_t *deg, uint32_t *fallback)
{
	uint32_t i = hash_u64(k) & (T.H - 1);
	for (uint32_t probe = 0; probe < T.H; ++probe, i = (i + 1) & (T.H - 1)) {
		unsigned long long cur = *(volatile unsigned long long*)&T.key[i];
		if (cur == k) return i;
		if (cur == KEY_EMPTY) {
			cur = atomicCAS(&T.key[i], KEY_EMPTY, k);
			if (cur == KEY_EMPTY) {
				const uint32_t id = atomicAdd(n_created, 1u);
				if (id < T.H) T.list[id] = i;
				if (2ull * id >= T.H) *fallback = 1u;      // load factor 1/2: the engine sized the table from the previous round with a wide margin
				if (deg) {
					const unsigned long long m = arc_mirror(k);
					atomicAdd(&deg[(uint32_t)(k >> 32)], 1u);
					if (m != k) atomicAdd(&deg[(uint32_t)(m >> 32)], 1u);
				}
				return i;
			}
			if (cur == k) return i;
		}
	}
	*fallback = 1u;
	return 0;
}
// read-only look-up; ARC_NONE if absent
__device__ __forceinline__ uint32_t arc_find(const ArcTable &T, unsigned long long k)
{
	uint32_t i = hash_u64(k) & (T.H - 1);
	for (uint32_t probe = 0; probe < T.H; ++probe, i = (i + 1) & (T.H - 1)) {
		const unsigned long long cur = T.key[i];
		if (cur == k) return i;
		if (cur == KEY_EMPTY) return ARC_NONE;
	}
	return ARC_NONE;
}

struct ArcSrc {      // the arc walk's input: survivors in (contig, cm) order (indices into the active columns) and those columns
	const uint32_t *sv;
	const int32_t *ctg, *cm, *gidp, *sori, *sdom, *pdom0; const uint32_t *bits;
	const int32_t *g2s, *prot_gid, *ctg2g;
	int32_t use_ori, n_seg0, Gl;
};

// one thread per survivor: everything the walk needs of it in 16 contiguous bytes, and the segment occupancy (graph.c:113,125-126)
__global__ void k_survivor_info(ArcSrc e, const DevCounts *__restrict__ dc, SvInfo *info, uint32_t *seg_seen, int32_t *seg_ng, int32_t *seg_tot, uint32_t *err)
{
	const int32_t m = (int32_t)dc->m_sv;
	PGX_GS(t, m) {
		const uint32_t h = e.sv[t];
		const int32_t c = e.ctg[h], so = e.sori[h], sd = e.sdom[h], p0 = e.pdom0[h];
		const uint32_t b = e.bits[h];
		int32_t sid = e.g2s[e.gidp[h] >> 1];
		if (sid < 0) { atomicOr(err, 32u); sid = 0; }                               // graph.c:111 assert
		SvInfo r;
		r.w = (uint32_t)sid << 1 | (b & PG_HB_REV); r.cm = e.cm[h]; r.ctg = c;
		r.sc = (e.use_ori || so > sd || p0 < 0 || e.g2s[e.prot_gid[p0]] >= 0)? so : sd;   // graph.c:82-85
		info[t] = r;
		const size_t bit = (size_t)e.ctg2g[c] * e.n_seg0 + sid;
		const uint32_t mk = 1u << (bit & 31);
		if (!(atomicOr(&seg_seen[bit >> 5], mk) & mk)) atomicAdd(&seg_ng[sid], 1);
		atomicAdd(&seg_tot[sid], 1);
	}
}

// graph.c:106-122: survivor t and its predecessor on the same contig give (v -> w) and (w^1 -> v^1); here: the canonical one of the two
__global__ void k_arc_insert(const SvInfo *__restrict__ info, const int32_t *__restrict__ ctg2g, int32_t Gl, ArcTable T, DevCounts *dc, uint32_t *rec, uint32_t *deg, uint32_t *tie_cnt)
{
	const int32_t m = (int32_t)dc->m_sv;
	PGX_GS(t, m) {
		uint32_t r = ARC_NONE;
		if (t > 0) {
			const SvInfo a = info[t - 1], b = info[t];
			if (a.ctg == b.ctg) {
				if (b.cm == a.cm) { atomicAdd(tie_cnt, 1u); dc->fallback = 1u; }             // two survivors with one (contig, cm): the reference's order of the two decides (careful path)
				const unsigned long long k0 = (unsigned long long)a.w << 32 | b.w, k1 = arc_mirror(k0);
				const unsigned long long kc = k0 < k1? k0 : k1;
				const uint32_t slot = arc_slot(T, kc, &dc->n_part, deg, &dc->fallback);
				r = slot | (k1 < k0? REC_SWAP : 0u) | (k1 == k0? REC_SELF : 0u);
				const size_t bit = (size_t)slot * Gl + ctg2g[b.ctg];
				const uint32_t mk = 1u << (bit & 31);
				if (k1 == k0) { atomicOr(&T.bm1[bit >> 5], mk); atomicOr(&T.bm2[bit >> 5], mk); }   // two records of one (key, genome) group
				else if (atomicOr(&T.bm1[bit >> 5], mk) & mk) atomicOr(&T.bm2[bit >> 5], mk);
			}
		}
		rec[t] = r;
	}
}

// A record of a (key, genome) group with several records (tandem duplicates: ~1 % of the records): the group is found -- or created -- in the
// second, small table and the record is added to its sum / count / maxima (graph.c:131-135); k_arc_ovf_fold collapses the groups afterwards.
__device__ __forceinline__ void ovf_group_add(const OvfTable &L, DevCounts *dc, uint32_t slot, int32_t gl, int32_t d, int32_t a1, int32_t a2)
{
	const unsigned long long k = (unsigned long long)slot << 32 | (uint32_t)gl;
	uint32_t j = hash_u64(k) & (L.H - 1), found = ARC_NONE;
	for (uint32_t probe = 0; probe < L.H; ++probe, j = (j + 1) & (L.H - 1)) {
		unsigned long long cur = *(volatile unsigned long long*)&L.key[j];
		if (cur == k) { found = j; break; }
		if (cur == KEY_EMPTY) {
			cur = atomicCAS(&L.key[j], KEY_EMPTY, k);
			if (cur == KEY_EMPTY) {
				const uint32_t id = atomicAdd(&dc->n_l2, 1u);
				if (id < L.H) L.list[id] = j;
				if (2ull * id >= L.H) dc->fallback = 1u;
				found = j; break;
			}
			if (cur == k) { found = j; break; }
		}
	}
	if (found == ARC_NONE) { dc->fallback = 1u; return; }
	atomicAdd(&L.n[found], 1u);
	atomicAdd(&L.dsum[found], (unsigned long long)(long long)d);                    // graph.c:133: uint64 sum of sign-extended distances
	atomicMax(&L.m1[found], a1); atomicMax(&L.m2[found], a2);                       // graph.c:131,134-135: maxima start at 0
}

// graph.c:128-145 for the groups of one record, graph.c:153-163 (the additive part) for all of them; the records of larger groups go straight
// into the overflow table (no list of them, no launch of its own)
__global__ void k_arc_fold(const SvInfo *__restrict__ info, const uint32_t *__restrict__ rec, const int32_t *__restrict__ ctg2g, ArcTable T, DevCounts *dc, int32_t Gl, OvfTable L)
{
	const int32_t m = (int32_t)dc->m_sv;
	PGX_GS(t, m) {
		const uint32_t r = rec[t];
		if (r == ARC_NONE) continue;
		const SvInfo a = info[t - 1], b = info[t];
		const uint32_t slot = r & REC_SLOT;
		const int32_t d = b.cm - a.cm, gl = ctg2g[b.ctg];
		const int32_t s1 = (r & REC_SWAP)? b.sc : a.sc, s2 = (r & REC_SWAP)? a.sc : b.sc;   // graph.c:116,118: (s1, s2) of the canonical record
		const size_t bit = (size_t)slot * Gl + gl;
		if (T.bm2[bit >> 5] >> (bit & 31) & 1u) {
			atomicAdd(&dc->n_ovf, (r & REC_SELF)? 2u : 1u);                              // (statistics only)
			ovf_group_add(L, dc, slot, gl, d, s1, s2);
			if (r & REC_SELF) ovf_group_add(L, dc, slot, gl, d, s2, s1);                  // an arc that is its own mirror: both records of the pair
		} else {
			// one record: n = 1, dist = (int32)((double)d / 1 + .499) = d (d >= 0: survivors are in cm order), max scores start at 0 (graph.c:131)
			ArcAcc *acc = &T.acc[slot];
			atomicAdd(&acc->a, 1ull << ACC_SHIFT | (unsigned long long)(s1 > 0? s1 : 0));
			atomicAdd(&acc->b, 1ull << ACC_SHIFT | (unsigned long long)(s2 > 0? s2 : 0));
			if (d) atomicAdd(&acc->dist, (unsigned long long)(long long)d);
		}
	}
}

// one thread per (key, genome) group of the overflow: the reference's per-genome collapse, then into the key's sums; the entry is
// handed back empty, so the table needs no clearing between rounds
__global__ void k_arc_ovf_fold(OvfTable L, ArcTable T, DevCounts *dc)
{
	const uint32_t n = dc->n_l2 < L.H? dc->n_l2 : L.H;
	PGX_GS(i, n) {
		const uint32_t j = L.list[i];
		const uint32_t slot = (uint32_t)(L.key[j] >> 32), cnt = L.n[j];
		const unsigned long long dsum = L.dsum[j];
		const int32_t m1 = L.m1[j], m2 = L.m2[j];
		const int32_t dg = (int32_t)((double)dsum / cnt + .499);                      // graph.c:141
		ArcAcc *acc = &T.acc[slot];
		atomicAdd(&acc->a, 1ull << ACC_SHIFT | (unsigned long long)m1);
		atomicAdd(&acc->b, (unsigned long long)cnt << ACC_SHIFT | (unsigned long long)m2);
		atomicAdd(&acc->dist, (unsigned long long)(long long)dg * (unsigned long long)cnt);   // graph.c:160
		L.key[j] = KEY_EMPTY; L.n[j] = 0; L.dsum[j] = 0; L.m1[j] = 0; L.m2[j] = 0;
	}
}

__device__ __forceinline__ ArcPartial acc_to_partial(unsigned long long x, const ArcAcc &c)
{
	ArcPartial p;
	p.x = x; p.n_genome = (int32_t)(c.a >> ACC_SHIFT); p.tot_cnt = (int32_t)(c.b >> ACC_SHIFT);
	p.sum_dist = (int64_t)c.dist; p.sum_s1 = (int64_t)(c.a & ACC_MASK); p.sum_s2 = (int64_t)(c.b & ACC_MASK);
	return p;
}

// ---- N ranks: every rank's table travels as [header, records...]; the sums are integers, so adding them in any order is exact -------
struct ArcWire { unsigned long long x, a, b, dist; };          // header: x = number of records, a = the sender's fallback flag
__global__ void k_arc_pack(ArcTable T, DevCounts *dc, ArcWire *send, uint32_t cap)
{
	uint32_t n = dc->n_part;
	if (n > cap) { n = cap; if (blockIdx.x == 0 && threadIdx.x == 0) dc->fallback = 1u; }
	if (blockIdx.x == 0 && threadIdx.x == 0) { ArcWire h; h.x = n; h.a = (dc->fallback || dc->n_part > cap)? 1 : 0; h.b = h.dist = 0; send[0] = h; }
	PGX_GS(i, n) { const uint32_t s = T.list[i]; const ArcAcc c = T.acc[s]; ArcWire w; w.x = T.key[s]; w.a = c.a; w.b = c.b; w.dist = c.dist; send[1 + i] = w; }
}
__global__ void k_arc_merge(const ArcWire *__restrict__ all, int32_t nranks, uint32_t cap, ArcTable G, DevCounts *dc, uint32_t *deg)
{
	const unsigned long long total = (unsigned long long)nranks * cap;
	for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (unsigned long long)gridDim.x * blockDim.x) {
		const uint32_t r = (uint32_t)(j / cap), i = (uint32_t)(j % cap);
		const ArcWire *blk = all + (size_t)r * (cap + 1);
		if (i == 0 && blk[0].a) dc->fallback = 1u;
		if (i >= (uint32_t)blk[0].x) continue;
		const ArcWire p = blk[1 + i];
		ArcAcc *acc = &G.acc[arc_slot(G, p.x, &dc->n_gpart, deg, &dc->fallback)];
		atomicAdd(&acc->a, p.a); atomicAdd(&acc->b, p.b);
		if (p.dist) atomicAdd(&acc->dist, p.dist);
	}
}

// ---- sorted arc[] and idx[] from the table: counting sort by source vertex, rank by key inside a vertex --------------------------------
// deg[NV] -> off[NV + 1] (arc offsets; total = n_arc) and nl_off[NV + 1] (offsets into the table of arc pairs (i <= k) of every vertex with
// >= 2 out-arcs, branch.c:60-98; total = n_nl), one pass with two look-back chains.  status: 2 * tiles words, zeroed; ticket zeroed.
__global__ void __launch_bounds__(SCAN_B) k_vertex_scan(const uint32_t *__restrict__ deg, int32_t NV, uint32_t *off, uint32_t *nl_off, DevCounts *dc,
	uint32_t arc_cap, unsigned long long *nl_total, unsigned long long *status, uint32_t *ticket)
{
	__shared__ uint32_t wsum[33];
	__shared__ uint32_t s_tile, s_pa, s_pb;
	const int tid = threadIdx.x;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile, n_tiles = gridDim.x;
	const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)tid * SCAN_I;
	uint32_t a[SCAN_I], b[SCAN_I], sa = 0, sb = 0;
	unsigned long long sb64 = 0;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) {
		const int64_t i = base + k;
		a[k] = i < NV? deg[i] : 0u;
		b[k] = a[k] >= 2u? a[k] * (a[k] + 1u) / 2u : 0u;
		if (a[k] >= 65536u) { b[k] = 0u; dc->fallback = 1u; }                        // (a vertex with 65 k out-arcs: leave it to the careful path)
		sa += a[k]; sb += b[k]; sb64 += b[k];
	}
	if (sb64) atomicAdd(nl_total, sb64);                                              // the 32-bit offsets below may wrap: consumers check this total first
	uint32_t ta, tb;
	const uint32_t ea = block_exclusive_scan<SCAN_B>(sa, ta, wsum);
	const uint32_t eb = block_exclusive_scan<SCAN_B>(sb, tb, wsum);
	if (tid < 32) {
		const uint32_t pa = lookback_prefix(status, tile, ta);
		const uint32_t pb = lookback_prefix(status + n_tiles, tile, tb);
		if (tid == 0) { s_pa = pa; s_pb = pb; }
	}
	__syncthreads();
	uint32_t ra = s_pa + ea, rb = s_pb + eb;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) {
		const int64_t i = base + k;
		if (i <= NV) { off[i] = ra; nl_off[i] = rb; }
		if (i == NV) {
			dc->n_arc = ra; dc->n_nl = rb;
			if (ra > arc_cap) { dc->fallback = 1u; dc->n_arc = 0; }                      // more arcs than the arrays hold: nothing is placed
		}
		ra += a[k]; rb += b[k];
	}
}

// every canonical key gives its arc and the mirror arc (scores exchanged); tmp.bits remembers where it came from: slot << 1 | is_mirror
__global__ void k_arc_place(ArcTable T, const uint32_t *__restrict__ n_created, const uint32_t *__restrict__ off, uint32_t *cur, const DevCounts *__restrict__ dc, ArcDev *tmp)
{
	if (dc->n_arc == 0) return;                           // nothing, or more than the arrays hold (fallback raised by k_vertex_scan)
	const uint32_t n = *n_created < T.H? *n_created : T.H;
	PGX_GS(i, n) {
		const uint32_t s = T.list[i];
		const unsigned long long x = T.key[s], xm = arc_mirror(x);
		ArcDev a = arc_finalize(acc_to_partial(x, T.acc[s]));
		a.bits = s << 1;
		const uint32_t v = (uint32_t)(x >> 32);
		tmp[off[v] + atomicAdd(&cur[v], 1u)] = a;
		if (xm != x) {
			const uint32_t vm = (uint32_t)(xm >> 32);
			const int32_t t = a.s1; a.s1 = a.s2; a.s2 = t; a.x = xm; a.bits = s << 1 | 1u;
			tmp[off[vm] + atomicAdd(&cur[vm], 1u)] = a;
		}
	}
}
// pos_of[slot * 2 + is_mirror] = where the arc ended up (k_branch_hits_fast finds an adjacency's arcs through it); the same launch stores
// the segment occupancy of this arc build (graph.c:125-126; summed over ranks by then) -- independent work, one launch less per round
__global__ void k_arc_rank(const ArcDev *__restrict__ tmp, const uint32_t *__restrict__ off, const DevCounts *__restrict__ dc, ArcDev *arc, unsigned long long *idx, uint32_t *pos_of,
	SegDev *seg, const int32_t *__restrict__ ng, const int32_t *__restrict__ tot)
{
	PGX_GS(p, dc->n_arc) {
		ArcDev a = tmp[p];
		const uint32_t v = (uint32_t)(a.x >> 32), b = off[v], e = off[v + 1];
		uint32_t rank = 0;
		for (uint32_t q = b; q < e; ++q) rank += tmp[q].x < a.x? 1u : 0u;                // keys are distinct
		pos_of[a.bits] = b + rank;
		a.bits = 0;
		arc[b + rank] = a;
		if (rank == 0) idx[v] = (unsigned long long)b << 32 | (e - b);                  // graph.c:202-211
	}
	if (seg) { PGX_GS(s, dc->n_seg) { seg[s].n_genome = ng[s]; seg[s].tot_cnt = tot[s]; } }
}
// ---- which tiles of the next overlap sweep can a changed hit influence? ---------------------------------------------------------------
// A hit's sweep result (shadow bit, pid_dom; overlap.c:101-178) depends on its own flag word and on the flag words of its window partners,
// nothing else that changes after the load-time block.  When hit h changes its flt / weak_br bits, the hits that have h in their window
// are: later ones that start before h ends, earlier ones that end after h starts -- an index range, because cs and the running maximum of
// ce are monotone inside a contig.  The range (in 256-hit sweep tiles) depends on coordinates only: k_reach computes it once per active
// set; a changed hit marks its tiles; k_sweep_tiled<SW_SHADOW> skips every unmarked tile.
struct DirtyCols { const int2 *reach; uint32_t *tiles; };
__global__ void k_reach(const int32_t *__restrict__ ctg, const int32_t *__restrict__ cs, const int32_t *__restrict__ ce, const int32_t *__restrict__ pmax, int32_t n, int32_t tile, int2 *reach)
{
	PGX_GS(h, n) {
		const int32_t c = ctg[h], hs = cs[h], he = ce[h];
		int32_t lo = 0, hi = h;                                 // contig start: first index with ctg >= c (ctg is non-decreasing)
		while (lo < hi) { const int32_t m = (lo + hi) >> 1; if (ctg[m] < c) lo = m + 1; else hi = m; }
		int32_t first = lo; hi = h;                             // first hit of the contig whose running max of ce exceeds my start
		while (first < hi) { const int32_t m = (first + hi) >> 1; if (pmax[m] > hs) hi = m; else first = m + 1; }
		lo = h + 1; hi = n;                                     // first index past h that is on another contig or starts at/after my end
		while (lo < hi) { const int32_t m = (lo + hi) >> 1; if (ctg[m] == c && cs[m] < he) lo = m + 1; else hi = m; }
		reach[h] = make_int2(first / tile, (lo - 1) / tile);
	}
}
__device__ __forceinline__ void mark_dirty(const DirtyCols &D, int32_t h)
{
	const int2 r = D.reach[h];
	for (int32_t t = r.x; t <= r.y; ++t) atomicOr(&D.tiles[t >> 5], 1u << (t & 31));
}

// ---- branch round -----------------------------------------------------------------------------------------------------------
// The locality table of a branch round: one entry per (vertex, pair of its out-arcs): is there a LOCAL genome in which the two target genes
// are local to each other (branch.c:31-46)?  Every use of pg_n_local() in branch.c:60-98 only asks "== 0 ?" / "> 0 ?", so an entry is settled
// by the first genome that says yes.  A warp takes 32 consecutive entries: every lane decodes ONE entry (owner vertex by binary search in the
// offsets, position in the upper triangle, the two gene ids -- a chain of ~25 dependent loads, paid once per 32 entries instead of once per
// entry), then the warp goes through its entries with the lanes spread over the genomes.  ord_sensitive: an ordinal of a cs-tied survivor
// (GenePos.pad, k_cs_tie_mark) sits within 8 of the local_count threshold in a comparison that the distance clause does not settle, and no
// other genome settles the entry: the reference's numbering of those hits may differ (DESIGN.md "tie order").
__global__ void k_branch_local(const unsigned long long *__restrict__ idx, const DevCounts *__restrict__ dc, const ArcDev *__restrict__ arc, const SegDev *__restrict__ seg, const GenePos *__restrict__ pos, int32_t G,
	const uint32_t *__restrict__ nl_off, int32_t NV, uint32_t nl_cap, const unsigned long long *__restrict__ nl_total, BranchOpt o, uint8_t *nl, uint32_t *ord_sensitive)
{
	const int32_t n_vtx = 2 * (int32_t)dc->n_seg < NV? 2 * (int32_t)dc->n_seg : NV, lane = threadIdx.x & 31, n_warps = (gridDim.x * blockDim.x) >> 5;
	const int32_t n_nl = *nl_total <= (unsigned long long)nl_cap? (int32_t)dc->n_nl : 0;   // too many pairs for the table: k_branch_decide_fast raises the fallback
	for (int32_t base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < n_nl; base += n_warps * 32) {
		const int32_t ent = base + lane;
		int32_t g1 = 0, g2 = 0;
		if (ent < n_nl) {
			int32_t lo = 0, hi = n_vtx;                         // vertex owning this entry: last v with nl_off[v] <= ent
			while (lo < hi) { const int32_t m = (lo + hi) >> 1; if (nl_off[m] <= (uint32_t)ent) lo = m + 1; else hi = m; }
			const int32_t v = lo - 1;
			const unsigned long long e = idx[v];
			const uint32_t n = (uint32_t)e;
			uint32_t r = (uint32_t)ent - nl_off[v], i = 0;
			while (r >= n - i) { r -= n - i; ++i; }             // row i of the upper triangle holds n - i entries (k = i .. n-1)
			const ArcDev *a = arc + (e >> 32);
			g1 = seg[(uint32_t)a[i].x >> 1].gid; g2 = seg[(uint32_t)a[i + r].x >> 1].gid;
		}
		const int32_t n_here = n_nl - base < 32? n_nl - base : 32;
		bool my_any = false, my_shaky = false;
		for (int32_t q = 0; q < n_here; ++q) {
			const GenePos *p1 = pos + (size_t)__shfl_sync(0xffffffffu, g1, q) * G, *p2 = pos + (size_t)__shfl_sync(0xffffffffu, g2, q) * G;
			bool any = false, sens = false, settled = false;
			for (int32_t j0 = 0; j0 < G && !settled; j0 += 32) {
				const int32_t j = j0 + lane;
				bool robust = false;
				if (j < G) {
					const GenePos x = p1[j], y = p2[j];
					if (x.cid >= 0 && y.cid >= 0 && (o.frag_mode || x.cid == y.cid)) {
						const long long dd = (long long)x.cm - y.cm; const int32_t dcnt = x.ord - y.ord;
						const bool by_dist = dd >= -o.local_dist && dd <= o.local_dist, by_cnt = dcnt >= -o.local_count && dcnt <= o.local_count;
						bool shaky = false;
						if ((x.pad | y.pad) && !by_dist) {
							const long long ac = dcnt < 0? -(long long)dcnt : dcnt;
							shaky = ac > (long long)o.local_count - 8 && ac <= (long long)o.local_count + 8;
						}
						any = any || by_dist || by_cnt;
						sens = sens || shaky;
						robust = (by_dist || by_cnt) && !shaky;
					}
				}
				settled = __any_sync(0xffffffffu, robust);
			}
			any = __any_sync(0xffffffffu, any); sens = __any_sync(0xffffffffu, sens);
			if (lane == q) { my_any = any; my_shaky = sens && !settled; }
		}
		if (ent < n_nl) { nl[ent] = my_any? 1 : 0; if (my_shaky) *ord_sensitive = 1u; }
	}
}

__device__ __forceinline__ uint32_t tri_index_u(uint32_t i, uint32_t k, uint32_t n) { return i * n - i * (i - 1) / 2 + (k - i); }   // i <= k

// one thread per oriented vertex: branch.c:56-98 with the (summed over ranks) locality table.  tie_round: {tied survivors this round,
// long run, ord_sensitive} (Session::C_TIE_ROUND..): the round needs the reference's numbering of cs-tied survivors -> careful path.
__global__ void k_branch_decide_fast(const unsigned long long *__restrict__ idx, DevCounts *dc, ArcDev *arc, SegDev *seg, const uint32_t *__restrict__ nl_off, const uint8_t *__restrict__ nl,
	uint32_t nl_cap, const unsigned long long *__restrict__ nl_total, BranchOpt o, int32_t *grp_scratch, const uint32_t *__restrict__ tie_round)
{
	if (blockIdx.x == 0 && threadIdx.x == 0 && tie_round[0] && (tie_round[1] || tie_round[2])) dc->fallback = 1u;
	if (*nl_total > (unsigned long long)nl_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) dc->fallback = 1u; return; }   // the pair table does not fit: this run is redone
	PGX_GS(v, 2 * (int32_t)dc->n_seg) {
		const unsigned long long e = idx[v];
		const int32_t n = (int32_t)(uint32_t)e;
		if (n < 2) { seg[v >> 1].n_dist_loci[v & 1] = 0; continue; }                      // branch.c:56-57,66
		ArcDev *a = arc + (e >> 32);
		const uint8_t *t = nl + nl_off[v];
		int32_t best = 0;
		for (int32_t i = 0; i < n; ++i) best = best > a[i].s1? best : a[i].s1;
		for (int32_t i = 0; i < n; ++i) {
			const double r = 1.0 - (double)a[i].s1 / best;
			if (r > o.branch_diff) {
				int32_t sum = 0;
				for (int32_t k = 0; k < n; ++k)
					if (a[k].s1 == best) sum += t[k < i? tri_index_u(k, i, n) : tri_index_u(i, k, n)];
				const uint32_t wb = ((sum == 0 && r > o.branch_diff_dist) || r > o.branch_diff_cut)? 2u : 1u;   // branch.c:76-77
				a[i].bits = (a[i].bits & ~6u) | wb << 1;            // pg_arc_t: del:1, weak_br:2
			}
		}
		int32_t n_grp = 0;                                          // branch.c:82-90
		int32_t *grp = grp_scratch + (e >> 32);
		for (int32_t i = 0; i < n; ++i) grp[i] = 0;
		for (int32_t i = 0; i < n; ++i) {
			if (grp[i] == 0) grp[i] = ++n_grp;
			for (int32_t k = i + 1; k < n; ++k)
				if (grp[k] == 0 && t[tri_index_u(i, k, n)] > 0) grp[k] = grp[i];
		}
		seg[v >> 1].n_dist_loci[v & 1] = n_grp;
	}
}

// branch.c:108-145 + PG_SET_FILTER(weak_br == 2) (graph.c:309).  The two arcs an adjacency sits on (pg_get_arc, pgpriv.h:99-107) are found
// through the table of canonical keys and pos_of[]; a hit whose flt / weak_br bits change marks the sweep tiles it can influence.
// rec (optional): the pair record of the adjacency, written by k_arc_insert against THIS table (single rank: no look-up needed)
__device__ __forceinline__ uint32_t arc_weak_of(const ArcTable &T, const uint32_t *__restrict__ pos_of, const ArcDev *__restrict__ arc, unsigned long long x, const uint32_t *__restrict__ rec, int32_t t)
{
	const unsigned long long xm = arc_mirror(x);
	uint32_t s;
	if (rec) { const uint32_t r = rec[t]; s = r == ARC_NONE? ARC_NONE : r & REC_SLOT; } else s = arc_find(T, x < xm? x : xm);
	if (s == ARC_NONE) return 0u;
	return arc[pos_of[s * 2 + (xm < x? 1u : 0u)]].bits >> 1 & 3u;
}
__global__ void k_branch_hits_fast(const uint32_t *__restrict__ sv, const SvInfo *__restrict__ info, const DevCounts *__restrict__ dc, uint32_t *bits,
	const ArcDev *__restrict__ arc, ArcTable T, const uint32_t *__restrict__ pos_of, const uint32_t *__restrict__ rec, DirtyCols D, uint32_t *n_changed)
{
	const int32_t m = (int32_t)dc->m_sv;
	PGX_GS(t, m) {
		const uint32_t h = sv[t], b_in = bits[h];
		const SvInfo me = info[t];
		uint32_t wb = (b_in & PG_HB_WBR_MASK) >> PG_HB_WBR_SHIFT;
		if (t > 0) {
			const SvInfo p = info[t - 1];
			if (p.ctg == me.ctg) {                              // I am the later hit of (prev, me): arc (w^1 -> v^1), branch.c:130-132
				const uint32_t x = arc_weak_of(T, pos_of, arc, (unsigned long long)(me.w ^ 1u) << 32 | (p.w ^ 1u), rec, t);
				wb = wb > x? wb : x;
			}
		}
		if (t + 1 < m) {
			const SvInfo nx = info[t + 1];
			if (nx.ctg == me.ctg) {                             // I am the earlier hit of (me, next): arc (v -> w), branch.c:127-129
				const uint32_t x = arc_weak_of(T, pos_of, arc, (unsigned long long)me.w << 32 | nx.w, rec, t + 1);
				wb = wb > x? wb : x;
			}
		}
		uint32_t b = (b_in & ~PG_HB_WBR_MASK) | wb << PG_HB_WBR_SHIFT;
		if (wb == 2) b |= PG_HB_FLT;                               // graph.c:309
		if (b != b_in) { bits[h] = b; atomicAdd(n_changed, 1u); if (D.tiles) mark_dirty(D, (int32_t)h); }
	}
}
// graph.c:61-69 + PG_SET_FILTER(vtx == 0) (graph.c:312).  n_seg_new (optional): the segment count after the compaction that precedes this
// launch (k_seg_compact_fast still read the old count; nothing in this kernel reads either) -- stored by one thread instead of a launch of its own
__global__ void k_flag_vtx_fast(const int32_t *__restrict__ gidp, uint32_t *bits, int32_t n, const int32_t *__restrict__ g2s, DirtyCols D, DevCounts *dc, const uint32_t *__restrict__ n_seg_new)
{
	const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0 && n_seg_new) dc->n_seg = *n_seg_new;
	if (i >= n) return;
	const uint32_t b_in = bits[i];
	uint32_t b = b_in;
	if (g2s[gidp[i] >> 1] >= 0) b |= PG_HB_VTX; else { b &= ~PG_HB_VTX; b |= PG_HB_FLT; }
	if (b != b_in) { bits[i] = b; if (D.tiles && ((b ^ b_in) & PG_HB_FLT)) mark_dirty(D, i); }
}

// graph.c:237-263 in two launches: delete marks (occupancy, out-degree from idx[], distant loci) scanned in one pass, then the
// order-preserving compaction (pg_hard_delete) with the gene -> segment map (pg_gen_g2s; g2s is filled with -1 beforehand).
__global__ void __launch_bounds__(SCAN_B) k_seg_keep_scan(const SegDev *__restrict__ seg, const DevCounts *__restrict__ dc, const unsigned long long *__restrict__ idx,
	int32_t max_occ_total, int32_t max_degree, int32_t max_dist_loci, uint32_t *pos, uint32_t *n_keep, unsigned long long *status, uint32_t *ticket)
{
	__shared__ uint32_t wsum[33];
	__shared__ uint32_t s_tile, s_prefix;
	const int tid = threadIdx.x;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	const int32_t n_seg = (int32_t)dc->n_seg;
	const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)tid * SCAN_I;
	uint32_t fl = 0, sum = 0;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) {
		const int64_t s = base + k;
		if (s < n_seg) {
			const SegDev r = seg[s];
			bool del = r.tot_cnt > max_occ_total;                                            // graph.c:241
			del = del || (int32_t)(uint32_t)idx[2 * s] > max_degree || (int32_t)(uint32_t)idx[2 * s + 1] > max_degree;   // graph.c:244-248
			const int32_t m = r.n_dist_loci[0] > r.n_dist_loci[1]? r.n_dist_loci[0] : r.n_dist_loci[1];
			del = del || m > max_dist_loci;                                                  // graph.c:251-255
			if (!del) { fl |= 1u << k; ++sum; }
		}
	}
	uint32_t tile_sum;
	const uint32_t excl = block_exclusive_scan<SCAN_B>(sum, tile_sum, wsum);
	if (tid < 32) { const uint32_t p = lookback_prefix(status, tile, tile_sum); if (tid == 0) s_prefix = p; }
	__syncthreads();
	uint32_t run = s_prefix + excl;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) {
		const int64_t s = base + k;
		if (s < n_seg) pos[s] = (fl >> k & 1u)? run : ARC_NONE;
		if (s == n_seg) *n_keep = run;
		run += fl >> k & 1u;
	}
}
__global__ void k_seg_compact_fast(const SegDev *__restrict__ in, SegDev *out, DevCounts *dc, const uint32_t *__restrict__ pos, const uint32_t *__restrict__ n_keep, int32_t *g2s)
{
	const int32_t n_seg = (int32_t)dc->n_seg;                  // (the count is replaced by the LAST block to finish reading it: a separate launch does it)
	PGX_GS(s, n_seg) { const uint32_t p = pos[s]; if (p != ARC_NONE) { const SegDev r = in[s]; out[p] = r; g2s[r.gid] = (int32_t)p; } }
}
// [emu-end round]

} // namespace pgx
