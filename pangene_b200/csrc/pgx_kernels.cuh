// pgx_kernels.cuh -- the hot-path kernels (sm_100a).  Integer / indexing work: no tensor cores.
// Each kernel cites the reference lines whose result it must reproduce bit for bit; the order-free
// formulation they implement is spelled out (and pinned against the compiled reference) in oracle/pg_oracle.c.
#pragma once
#include "pgx_util.cuh"
#include "pgx_graphtypes.cuh"
#include "pangene_abi.h"

namespace pgx {

// internal marker kept in the unused high bits of the pg_hit_t bit-field word; stripped on materialisation
constexpr uint32_t HB_FIRST = 0x80000000u;       // first hit of its genome in (contig, cs) order (overlap.c:108 quirk)
constexpr uint32_t HB_ABI_MASK = 0x7ffu;

constexpr int AOS_WORDS = 22;                    // sizeof(pg_hit_t) / 4
// word offsets inside pg_hit_t (include/pangene_abi.h)
constexpr int W_PID = 0, W_CID = 3, W_RANK = 7, W_SORI = 8, W_SADJ = 9, W_SDOM = 10, W_NEXON = 11, W_OFFEXON = 12, W_PDOM = 13, W_PDOM0 = 14, W_BITS = 15, W_CS = 16, W_CM = 18, W_CE = 20;

__device__ __forceinline__ uint32_t mix32(uint32_t k)   // pgpriv.h:88-97
{
	k += ~(k << 15); k ^= k >> 10; k += k << 3; k ^= k >> 6; k += ~(k << 11); k ^= k >> 16;
	return k;
}
// overlap.c:82,137: score_adj<<33 | preferred<<32 | hash(pid)
__device__ __forceinline__ uint64_t score_key(int32_t sadj, int32_t gidp, int32_t pid)
{
	return (uint64_t)(int64_t)sadj << 33 | (uint64_t)(gidp & 1) << 32 | mix32((uint32_t)pid);
}

__device__ __forceinline__ int upper_bound_i32(const int32_t *a, int n, int32_t x)   // first k with a[k] > x
{
	int lo = 0, hi = n;
	while (lo < hi) { int m = (lo + hi) >> 1; if (a[m] <= x) lo = m + 1; else hi = m; }
	return lo;
}

// ------------------------------------------------------------------------------------------------
// stage A: raw (file) order -- unpack, per-(genome,protein) pseudogene flags (hit.c:66-105)
// ------------------------------------------------------------------------------------------------
// what the device needs of a pg_hit_t, packed by the host while it fills the pinned upload chunks (pgx_hostpipe.hpp)
struct RawHit { int32_t pid, cid, rank, sori, sadj, nexon, offexon, rev, cs, cm, ce, lof; };   // 48 bytes (lof: never read by the path itself, only handed on to the walk tokens)
// what changed, packed by the device for the way back
struct OutHit { uint32_t src; int32_t rank, sdom, pdom, pdom0; uint32_t bits; };               // 24 bytes

// hoff[G+1]: hit offsets of the local genomes; coff[G+1]: contig-key offsets; one thread per raw hit.
__global__ void k_unpack(const RawHit *__restrict__ raw, int32_t n, const int32_t *__restrict__ hoff, const int32_t *__restrict__ coff, int G,
	int32_t *r_rank, int32_t *r_gl, uint64_t *key, int32_t n_prot, uint32_t *err)
{
	int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	RawHit h = raw[t];
	int gl = upper_bound_i32(hoff, G + 1, t) - 1;
	uint32_t bad = 0;
	if (h.cid < 0 || h.cid >= coff[gl + 1] - coff[gl]) bad |= 2;
	if (h.pid < 0 || h.pid >= n_prot) bad |= 4;
	if (h.nexon < 1) bad |= 8;
	if (bad) { atomicOr(err, bad); h.cid = 0; h.cs = 0; }
	r_rank[t] = h.rank; r_gl[t] = gl;
	key[t] = (uint64_t)(uint32_t)(coff[gl] + h.cid) << 32 | (uint32_t)h.cs;
}

__global__ void k_fp_minmax(const RawHit *__restrict__ raw, const int32_t *__restrict__ r_gl, int32_t n, int32_t n_prot, int32_t *t_max, int32_t *t_min)
{
	int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	size_t k = (size_t)r_gl[t] * n_prot + raw[t].pid;
	int32_t ne = raw[t].nexon;
	atomicMax(&t_max[k], ne);
	atomicMin(&t_min[k], ne);
}

__device__ __forceinline__ bool fp_group_flagged(int32_t M, int32_t m) { return M > 1 && (m == 1 || m * 2 <= M); }   // hit.c:84

__global__ void k_fp_flag(const RawHit *__restrict__ raw, const int32_t *__restrict__ r_rank, const int32_t *__restrict__ r_gl,
	int32_t n, int32_t n_prot, const int32_t *__restrict__ t_max, const int32_t *__restrict__ t_min, int32_t *t_r1, uint8_t *r_pseudo)
{
	int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	size_t k = (size_t)r_gl[t] * n_prot + raw[t].pid;
	int32_t M = t_max[k], m = t_min[k], ne = raw[t].nexon;
	uint8_t ps = 0;
	if (fp_group_flagged(M, m)) {
		if (ne == 1 || ne * 2 <= M) ps = 1;                 // hit.c:88-89
		else atomicMin(&t_r1[k], r_rank[t]);                // lowest-rank survivor = "j1" of hit.c:90-91
	}
	r_pseudo[t] = ps;
}

__global__ void k_fp_rank(const RawHit *__restrict__ raw, const int32_t *__restrict__ r_gl, int32_t n, int32_t n_prot,
	const int32_t *__restrict__ t_max, const int32_t *__restrict__ t_min, const int32_t *__restrict__ t_r1, int32_t *r_rank)
{
	int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	size_t k = (size_t)r_gl[t] * n_prot + raw[t].pid;
	if (!fp_group_flagged(t_max[k], t_min[k])) return;
	int32_t r1 = t_r1[k], r = r_rank[t];
	if (r1 <= 0 || r1 == INT32_MAX) return;                 // hit.c:94
	if (r < r1) r_rank[t] = r + 1; else if (r == r1) r_rank[t] = 0;   // hit.c:95-97
}

// ------------------------------------------------------------------------------------------------
// stage B: gather into (contig, cs) order -- the one and only physical reordering of hits
// ------------------------------------------------------------------------------------------------
struct FullCols {            // all hits, (genome, contig, cs) order, packed int32 columns
	int32_t *ctg, *cs, *ce, *cm, *pid, *gidp, *sadj, *rank, *sori, *sdom, *pdom0, *pdom, *cds, *pmax;
	uint32_t *bits, *nex, *eoff, *esrc;
	uint64_t *key;
};

__global__ void k_gather(const RawHit *__restrict__ raw, const uint32_t *__restrict__ src, int32_t n, const int32_t *__restrict__ hoff, const int32_t *__restrict__ coff,
	const uint32_t *__restrict__ eoff_g, const int32_t *__restrict__ r_rank, const uint8_t *__restrict__ r_pseudo, const int32_t *__restrict__ r_gl,
	const int32_t *__restrict__ prot_gid, const uint8_t *__restrict__ gene_pref, FullCols c)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t s = src[i];
	RawHit h = raw[s];
	int gl = r_gl[s];
	int32_t gid = prot_gid[h.pid];
	uint32_t b = (uint32_t)h.rev & PG_HB_REV;
	if (r_pseudo[s]) b |= PG_HB_PSEUDO | PG_HB_FLT;          // read.c:246
	if (i == hoff[gl]) b |= HB_FIRST;
	c.ctg[i] = coff[gl] + h.cid;
	c.cs[i] = h.cs; c.cm[i] = h.cm; c.ce[i] = h.ce;
	c.pid[i] = h.pid; c.gidp[i] = gid << 1 | gene_pref[gid];
	c.sadj[i] = h.sadj; c.sori[i] = h.sori; c.rank[i] = r_rank[s];
	c.key[i] = score_key(h.sadj, gid << 1 | gene_pref[gid], h.pid);
	c.sdom[i] = 0; c.pdom0[i] = -1; c.pdom[i] = -1;          // parser state: read.c:133-134
	c.bits[i] = b;
	c.nex[i] = (uint32_t)h.nexon;
	c.esrc[i] = eoff_g[gl] + (uint32_t)h.offexon;
}

// copy each hit's exon list to its new home (contiguous in hit order) and sum the CDS length (overlap.c:45-51)
__global__ void k_gather_exons(const int2 *__restrict__ ex_raw, const uint32_t *__restrict__ esrc, const uint32_t *__restrict__ eoff, int32_t n, int2 *ex, int32_t *cds)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t b = eoff[i], e = eoff[i + 1], s = esrc[i];
	int32_t len = 0;
	for (uint32_t k = 0; k < e - b; ++k) { int2 x = ex_raw[s + k]; ex[b + k] = x; len += x.y - x.x; }
	cds[i] = len;
}

// ---- overlap.c:108 quirk, the moving part -------------------------------------------------------------------------------------
// pg_shadow starts its loop at i = 1: the hit at the FRONT of the reference's array never has its shadow bit reset (HB_FIRST).  That
// hit has the smallest cs of the genome's first contig; when several hits share that (contig, cs), which of them is in front changes
// from sort to sort of the reference's unstable radix-sort chain.  k_first_group finds the genomes where that can
// happen; Session::prepare_first_replay replays the chain for their first contig on the host (exact permutation, pgx_host.hpp) and
// k_set_first moves HB_FIRST to the right hit before every later sweep.  Genomes without such a tie (the usual case) cost nothing.
struct FirstGroup { int32_t i0, m, len; };   // first hit of the genome (F order); hits sharing its (contig, cs) -- 0 if alone; hits in that contig

__global__ void k_first_group(const int32_t *__restrict__ hoff, int32_t G, const int32_t *__restrict__ ctg, const int32_t *__restrict__ cs, FirstGroup *out, uint32_t *n_flagged)
{
	int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= G) return;
	const int32_t i0 = hoff[g], e = hoff[g + 1];
	FirstGroup r; r.i0 = i0; r.m = 0; r.len = 0;
	if (e - i0 >= 2) {
		const int32_t c = ctg[i0], s = cs[i0];
		int32_t m = 1;
		while (i0 + m < e && ctg[i0 + m] == c && cs[i0 + m] == s) ++m;
		if (m >= 2) {
			int32_t lo = i0 + m, hi = e;   // end of the contig: ctg is non-decreasing inside a genome
			while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (ctg[mid] == c) lo = mid + 1; else hi = mid; }
			r.m = m; r.len = lo - i0;
			atomicAdd(n_flagged, 1u);
		}
	}
	out[g] = r;
}

// one thread per flagged genome: the members of the first group that are still active sit at the head of the genome's active range
__global__ void k_set_first(const int32_t *__restrict__ flagged_g, const int32_t *__restrict__ target, int32_t n_flagged, const FirstGroup *__restrict__ fg,
	const int32_t *__restrict__ goff_act, const int32_t *__restrict__ full, uint32_t *bits)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_flagged) return;
	const int32_t g = flagged_g[k], t = target[k];
	const FirstGroup r = fg[g];
	for (int32_t a = goff_act[g]; a < goff_act[g + 1] && full[a] < r.i0 + r.m; ++a)
		bits[a] = full[a] == t? (bits[a] | HB_FIRST) : (bits[a] & ~HB_FIRST);
}

// ------------------------------------------------------------------------------------------------
// the interval sweep (overlap.c:58-178) as a set of pairs, evaluated from both members of every pair
// ------------------------------------------------------------------------------------------------
// [emu-begin sweep]  (tests/emu/build.py compiles the text up to [emu-end sweep] for the host: tests/test_sweep_emu.py)
struct SweepCols {
	int32_t n;
	const int32_t *ctg, *cs, *ce, *pid, *gidp, *rank, *cds, *sori, *pmax;
	const uint64_t *key;         // score_key() of every hit, immutable
	const uint32_t *bits, *eoff;
	const int2 *ex;
};
struct SweepOpt { double min_ov_ratio; int32_t check_strand; };

enum { SW_LOAD = 0, SW_SHADOW = 1, SW_ISO = 2 };

// ---- accessors: the pair logic below is written once and instantiated over "columns in global memory" and
// ---- "columns staged in shared memory by TMA bulk copies, global memory outside the staged window"
struct AccGlobal {
	SweepCols c;
	__device__ __forceinline__ int32_t ctg(int32_t p) const { return c.ctg[p]; }
	__device__ __forceinline__ int32_t cs(int32_t p) const { return c.cs[p]; }
	__device__ __forceinline__ int32_t ce(int32_t p) const { return c.ce[p]; }
	__device__ __forceinline__ int32_t pmax(int32_t p) const { return c.pmax[p]; }
	__device__ __forceinline__ uint32_t bits(int32_t p) const { return c.bits[p]; }
	__device__ __forceinline__ int32_t gidp(int32_t p) const { return c.gidp[p]; }
	__device__ __forceinline__ int32_t rank(int32_t p) const { return c.rank[p]; }
	__device__ __forceinline__ int32_t cds(int32_t p) const { return c.cds[p]; }
	__device__ __forceinline__ uint64_t key(int32_t p) const { return c.key[p]; }
	__device__ __forceinline__ uint32_t eoff(int32_t p) const { return c.eoff[p]; }
	__device__ __forceinline__ int2 exon(uint32_t k) const { return c.ex[k]; }
};


struct AccTile {
	SweepCols c;
	int32_t lo, hi;              // staged hits [lo, hi)
	uint32_t elo, ehi;           // staged exons [elo, ehi)
	const int32_t *s_ctg, *s_cs, *s_ce, *s_pmax, *s_gidp, *s_rank, *s_cds;
	const uint32_t *s_bits, *s_eoff;
	const uint64_t *s_key;
	const int2 *s_ex;
	__device__ __forceinline__ bool in(int32_t p) const { return p >= lo && p < hi; }
	__device__ __forceinline__ int32_t ctg(int32_t p) const { return in(p)? s_ctg[p - lo] : c.ctg[p]; }
	__device__ __forceinline__ int32_t cs(int32_t p) const { return in(p)? s_cs[p - lo] : c.cs[p]; }
	__device__ __forceinline__ int32_t ce(int32_t p) const { return in(p)? s_ce[p - lo] : c.ce[p]; }
	__device__ __forceinline__ int32_t pmax(int32_t p) const { return in(p)? s_pmax[p - lo] : c.pmax[p]; }
	__device__ __forceinline__ uint32_t bits(int32_t p) const { return in(p)? s_bits[p - lo] : c.bits[p]; }
	__device__ __forceinline__ int32_t gidp(int32_t p) const { return in(p)? s_gidp[p - lo] : c.gidp[p]; }
	__device__ __forceinline__ int32_t rank(int32_t p) const { return (s_rank && in(p))? s_rank[p - lo] : c.rank[p]; }
	__device__ __forceinline__ int32_t cds(int32_t p) const { return in(p)? s_cds[p - lo] : c.cds[p]; }
	__device__ __forceinline__ uint64_t key(int32_t p) const { return in(p)? s_key[p - lo] : c.key[p]; }
	__device__ __forceinline__ uint32_t eoff(int32_t p) const { return (p >= lo && p <= hi)? s_eoff[p - lo] : c.eoff[p]; }
	__device__ __forceinline__ int2 exon(uint32_t k) const { return (k >= elo && k < ehi)? s_ex[k - elo] : c.ex[k]; }
};

// |CDS(lo) & CDS(hi)| (overlap.c:6-42); lo precedes hi in (contig, cs) order
template<class Acc> __device__ __forceinline__ int32_t cds_overlap(const Acc &A, uint32_t b0, uint32_t e0, int32_t cs0, int32_t ce0, uint32_t b1, uint32_t e1, int32_t cs1, int32_t ce1)
{
	if (!(cs0 < ce1 && ce0 > cs1)) return 0;
	int32_t inter = 0;
	int2 x0 = A.exon(b0), x1 = A.exon(b1);
	while (true) {
		int32_t s0 = cs0 + x0.x, t0 = cs0 + x0.y, s1 = cs1 + x1.x, t1 = cs1 + x1.y;
		if (s0 < s1) {
			if (t0 < t1) { int32_t o = t0 - s1; inter += o > 0? o : 0; if (++b0 == e0) break; x0 = A.exon(b0); }
			else { inter += x1.y - x1.x; if (++b1 == e1) break; x1 = A.exon(b1); }
		} else {
			if (t1 < t0) { int32_t o = t1 - s0; inter += o > 0? o : 0; if (++b1 == e1) break; x1 = A.exon(b1); }
			else { inter += x0.y - x0.x; if (++b0 == e0) break; x0 = A.exon(b0); }
		}
	}
	return inter;
}

// Same walk, but only answers "are the two hits in relation?", i.e. "does the overlap reach `need`?".  The running
// overlap only grows, so the walk stops as soon as the answer is yes.
template<class Acc> __device__ __forceinline__ bool cds_related(const Acc &A, uint32_t b0, uint32_t e0, int32_t cs0, int32_t ce0, uint32_t b1, uint32_t e1, int32_t cs1, int32_t ce1, int32_t need)
{
	if (!(cs0 < ce1 && ce0 > cs1)) return false;
	int32_t inter = 0;
	int2 x0 = A.exon(b0), x1 = A.exon(b1);
	while (true) {
		int32_t s0 = cs0 + x0.x, t0 = cs0 + x0.y, s1 = cs1 + x1.x, t1 = cs1 + x1.y;
		bool adv0;
		if (s0 < s1) {
			if (t0 < t1) { int32_t o = t0 - s1; inter += o > 0? o : 0; adv0 = true; }
			else { inter += x1.y - x1.x; adv0 = false; }
		} else {
			if (t1 < t0) { int32_t o = t1 - s0; inter += o > 0? o : 0; adv0 = false; }
			else { inter += x0.y - x0.x; adv0 = true; }
		}
		if (inter >= need) return true;
		if (adv0) { if (++b0 == e0) break; x0 = A.exon(b0); } else { if (++b1 == e1) break; x1 = A.exon(b1); }
	}
	return false;
}

// smallest overlap that puts two hits in relation: 1 for hits of one gene (overlap.c:81,132: any CDS overlap), else the
// smallest integer ov with !((double)ov / shorter < ratio) (overlap.c:134-136) -- found with the SAME double division the
// reference uses, so the predicate is reproduced exactly; two or three divisions per pair instead of one per walk step.
__device__ __forceinline__ int32_t related_threshold(bool same_gene, int32_t cds_a, int32_t cds_b, double ratio)
{
	if (same_gene) return 1;
	const int32_t shorter = cds_a < cds_b? cds_a : cds_b;
	if (shorter <= 0) return 1;
	const double x = ratio * (double)shorter;
	double t = ceil(x);
	// Away from the rounding boundary the candidate needs no division at all: with shorter < 2^22 and both t - x and x - (t - 1) above
	// 2^-20, the two quotients below differ from `ratio` by more than 2^-42, far beyond the 2^-53 relative error of a double division;
	// and when x is an integer AND the product was exact (fma residue 0), t / shorter IS ratio (not less), (t - 1) / shorter is less.
	if (shorter < (1 << 22) && t >= 1.0 && t < 2147483000.0) {
		const double frac = t - x;
		if ((frac > 9.5367431640625e-07 && frac < 1.0 - 9.5367431640625e-07) || (frac == 0.0 && fma(ratio, (double)shorter, -x) == 0.0)) return (int32_t)t;
	}
	int32_t need = t < 1.0? 1 : (t > 2147483000.0? 2147483000 : (int32_t)t);
	while (need > 1 && !((double)(uint64_t)(need - 1) / shorter < ratio)) --need;
	while (need < 2147483000 && (double)(uint64_t)need / shorter < ratio) ++need;
	return need;
}

// One hit against its window.  Every pair is visited from both members, but only the member that WOULD LOSE if the pair
// is in relation ever walks the exon lists, and it skips the walk when the outcome cannot change its result any more
// (already shadowed and the partner's key cannot beat the current dominator).  Exactness: shadow = OR over losing related
// pairs; dominator = max key over them, ties to the smallest partner index (overlap.c:150,153 strict '<' in ascending
// partner order) -- a skipped pair can alter neither.
//   SW_LOAD   = pg_shadow(cal_dom_sc=1) followed by read.c:249-253: writes pid_dom0, score_dom; shadow/pid_dom stay reset
//   SW_SHADOW = pg_shadow(cal_dom_sc=0): writes the shadow bit and pid_dom
//   SW_ISO    = pg_flt_ov_isoform: writes flt_iso_ov|flt
template<int MODE, class Acc> __device__ __forceinline__ void sweep_hit(const Acc &A, const SweepOpt &o, int32_t i, int32_t n, uint32_t mbits, uint32_t *bits_out, int32_t *pdom_out, int32_t *sdom_out, uint32_t *tie_cnt)
{
	const int32_t mctg = A.ctg(i), mcs = A.cs(i), mce = A.ce(i), mgidp = A.gidp(i), mcds = A.cds(i);
	const uint32_t meb = A.eoff(i), mee = A.eoff(i + 1);
	const uint64_t mkey = A.key(i);
	int32_t mrank = INT32_MIN;           // loaded on the first key tie only
	bool lose = false;
	uint64_t best_key = 0; int32_t best_p = INT32_MAX;
#pragma unroll 1
	for (int side = 0; side < 2; ++side) {
		// side 0: earlier hits still open at my start (overlap.c:67-71,114-120), walking back; side 1: later hits that start before my end
#pragma unroll 1
		for (int32_t p = side == 0? i - 1 : i + 1; p >= 0 && p < n; p += side == 0? -1 : 1) {
			if (A.ctg(p) != mctg) break;
			int32_t pcs, pce;
			if (side == 0) {
				if (A.pmax(p) <= mcs) break;               // nothing further back reaches my start
				pce = A.ce(p);
				if (pce <= mcs) continue;
				pcs = A.cs(p);
			} else {
				pcs = A.cs(p);
				if (pcs >= mce) break;
				pce = A.ce(p);
			}
			const uint32_t pbits = A.bits(p);
			if (pbits & PG_HB_FLT) continue;                                              // overlap.c:70,122
			const int32_t pgidp = A.gidp(p);
			const bool same_gene = (pgidp >> 1) == (mgidp >> 1);
			if (MODE == SW_ISO && !same_gene) continue;                                   // overlap.c:77
			if (o.check_strand && ((pbits ^ mbits) & PG_HB_REV)) continue;                // overlap.c:78,128
			const uint64_t pkey = A.key(p);
			// would I be the loser?  the reference calls the later hit "i", the earlier "j" (overlap.c:139-147)
			bool me_loses;
			const uint32_t mw = mbits & PG_HB_WBR_MASK, pw = pbits & PG_HB_WBR_MASK;
			if (MODE == SW_ISO || same_gene || mw == pw) {
				if (mkey != pkey) me_loses = mkey < pkey;
				else {
					if (mrank == INT32_MIN) mrank = A.rank(i);
					int32_t prank = A.rank(p);
					// hi loses iff rank_hi > rank_lo; otherwise lo loses
					me_loses = side == 0? (mrank > prank) : !(prank > mrank);
				}
			} else me_loses = mw > pw;
			if (!me_loses) continue;
			const bool improves = pkey > best_key || (pkey == best_key && pkey > 0 && p < best_p);
			if (MODE == SW_LOAD && pkey == best_key && pkey > 0 && A.cs(p) == A.cs(best_p)) atomicAdd(tie_cnt, 1u);
			if (MODE != SW_ISO && lose && !improves) continue;
			const int32_t need = related_threshold(MODE == SW_ISO || same_gene, mcds, MODE == SW_ISO || same_gene? 0 : A.cds(p), o.min_ov_ratio);
			const uint32_t pb = A.eoff(p), pe = A.eoff(p + 1);
			const bool rel = side == 0? cds_related(A, pb, pe, pcs, pce, meb, mee, mcs, mce, need)
			                          : cds_related(A, meb, mee, mcs, mce, pb, pe, pcs, pce, need);
			if (!rel) continue;
			lose = true;
			if (MODE == SW_ISO) { side = 2; break; }                                      // one loss is all there is to know
			if (improves) best_key = pkey, best_p = p;
		}
	}
	if (MODE == SW_ISO) {
		bits_out[i] = lose? (mbits | PG_HB_FLT_OV | PG_HB_FLT) : mbits;                  // overlap.c:85-91
	} else if (MODE == SW_SHADOW) {
		uint32_t b = mbits;
		if (!(b & HB_FIRST)) b &= ~PG_HB_SHADOW;                                          // overlap.c:108,113: hit 0 is never reset
		if (lose) b |= PG_HB_SHADOW;
		bits_out[i] = b;
		pdom_out[i] = best_key > 0? A.c.pid[best_p] : -1;                                 // overlap.c:160-165
	} else {
		bits_out[i] = mbits;
		int32_t pd = -1, sd = -1;
		if (best_key > 0) {
			pd = A.c.pid[best_p];
			const int32_t pcs = A.cs(best_p), pce = A.ce(best_p);
			const uint32_t pb = A.eoff(best_p), pe = A.eoff(best_p + 1);
			const int32_t ov = best_p < i? cds_overlap(A, pb, pe, pcs, pce, meb, mee, mcs, mce) : cds_overlap(A, meb, mee, mcs, mce, pb, pe, pcs, pce);
			// overlap.c:170, IEEE double, left to right; -fmad=false keeps the products and sums un-contracted
			sd = (int32_t)(A.c.sori[i] * (1.0 - (double)ov / mcds) + A.c.sori[best_p] * ((double)ov / A.cds(best_p)) + .499);
		}
		pdom_out[i] = pd; sdom_out[i] = sd;
	}
}

// plain variant: one thread per hit, everything through L1/L2.  bits_out may not alias c.bits.
template<int MODE> __global__ void __launch_bounds__(128) k_sweep(SweepCols c, SweepOpt o, uint32_t *bits_out, int32_t *pdom_out, int32_t *sdom_out, uint32_t *tie_cnt)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= c.n) return;
	uint32_t mbits = c.bits[i];
	if (mbits & PG_HB_FLT) { bits_out[i] = mbits; return; }                              // overlap.c:63,112
	AccGlobal A; A.c = c;
	sweep_hit<MODE>(A, o, i, c.n, mbits, bits_out, pdom_out, sdom_out, tie_cnt);
}

// ---- TMA plumbing (1-D bulk copies global -> shared, completion on an mbarrier) ------------------
#ifdef PGX_SIMT_EMU   // host build of the sweep kernels for tests/test_sweep_emu.py: the copy is synchronous, the barrier trivially complete
__device__ __forceinline__ void mbar_init(uint64_t*, uint32_t) { }
__device__ __forceinline__ void fence_mbar_init() { }
__device__ __forceinline__ void fence_proxy_async() { }
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) { }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t*, uint32_t) { return true; }
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
#endif

// ---- the production sweep: a CTA owns SWQ_TILE consecutive hits of the (contig, cs) order ---------------------------------
// Staging: the eight packed columns of [tile - halo, tile + halo) and the exon lists of those hits are pulled into shared memory
// by one thread issuing ten cp.async.bulk copies (TMA, UBLKCP in SASS) that complete on one mbarrier; five CTAs per SM overlap
// each other's copy and compute phases.
// What a hit needs from the sweep is (a) whether ANY related partner beats it and (b) the related partner with the LARGEST key
// (smallest index on ties).  ncu on the thread-per-hit kernel above: 5 of 32 lanes active, ~2900 thread-instructions per hit --
// every thread walks exon lists for several partners inside a branchy window loop.  So:
//   phase 0  (first sweep only) a sparse table over the staged keys: "strongest hit of an index range" in two look-ups;
//   phase 1  the would-be dominator of every live hit -- the partner with the largest key among those the hit would lose to --
//            from two binary searches for the window bounds (pmax and cs are monotone inside a contig) and two range-maximum
//            queries (1a), or, where the table does not apply, from a straight-line scan of the window (1b); no exon list touched;
//   phase 2  ONE exon walk per hit, against that partner.  If it is in relation, it is the dominator and the hit is shadowed: no
//            other partner can have a larger key.  The (hit, candidate) pairs of the tile are bucketed by walk length (a counting
//            sort in shared memory) and dealt out again longest first, one pair per thread, so that the lanes of a warp walk lists
//            of about the same length; in SW_LOAD the exact overlap of overlap.c:170 comes out of the same walk;
//   phase 3  what is left -- the strongest candidate was not in relation (a bounding-box overlap without enough CDS overlap), or
//            the window leaves the staged range -- is parked and taken one hit at a time by a warp: 32 partners at once.
// History of the design with the ncu numbers of every step: profiles/README.md.  tests/test_sweep_emu.py runs this kernel on the
// CPU (SIMT shim) against the thread-per-hit kernel.
constexpr int SWQ_TILE = 256, SWQ_HALO = 64, SWQ_STG = SWQ_TILE + 2 * SWQ_HALO + 8, SWQ_EXCAP = 2560;
constexpr size_t SWQ_SMEM = (size_t)SWQ_STG * 8 + (size_t)SWQ_STG * 4 * 7 + (size_t)(SWQ_STG + 4) * 4 + (size_t)SWQ_EXCAP * 8 + 64;

// One walk of two exon lists that live in ONE address space (both staged in shared memory, or both in global memory).
// Adds up exactly what cds_overlap()/cds_related() add (overlap.c:6-42): in every one of the reference's four cases the
// amount is min(t0,t1) - max(s0,s1), clipped at 0 where the reference clips (exon records have y >= x: read.c builds them
// from non-negative CIGAR lengths), and the list to advance is the reference's.
template<bool EARLY> __device__ __forceinline__ int32_t walk_exons(const int2 *ex, uint32_t i0, uint32_t e0, int32_t cs0, uint32_t i1, uint32_t e1, int32_t cs1, int32_t need)
{
	int2 x0 = ex[i0], x1 = ex[i1];
	// Exons of one list that end before the other list begins add nothing and only advance their own list (overlap.c:17-27: "x ends earlier",
	// o <= 0): step over them without the full merge step.  A long gene against a short hit inside it spends most of its walk here.
	{
		const int32_t f1 = cs1 + x1.x;
		while (cs0 + x0.y <= f1) { if (++i0 == e0) return 0; x0 = ex[i0]; }
		const int32_t f0 = cs0 + x0.x;
		while (cs1 + x1.y <= f0) { if (++i1 == e1) return 0; x1 = ex[i1]; }
	}
	int32_t s0 = cs0 + x0.x, t0 = cs0 + x0.y, s1 = cs1 + x1.x, t1 = cs1 + x1.y, inter = 0;
	while (true) {
		const bool adv0 = s0 < s1? t0 < t1 : !(t1 < t0);
		const int32_t o = (t0 < t1? t0 : t1) - (s0 > s1? s0 : s1);
		inter += o > 0? o : 0;
		if (EARLY && inter >= need) break;
		// advance ONE list, without a branch per list (the lanes of a warp take the two sides in any mix) and on 32-bit list positions
		i0 += adv0? 1u : 0u; i1 += adv0? 0u : 1u;
		const uint32_t in = adv0? i0 : i1;
		if (in == (adv0? e0 : e1)) break;
		const int2 x = ex[in];
		const int32_t base = adv0? cs0 : cs1, ns = base + x.x, nt = base + x.y;
		s0 = adv0? ns : s0; t0 = adv0? nt : t0; s1 = adv0? s1 : ns; t1 = adv0? t1 : nt;
	}
	return inter;
}

// the walk of one pair (the earlier hit first, overlap.c:80): staged exon lists when both are inside the staged range
template<bool EARLY> __device__ __forceinline__ int32_t pair_overlap(const int2 *s_ex, uint32_t elo, uint32_t ehi, const int2 *g_ex, uint32_t b0, uint32_t e0, int32_t cs0, int32_t ce0,
                                                                     uint32_t b1, uint32_t e1, int32_t cs1, int32_t ce1, int32_t need)
{
	if (!(cs0 < ce1 && ce0 > cs1)) return 0;                                         // overlap.c:13
	if (b0 >= elo && e0 <= ehi && b1 >= elo && e1 <= ehi) return walk_exons<EARLY>(s_ex, b0 - elo, e0 - elo, cs0, b1 - elo, e1 - elo, cs1, need);
	return walk_exons<EARLY>(g_ex, b0, e0, cs0, b1, e1, cs1, need);
}

// dirty (optional, SW_SHADOW with bits_out == c.bits): one bit per tile; a tile whose bit is clear holds no hit whose inputs changed since the
// previous sweep (pgx_round.cuh: mark_dirty) and keeps its shadow bits and dominators.
// VAR (profiles/README.md has the A/B numbers): 0 = the round-2 kernel as first measured (64-bit cluster maxima by compare-and-swap loops, pid / score
// columns read from global memory when a result is written, the isoform sweep on the cluster shortcut); 1 = cluster maxima on the upper key
// word with native 32-bit shared-memory atomics (a tie of that word at the top of a cluster sends its hits to the exact window scan), holder
// count and slot in ONE atomic, the pid / score_ori columns staged beside the others (no global load on the result path), warp prefixes from
// one 128-bit load, no flag-word store in SW_LOAD (the isoform sweep that follows writes every flag word), the isoform sweep on the window scan;
// 2 = as 1 with the isoform sweep on the cluster shortcut; 3 = as 1 without the cluster pass in SW_LOAD either (every hit scans its window).
// SW_SHADOW is the same code for every VAR.
template<int MODE, int VAR> __global__ void __launch_bounds__(SWQ_TILE, 5) k_sweep_tiled(SweepCols c, SweepOpt o, uint32_t *bits_out, int32_t *pdom_out, int32_t *sdom_out, uint32_t *tie_cnt, const uint32_t *dirty)
{
	constexpr bool NEWP0 = VAR != 0 && MODE != SW_SHADOW;                                // the new cluster pass + staged result columns
	if (dirty && !(dirty[blockIdx.x >> 5] >> (blockIdx.x & 31) & 1u)) return;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	unsigned char *sp = smem_raw;
	uint64_t *s_key = (uint64_t*)sp; sp += (size_t)SWQ_STG * 8;
	int32_t *s_ctg = (int32_t*)sp; sp += (size_t)SWQ_STG * 4;
	int32_t *s_cs = (int32_t*)sp; sp += (size_t)SWQ_STG * 4;
	int32_t *s_ce = (int32_t*)sp; sp += (size_t)SWQ_STG * 4;
	int32_t *s_pmax = (int32_t*)sp; sp += (size_t)SWQ_STG * 4;
	int32_t *s_gidp = (int32_t*)sp; sp += (size_t)SWQ_STG * 4;
	int32_t *s_cds = (int32_t*)sp; sp += (size_t)SWQ_STG * 4;
	uint32_t *s_bits = (uint32_t*)sp; sp += (size_t)SWQ_STG * 4;
	uint32_t *s_eoff = (uint32_t*)sp; sp += (size_t)(SWQ_STG + 4) * 4;
	int2 *s_ex = (int2*)sp; sp += (size_t)SWQ_EXCAP * 8;
	uint64_t *bar = (uint64_t*)sp; sp += 8;
	uint32_t *s_misc = (uint32_t*)sp;                                                    // [0],[1]: staged exon range
	__shared__ uint16_t s_cst[SWQ_STG + 8];      // s_cst[q]: first staged slot of the overlap cluster that slot q belongs to
	constexpr int P0N = SWQ_STG + 8;             // 400 entries; every quarter of s_p0 below starts on a 16-byte boundary
	__shared__ __align__(16) unsigned char s_p0[P0N * 16];
	// VAR 0: at a cluster's first slot the largest key in the cluster (64 bit), the leftmost slot that holds it, how many slots hold it
	unsigned long long *s_cmax = (unsigned long long*)s_p0;
	uint32_t *s_ctop = (uint32_t*)(s_p0 + P0N * 8), *s_ccnt = (uint32_t*)(s_p0 + P0N * 12);
	// VAR 1, 2: the largest UPPER key word (score_adj << 1 | preferred) of the cluster, holders << 16 | sum of their slots; staged score_ori and pid
	uint32_t *s_chi = (uint32_t*)s_p0, *s_cinfo = (uint32_t*)(s_p0 + P0N * 4);
	int32_t *s_sori = (int32_t*)(s_p0 + P0N * 8), *s_pid = (int32_t*)(s_p0 + P0N * 12);
	__shared__ __align__(16) int32_t s_wmax[SWQ_TILE / 32];
	__shared__ __align__(16) uint16_t s_wcnt[SWQ_TILE / 32];
	__shared__ uint16_t s_live[SWQ_TILE];
	__shared__ uint32_t s_hist[64], s_job[SWQ_TILE], s_njob, s_flags;

	const int32_t t0 = blockIdx.x * SWQ_TILE, tid = threadIdx.x;
	int32_t lo = t0 - SWQ_HALO; if (lo < 0) lo = 0;
	lo &= ~3;                                                                            // 16-byte aligned column offsets
	int32_t hi = t0 + SWQ_TILE + SWQ_HALO; if (hi > c.n) hi = c.n;
	const int32_t cnt4 = (hi - lo + 3) & ~3, cnte = (hi - lo + 1 + 3) & ~3;              // columns are allocated with slack past n
	if (tid < 64) s_hist[tid] = 0;
	if (tid == 0) {
		s_flags = 0;
		mbar_init(bar, 1);
		fence_mbar_init();
		uint32_t e0 = c.eoff[lo] & ~1u, e1 = c.eoff[hi];
		uint32_t ne = (e1 - e0 + 1) & ~1u;
		if (ne > SWQ_EXCAP) ne = SWQ_EXCAP;
		s_misc[0] = e0; s_misc[1] = e0 + ne;
		fence_proxy_async();
		mbar_expect_tx(bar, (uint32_t)cnt4 * 4u * (7u + (NEWP0 && MODE == SW_LOAD? 2u : 0u)) + (uint32_t)cnt4 * 8u + (uint32_t)cnte * 4u + ne * 8u);
		if (NEWP0 && MODE == SW_LOAD) { bulk_g2s(s_sori, c.sori + lo, cnt4 * 4, bar); bulk_g2s(s_pid, c.pid + lo, cnt4 * 4, bar); }
		bulk_g2s(s_ctg, c.ctg + lo, cnt4 * 4, bar); bulk_g2s(s_cs, c.cs + lo, cnt4 * 4, bar); bulk_g2s(s_ce, c.ce + lo, cnt4 * 4, bar);
		bulk_g2s(s_pmax, c.pmax + lo, cnt4 * 4, bar); bulk_g2s(s_gidp, c.gidp + lo, cnt4 * 4, bar); bulk_g2s(s_cds, c.cds + lo, cnt4 * 4, bar);
		bulk_g2s(s_bits, c.bits + lo, cnt4 * 4, bar); bulk_g2s(s_key, c.key + lo, cnt4 * 8, bar);
		bulk_g2s(s_eoff, c.eoff + lo, cnte * 4, bar);
		bulk_g2s(s_ex, c.ex + e0, ne * 8, bar);
	}
	__syncthreads();
	while (!mbar_try_wait(bar, 0)) { }

	// ---- phase 0 (first sweep only, not with -S): overlap clusters.  A cluster is a maximal run of staged hits in which every hit starts
	// before the running maximum of the ends before it (same contig): a hit's window (overlap.c:67-71,114-120) never leaves its cluster, so the
	// strongest hit of the cluster -- the leftmost maximum of the score key; filtered hits count as key 0 and get key 0 in the staged copy, nothing
	// below reads the key of a filtered hit for any other purpose -- is the strongest hit of every window that contains it.  One pass finds
	// the cluster of every slot (a block-wide running maximum over "I start a cluster"), one loop per cluster finds its strongest hit.
	// Only usable where the loser of a pair is decided by the keys alone, i.e. where no unfiltered hit of the tile carries a weak_br mark
	// (overlap.c:139-147): flag 1.
	const int32_t S = hi - lo;
	const bool rmq_mode = ((MODE == SW_LOAD && VAR != 3) || (MODE == SW_ISO && (VAR == 0 || VAR == 2))) && !o.check_strand;   // (the later sweeps run on the small active set and mostly see weak_br marks)
	// filtered hits (25 % in the first sweep) do no work: the live ones are compacted so that they fill whole warps.
	const int32_t i0 = t0 + tid;
	const bool r0 = i0 < c.n;
	const uint32_t bits0 = s_bits[r0? i0 - lo : 0];
	if (!(NEWP0 && MODE == SW_LOAD) && r0 && (bits0 & PG_HB_FLT)) bits_out[i0] = bits0;  // overlap.c:63,112
	const bool live0 = r0 && !(bits0 & PG_HB_FLT);
	const uint32_t bal = __ballot_sync(0xffffffffu, live0);
	if ((tid & 31) == 0) s_wcnt[tid >> 5] = (uint16_t)__popc(bal);
	int32_t cst0 = -1, cst1 = -1;                                                         // cluster starts of slots 2 tid and 2 tid + 1, before the carry from the threads to the left
	if (rmq_mode) {
		uint32_t flags = 0;
		const int32_t q0 = 2 * tid, q1 = q0 + 1;
		if (NEWP0) {   // every slot starts out as an empty cluster head: no barrier between finding the heads and the first atomics below
			if (q0 < S) { s_chi[q0] = 0u; s_cinfo[q0] = 0u; }
			if (q1 < S) { s_chi[q1] = 0u; s_cinfo[q1] = 0u; }
		}
		if (q0 < S) {
			const uint32_t bq = s_bits[q0];
			if (!(bq & PG_HB_FLT) && (bq & PG_HB_WBR_MASK)) flags |= 1u;
			if (q0 == 0 || s_ctg[q0] != s_ctg[q0 - 1] || s_pmax[q0 - 1] <= s_cs[q0]) cst0 = q0;
		}
		if (q1 < S) {
			const uint32_t bq = s_bits[q1];
			if (!(bq & PG_HB_FLT) && (bq & PG_HB_WBR_MASK)) flags |= 1u;
			cst1 = (s_ctg[q1] != s_ctg[q0] || s_pmax[q0] <= s_cs[q1])? q1 : cst0;
		} else cst1 = cst0;
		if (flags) atomicOr(&s_flags, flags);
		int32_t m = cst1;                                                                 // inclusive running maximum over the threads of the warp
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, m, d); if ((tid & 31) >= d && t > m) m = t; }
		if ((tid & 31) == 31) s_wmax[tid >> 5] = m;
		const int32_t left = __shfl_up_sync(0xffffffffu, m, 1);
		cst1 = m; if ((tid & 31) > 0 && left > cst0) cst0 = left;                         // (lane 0 gets the carry of the warps to its left below)
	}
	__syncthreads();
	int32_t n_live = 0;
	{
		int32_t base = 0;
		if (NEWP0) {   // the eight 16-bit warp counts in one 128-bit load; counts are <= 32, so the packed halves add without carries
			static_assert(SWQ_TILE / 32 == 8, "eight warps");
			const uint4 v = *(const uint4*)s_wcnt;
			const int W = tid >> 5;
			const uint32_t all = v.x + v.y + v.z + v.w;
			n_live = (int32_t)((all & 0xffffu) + (all >> 16));
			uint32_t pre = (W >= 2? v.x : 0u) + (W >= 4? v.y : 0u) + (W >= 6? v.z : 0u);   // whole words before mine
			const uint32_t mine = W < 2? v.x : W < 4? v.y : W < 6? v.z : v.w;
			pre += (W & 1)? (mine & 0xffffu) : 0u;                                         // the even warp of my word precedes the odd one
			base = (int32_t)((pre & 0xffffu) + (pre >> 16));
		} else {
#pragma unroll
			for (int w = 0; w < SWQ_TILE / 32; ++w) { if (w < (tid >> 5)) base += s_wcnt[w]; n_live += s_wcnt[w]; }
		}
		if (live0) s_live[base + __popc(bal & ((1u << (tid & 31)) - 1u))] = (uint16_t)tid;
	}
	if (rmq_mode) {
		int32_t carry = -1;
		if (NEWP0) {
			const int4 a = *(const int4*)s_wmax, b = *(const int4*)(s_wmax + 4);
			const int W = tid >> 5;
			if (W > 0 && a.x > carry) carry = a.x;
			if (W > 1 && a.y > carry) carry = a.y;
			if (W > 2 && a.z > carry) carry = a.z;
			if (W > 3 && a.w > carry) carry = a.w;
			if (W > 4 && b.x > carry) carry = b.x;
			if (W > 5 && b.y > carry) carry = b.y;
			if (W > 6 && b.z > carry) carry = b.z;
		} else {
#pragma unroll
			for (int w = 0; w < SWQ_TILE / 32; ++w) if (w < (tid >> 5) && s_wmax[w] > carry) carry = s_wmax[w];
		}
		if (carry > cst0) cst0 = carry;
		if (carry > cst1) cst1 = carry;
		const int32_t q0 = 2 * tid, q1 = q0 + 1;
		if (NEWP0) {
			if (q0 < S) { s_cst[q0] = (uint16_t)cst0; if (s_bits[q0] & PG_HB_FLT) s_key[q0] = 0ull; }
			if (q1 < S) { s_cst[q1] = (uint16_t)cst1; if (s_bits[q1] & PG_HB_FLT) s_key[q1] = 0ull; }
		} else {
			if (q0 < S) { s_cst[q0] = (uint16_t)cst0; if (s_bits[q0] & PG_HB_FLT) s_key[q0] = 0ull; if (cst0 == q0) { s_cmax[q0] = 0ull; s_ctop[q0] = 0xffffffffu; s_ccnt[q0] = 0u; } }
			if (q1 < S) { s_cst[q1] = (uint16_t)cst1; if (s_bits[q1] & PG_HB_FLT) s_key[q1] = 0ull; if (cst1 == q1) { s_cmax[q1] = 0ull; s_ctop[q1] = 0xffffffffu; s_ccnt[q1] = 0u; } }
		}
	}
	if (!(NEWP0 && rmq_mode)) __syncthreads();   // (block-uniform; the new cluster pass goes straight on: its tables were zeroed before the barrier above, and it reads only its own two slots until the next one)
	if (rmq_mode && NEWP0) {
		// strongest hit of every cluster, with native 32-bit atomics: the maximum of the UPPER key word (score_adj << 1 | preferred), then one
		// add per holder of that maximum: holders << 16 | sum of their slots.  One holder: the low half IS its slot, and it holds the largest
		// 64-bit key of the cluster (a larger key would have a larger or equal upper word).  Several holders (equal adjusted scores at the top
		// of a cluster: 0.3 % of the clusters of the bench workload) or none: the hits of the cluster take the window scan, which is exact.
		const int32_t q0 = 2 * tid, q1 = q0 + 1;
		const unsigned long long k0 = q0 < S? s_key[q0] : 0ull, k1 = q1 < S? s_key[q1] : 0ull;
		const uint32_t h0 = (uint32_t)(k0 >> 32), h1 = (uint32_t)(k1 >> 32);
		if (k0) atomicMax(&s_chi[cst0], h0);
		if (k1) atomicMax(&s_chi[cst1], h1);
		__syncthreads();
		if (k0 && h0 == s_chi[cst0]) atomicAdd(&s_cinfo[cst0], 0x10000u + (uint32_t)q0);
		if (k1 && h1 == s_chi[cst1]) atomicAdd(&s_cinfo[cst1], 0x10000u + (uint32_t)q1);
		__syncthreads();
	} else if (rmq_mode) {   // strongest hit of every cluster: maximum of the key, then the leftmost slot that holds it and how many do (shared-memory atomics)
		const int32_t q0 = 2 * tid, q1 = q0 + 1;
		const unsigned long long k0 = q0 < S? s_key[q0] : 0ull, k1 = q1 < S? s_key[q1] : 0ull;
		if (k0) atomicMax(&s_cmax[cst0], k0);
		if (k1) atomicMax(&s_cmax[cst1], k1);
		__syncthreads();
		if (k0 && k0 == s_cmax[cst0]) { atomicMin(&s_ctop[cst0], (uint32_t)q0); atomicAdd(&s_ccnt[cst0], 1u); }
		if (k1 && k1 == s_cmax[cst1]) { atomicMin(&s_ctop[cst1], (uint32_t)q1); atomicAdd(&s_ccnt[cst1], 1u); }
		__syncthreads();
	}
	const bool in_range = tid < n_live;
	const int32_t i = in_range? t0 + (int32_t)s_live[tid] : t0;
	const int32_t mq = i - lo;
	const uint32_t mbits = s_bits[mq];
	const uint32_t elo = s_misc[0], ehi = s_misc[1];
	const int32_t mctg = s_ctg[mq], mcs = s_cs[mq], mce = s_ce[mq], mgene = s_gidp[mq] >> 1;
	const uint32_t mw = mbits & PG_HB_WBR_MASK;
	const uint64_t mkey = s_key[mq];
	int32_t top_p = INT32_MAX; uint64_t top_key = 0;
	bool slow = false;

	// result of one hit (overlap.c:85-91,160-170); top = its dominator, ov = their exact CDS overlap (SW_LOAD)
	auto emit = [&](int32_t h, int32_t hq, uint32_t hbits, bool lose, bool dom, int32_t top, int32_t ov) {
		if (MODE == SW_ISO) bits_out[h] = lose? (hbits | PG_HB_FLT_OV | PG_HB_FLT) : hbits;
		else if (MODE == SW_SHADOW) {
			uint32_t b = hbits;
			if (!(b & HB_FIRST)) b &= ~PG_HB_SHADOW;                                      // overlap.c:108,113: hit 0 is never reset
			if (lose) b |= PG_HB_SHADOW;
			bits_out[h] = b;
			pdom_out[h] = (lose && dom)? c.pid[top] : -1;
		} else {
			if (!NEWP0) bits_out[h] = hbits;                                              // (unchanged; the isoform sweep that follows writes every flag word)
			int32_t pd = -1, sd = -1;
			if (lose && dom) {
				// overlap.c:170; `top` is a staged hit on this path, so with NEWP0 its pid and score come out of shared memory
				const int32_t so_h = NEWP0? s_sori[hq] : c.sori[h], so_t = NEWP0? s_sori[top - lo] : c.sori[top];
				pd = NEWP0? s_pid[top - lo] : c.pid[top];
				sd = (int32_t)(so_h * (1.0 - (double)ov / s_cds[hq]) + so_t * ((double)ov / s_cds[top - lo]) + .499);
			}
			pdom_out[h] = pd; sdom_out[h] = sd;
		}
	};

	if (in_range) {
	// ---- phase 1a: the would-be dominator from the cluster: if the strongest hit of my cluster is in my window, no partner can have a larger
	// key.  Whatever this cannot settle (the strongest hit does not reach me, its key occurs twice in the cluster -- same protein, same score:
	// ranks and the tie detector need the partner-by-partner view --, or I am the strongest hit of a cluster with such a twin) falls through to
	// the scan below; clusters that touch the ends of the staged range are finished by the parked path (phase 3).
	bool scan = true;
	if (rmq_mode && (MODE == SW_ISO || !(s_flags & 1u))) {   // (the isoform filter compares keys whatever the weak_br marks say, overlap.c:82-84)
		const int32_t head = s_cst[mq];
		const uint32_t info = NEWP0? s_cinfo[head] : 0u;
		const uint32_t n_top = NEWP0? info >> 16 : s_ccnt[head];
		const uint32_t tq = NEWP0? (n_top? info & 0xffffu : 0xffffffffu) : s_ctop[head];
		scan = false;
		if ((head == 0 && lo > 0) || (head == (int32_t)s_cst[S - 1] && hi < c.n)) slow = true;   // the cluster (and with it possibly my window) may continue outside the staged range
		else if (n_top > 1u) scan = true;                                                 // the strongest key (NEWP0: its upper word) occurs twice
		else if (tq == (uint32_t)mq || tq == 0xffffffffu) { }                             // nobody in my cluster beats me: top_p stays unset
		else if (MODE == SW_ISO && (s_gidp[tq] >> 1) != mgene) scan = true;               // overlap.c:77: only isoforms of my own gene count; the strongest of THOSE needs the scan
		else if ((int32_t)tq < mq? s_ce[tq] > mcs : s_cs[tq] < mce) { top_key = NEWP0? s_key[tq] : s_cmax[head]; top_p = lo + (int32_t)tq; }   // overlap.c:70,118: it is one of my partners
		else scan = true;
	}
	// ---- phase 1b: the would-be dominator by scanning the window.  One loop walks back, then forward: a lane whose backward
	// window is short starts its forward window while its neighbours are still going back.
	if (scan && !slow) {
		top_p = INT32_MAX; top_key = 0;
		int side = 0, step = -1;
#pragma unroll 1
		for (int32_t p = i - 1;; p += step) {
			bool end = false;
			int32_t q = 0, pcs = 0, pce = 0;
			if (p < lo) { slow = slow || p >= 0; end = true; }           // the window may continue outside the staged range
			else if (p >= hi) { slow = slow || p < c.n; end = true; }
			else {
				q = p - lo; pcs = s_cs[q]; pce = s_ce[q];
				end = s_ctg[q] != mctg || (side == 0? s_pmax[q] <= mcs : pcs >= mce);   // overlap.c:67-71,114-120: nothing further out can reach me
			}
			if (end) { if (side == 1) break; side = 1; step = 1; p = i; continue; }
			const uint32_t pbits = s_bits[q];
			const int32_t pgene = s_gidp[q] >> 1;
			const uint64_t pkey = s_key[q];
			const bool same_gene = pgene == mgene;
			bool ok = !(pbits & PG_HB_FLT) && pce > mcs;                 // overlap.c:70,118,122 (pce > mcs always holds on the forward side)
			if (MODE == SW_ISO) ok = ok && same_gene;                    // overlap.c:77
			if (o.check_strand) ok = ok && !((pbits ^ mbits) & PG_HB_REV);   // overlap.c:78,128
			const uint32_t pw = pbits & PG_HB_WBR_MASK;
			const bool key_rule = MODE == SW_ISO || same_gene || mw == pw;   // overlap.c:139-147
			bool loses = key_rule? mkey < pkey : mw > pw;
			if (ok && key_rule && mkey == pkey) {                        // same protein, same score: ranks decide (rare)
				const int32_t mrank = c.rank[i], prank = c.rank[p];
				loses = side == 0? (mrank > prank) : !(prank > mrank);
			}
			if (ok && loses) {
				if (MODE == SW_LOAD && pkey == top_key && top_p != INT32_MAX && pcs == s_cs[top_p - lo]) atomicAdd(tie_cnt, 1u);   // tie detector
				if (pkey > top_key || (pkey == top_key && p < top_p)) { top_key = pkey; top_p = p; }
			}
		}
	}
	}   // in_range

	// ---- phase 2: ONE walk per hit, against its would-be dominator; pairs bucketed by walk length and dealt out again
	const bool has_job = in_range && !slow && top_p != INT32_MAX;
	if (in_range && !slow && !has_job) emit(i, mq, mbits, false, false, 0, 0);       // nobody to lose to
	bool slow2 = false;
	int32_t i2 = t0;
	auto run_job = [&](int32_t jm, int32_t jt, bool dom) {
		i2 = lo + jm;
		const int32_t need = related_threshold(MODE == SW_ISO || (s_gidp[jt] >> 1) == (s_gidp[jm] >> 1), s_cds[jm], s_cds[jt], o.min_ov_ratio);
		const int32_t aq = jt < jm? jt : jm, bq = jt < jm? jm : jt;                  // the earlier hit goes first (overlap.c:80)
		const int32_t ov = pair_overlap<MODE != SW_LOAD>(s_ex, elo, ehi, c.ex, s_eoff[aq], s_eoff[aq + 1], s_cs[aq], s_ce[aq], s_eoff[bq], s_eoff[bq + 1], s_cs[bq], s_ce[bq], need);
		if (ov >= need) emit(i2, jm, s_bits[jm], true, dom, lo + jt, ov);
		else slow2 = true;                                                           // the strongest candidate did not hold: weaker ones must be tried
	};
	if (MODE == SW_ISO) {   // isoforms of one gene: any CDS overlap counts, the walk stops at the first common base -- nothing to balance
		if (has_job) run_job(mq, top_p - lo, top_key > 0);
	} else {
		uint32_t bin = 0, pos = 0;
		if (has_job) {
			const int32_t tq = top_p - lo;
			const uint32_t len = (s_eoff[mq + 1] - s_eoff[mq]) + (s_eoff[tq + 1] - s_eoff[tq]);
			bin = len >= 64u? 0u : 63u - len;                                        // longest first
			pos = atomicAdd(&s_hist[bin], 1u);
		}
		__syncthreads();
		if (tid < 32) {   // exclusive scan over the 64 buckets
			const uint32_t a = s_hist[2 * tid], b = s_hist[2 * tid + 1];
			uint32_t incl = a + b;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += v; }
			const uint32_t excl = incl - (a + b);
			s_hist[2 * tid] = excl; s_hist[2 * tid + 1] = excl + a;
			if (tid == 31) s_njob = incl;
		}
		__syncthreads();
		if (has_job) s_job[s_hist[bin] + pos] = (uint32_t)mq | (uint32_t)(top_p - lo) << 9 | (top_key > 0? 0x80000000u : 0u);
		__syncthreads();
		if (tid < (int32_t)s_njob) {
			const uint32_t job = s_job[tid];
			run_job((int32_t)(job & 511u), (int32_t)(job >> 9 & 511u), (job >> 31) != 0);
		}
	}
	// ---- phase 3: the hits that could not be settled above (a few per cent) are parked and then taken one at a time by a warp: the 32 lanes
	// try 32 partners at once -- all would-be-loser pairs walk their exon lists in parallel, a shuffle reduction picks the
	// related partner with the largest key (smallest index on ties).  Round 0: hits whose window leaves the staged range;
	// round 1: hits whose strongest candidate was not in relation.
	{
		const int lane = tid & 31;
		AccTile T;
		T.c = c; T.lo = lo; T.hi = hi; T.elo = elo; T.ehi = ehi;
		T.s_ctg = s_ctg; T.s_cs = s_cs; T.s_ce = s_ce; T.s_pmax = s_pmax; T.s_gidp = s_gidp; T.s_rank = nullptr; T.s_cds = s_cds; T.s_bits = s_bits; T.s_eoff = s_eoff; T.s_key = s_key; T.s_ex = s_ex;
#pragma unroll 1
		for (int rnd = 0; rnd < 2; ++rnd) {
		uint32_t parked = __ballot_sync(0xffffffffu, rnd == 0? slow : slow2);
		const int32_t isel = rnd == 0? i : i2;
#pragma unroll 1
		while (parked) {
			const int src_lane = __ffs(parked) - 1;
			parked &= parked - 1;
			const int32_t h = __shfl_sync(0xffffffffu, isel, src_lane), hq = h - lo;
			const int32_t hctg = s_ctg[hq], hcs = s_cs[hq], hce = s_ce[hq], hgene = s_gidp[hq] >> 1, hcds = s_cds[hq];
			const uint32_t hbits = s_bits[hq], hw = hbits & PG_HB_WBR_MASK, heb = s_eoff[hq], hee = s_eoff[hq + 1];
			const uint64_t hkey = s_key[hq];
			bool w_lose = false; uint64_t w_key = 0; int32_t w_p = INT32_MAX;
#pragma unroll 1
			for (int side = 0; side < 2; ++side) {
#pragma unroll 1
				for (int32_t base = 1;; base += 32) {
					const int32_t p = side == 0? h - base - lane : h + base + lane;
					bool inwin = p >= 0 && p < c.n;
					int32_t pcs = 0, pce = 0;
					if (inwin) {
						inwin = T.ctg(p) == hctg;
						if (inwin) { pcs = T.cs(p); pce = T.ce(p); inwin = side == 0? T.pmax(p) > hcs : pcs < hce; }
					}
					if (inwin) {
						const uint32_t pbits = T.bits(p);
						const int32_t pgene = T.gidp(p) >> 1;
						const uint64_t pkey = T.key(p);
						const bool same_gene = pgene == hgene;
						bool ok = !(pbits & PG_HB_FLT) && pce > hcs;
						if (MODE == SW_ISO) ok = ok && same_gene;
						if (o.check_strand) ok = ok && !((pbits ^ hbits) & PG_HB_REV);
						const uint32_t pw = pbits & PG_HB_WBR_MASK;
						const bool key_rule = MODE == SW_ISO || same_gene || hw == pw;
						bool loses = key_rule? hkey < pkey : hw > pw;
						if (ok && key_rule && hkey == pkey) { const int32_t r0 = c.rank[h], r1 = c.rank[p]; loses = side == 0? (r0 > r1) : !(r1 > r0); }
						if (ok && loses) {
							const int32_t need = related_threshold(MODE == SW_ISO || same_gene, hcds, T.cds(p), o.min_ov_ratio);
							const uint32_t pb = T.eoff(p), pe = T.eoff(p + 1);
							const bool pf = side == 0;                                   // the partner precedes the parked hit
							const bool rel = pair_overlap<true>(s_ex, elo, ehi, c.ex, pf? pb : heb, pf? pe : hee, pf? pcs : hcs, pf? pce : hce, pf? heb : pb, pf? hee : pe, pf? hcs : pcs, pf? hce : pce, need) >= need;
							if (rel) {
								w_lose = true;
								if (pkey > w_key || (pkey == w_key && p < w_p)) { w_key = pkey; w_p = p; }
							}
						}
					}
					if (!__any_sync(0xffffffffu, inwin && lane == 31)) break;   // the window predicate is monotone in distance: lane 31 is the farthest
				}
			}
			// warp reduction: any loss; largest key, then smallest index (overlap.c:150,153)
			w_lose = __any_sync(0xffffffffu, w_lose);
#pragma unroll
			for (int sft = 16; sft > 0; sft >>= 1) {
				const uint64_t ok2 = __shfl_xor_sync(0xffffffffu, w_key, sft); const int32_t op = __shfl_xor_sync(0xffffffffu, w_p, sft);
				if (ok2 > w_key || (ok2 == w_key && op < w_p)) { w_key = ok2; w_p = op; }
			}
			if (lane == 0) {
				if (MODE == SW_ISO) bits_out[h] = w_lose? (hbits | PG_HB_FLT_OV | PG_HB_FLT) : hbits;
				else if (MODE == SW_SHADOW) {
					uint32_t b = hbits;
					if (!(b & HB_FIRST)) b &= ~PG_HB_SHADOW;
					if (w_lose) b |= PG_HB_SHADOW;
					bits_out[h] = b;
					pdom_out[h] = (w_lose && w_key > 0)? c.pid[w_p] : -1;
				} else {
					if (!NEWP0) bits_out[h] = hbits;
					int32_t pd = -1, sd = -1;
					if (w_lose && w_key > 0) {
						pd = c.pid[w_p];
						const bool pf = w_p < h;
						const uint32_t wb2 = T.eoff(w_p), we2 = T.eoff(w_p + 1);
						const int32_t wcs = T.cs(w_p), wce = T.ce(w_p);
						const int32_t ov = pair_overlap<false>(s_ex, elo, ehi, c.ex, pf? wb2 : heb, pf? we2 : hee, pf? wcs : hcs, pf? wce : hce, pf? heb : wb2, pf? hee : we2, pf? hcs : wcs, pf? hce : wce, 0);
						sd = (int32_t)(c.sori[h] * (1.0 - (double)ov / hcds) + c.sori[w_p] * ((double)ov / T.cds(w_p)) + .499);   // overlap.c:170
					}
					pdom_out[h] = pd; sdom_out[h] = sd;
				}
			}
		}
		}
	}
}
// [emu-end sweep]

// ------------------------------------------------------------------------------------------------
// rest of the load-time block (hit.c:107-146)
// ------------------------------------------------------------------------------------------------
__global__ void k_chain_mark(const int32_t *__restrict__ ctg, const int32_t *__restrict__ ctg2g, const int32_t *__restrict__ pid, const uint32_t *__restrict__ bits, int32_t n, int32_t n_prot, uint8_t *t_kept)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (!(bits[i] & PG_HB_FLT_OV)) t_kept[(size_t)ctg2g[ctg[i]] * n_prot + pid[i]] = 1;   // hit.c:136-138
}

__global__ void k_chain_apply(const int32_t *__restrict__ ctg, const int32_t *__restrict__ ctg2g, const int32_t *__restrict__ pdom0, uint32_t *bits, int32_t n, int32_t n_prot, const uint8_t *__restrict__ t_kept)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int32_t p0 = pdom0[i];
	if (p0 >= 0 && !t_kept[(size_t)ctg2g[ctg[i]] * n_prot + p0]) bits[i] |= PG_HB_FLT | PG_HB_FLT_CHAIN;   // hit.c:141-142
}

// hit.c:112-118 as one atomicMax: a positive score wins by value, first in order on ties; any negative score always
// replaces the incumbent in the reference's signed-vs-unsigned compare, so the LAST negative in order wins outright.
__global__ void k_subopt_best(const int32_t *__restrict__ ctg, const int32_t *__restrict__ ctg2g, const int32_t *__restrict__ gidp, const int32_t *__restrict__ sadj,
	const int32_t *__restrict__ rank, const uint32_t *__restrict__ bits, int32_t n, int32_t n_gene, unsigned long long *t_best)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if ((bits[i] & PG_HB_FLT) || rank[i] > 0) return;
	int32_t s = sadj[i];
	if (s == 0) return;
	unsigned long long enc = s > 0? ((unsigned long long)(uint32_t)s << 32 | (0xffffffffu - (uint32_t)i)) : (1ull << 63 | (uint32_t)i);
	atomicMax(&t_best[(size_t)ctg2g[ctg[i]] * n_gene + (gidp[i] >> 1)], enc);
}

__global__ void k_subopt_apply(const int32_t *__restrict__ ctg, const int32_t *__restrict__ ctg2g, const int32_t *__restrict__ gidp, const int32_t *__restrict__ pid,
	uint32_t *bits, int32_t n, int32_t n_gene, const unsigned long long *__restrict__ t_best)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t b = bits[i];
	if (b & PG_HB_FLT) return;
	unsigned long long e = t_best[(size_t)ctg2g[ctg[i]] * n_gene + (gidp[i] >> 1)];
	int32_t keep = 0;                                                                  // hit.c:111,123: no winner -> protein 0
	if (e) { uint32_t w = (e >> 63)? (uint32_t)e : 0xffffffffu - (uint32_t)e; keep = pid[w]; }
	if (pid[i] != keep) bits[i] = b | PG_HB_FLT | PG_HB_FLT_SUB;
}

// ------------------------------------------------------------------------------------------------
// pg_post_process pieces (hit.c:153-247)
// ------------------------------------------------------------------------------------------------
__global__ void k_cap_max(const int32_t *__restrict__ pid, const int32_t *__restrict__ sori, int32_t n, int32_t *prot_max)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) atomicMax(&prot_max[pid[i]], sori[i]);                                  // hit.c:231-238 (floor 0 = initial value)
}
__global__ void k_cap_apply(const int32_t *__restrict__ pid, int32_t *sdom, int32_t n, const int32_t *__restrict__ prot_max)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { int32_t m = prot_max[pid[i]]; if (sdom[i] > m) sdom[i] = m; }         // hit.c:243-244
}

// per-protein sums over rank-0 unfiltered hits: z = sum(score_adj<<32|1) (hit.c:201), c[w]/s[w] (hit.c:163-167)
__global__ void k_prot_stats(const int32_t *__restrict__ pid, const int32_t *__restrict__ sadj, const int32_t *__restrict__ sori, const int32_t *__restrict__ rank,
	const uint32_t *__restrict__ bits, const uint32_t *__restrict__ eoff, int32_t n, unsigned long long *z, unsigned long long *cs4)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if ((bits[i] & PG_HB_FLT) || rank[i] != 0) return;
	int32_t p = pid[i];
	atomicAdd(&z[p], (unsigned long long)(long long)sadj[i] << 32 | 1ull);
	int w = (eoff[i + 1] - eoff[i]) != 1;
	atomicAdd(&cs4[(size_t)p * 4 + w], 1ull);                                          // c[w]
	atomicAdd(&cs4[(size_t)p * 4 + 2 + w], (unsigned long long)(long long)sori[i]);    // s[w]
}

__global__ void k_pseudo_joint(const int32_t *__restrict__ pid, uint32_t *bits, const uint32_t *__restrict__ eoff, int32_t n, const unsigned long long *__restrict__ cs4,
	int32_t n_genome_total, double min_vertex_ratio, int32_t drop_sgl)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t b = bits[i];
	if ((b & (PG_HB_FLT | PG_HB_PSEUDO)) || eoff[i + 1] - eoff[i] != 1) return;         // hit.c:175-176
	const unsigned long long *q = &cs4[(size_t)pid[i] * 4];
	int32_t c0 = (int32_t)q[0], c1 = (int32_t)q[1];
	long long s0 = (long long)q[2], s1 = (long long)q[3];
	double thr = n_genome_total * min_vertex_ratio;
	bool ps = false;
	if (c1 > 0 && c1 >= thr && ((double)s1 / c1) / ((double)s0 / c0) >= 0.99) ps = true;   // hit.c:176-179
	else if ((c1 == 0 || c1 <= thr) && drop_sgl) ps = true;                           // hit.c:180-181
	if (ps) bits[i] = b | PG_HB_PSEUDO;
}

// ------------------------------------------------------------------------------------------------
// generic flag passes (pgpriv.h:109-116 and friends)
// ------------------------------------------------------------------------------------------------
enum { SF_PSEUDO = 0, SF_SHADOW = 1, SF_WBR2 = 2 };
__global__ void k_set_filter(uint32_t *bits, int32_t n, int mode)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t b = bits[i];
	bool hit = mode == SF_PSEUDO? (b & PG_HB_PSEUDO) != 0 : mode == SF_SHADOW? (b & PG_HB_SHADOW) != 0 : ((b & PG_HB_WBR_MASK) >> PG_HB_WBR_SHIFT) == 2;
	if (hit && !(b & PG_HB_FLT)) bits[i] = b | PG_HB_FLT;
}

// graph.c:61-69 + PG_SET_FILTER(vtx == 0)
__global__ void k_flag_vtx(const int32_t *__restrict__ gidp, uint32_t *bits, int32_t n, const int32_t *__restrict__ g2s, int set_filter)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t b = bits[i];
	if (g2s[gidp[i] >> 1] >= 0) b |= PG_HB_VTX; else { b &= ~PG_HB_VTX; if (set_filter) b |= PG_HB_FLT; }
	bits[i] = b;
}

__global__ void k_flag_unfiltered(const uint32_t *__restrict__ bits, int32_t n, uint32_t *flag)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = (bits[i] & PG_HB_FLT)? 0u : 1u;
}

// ------------------------------------------------------------------------------------------------
// active set W: the hits that survived the load-time block, same order, compact columns
// ------------------------------------------------------------------------------------------------
struct ActCols {
	int32_t *ctg, *cs, *ce, *cm, *pid, *gidp, *sadj, *rank, *sori, *sdom, *pdom0, *pdom, *cds, *pmax, *full;
	uint32_t *bits, *bits_alt, *nex, *eoff;
	uint64_t *key;
};

__global__ void k_build_active(FullCols f, const uint32_t *__restrict__ pos, int32_t n, ActCols a)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t p = pos[i];
	if (pos[i + 1] == p) return;                            // filtered
	a.ctg[p] = f.ctg[i]; a.cs[p] = f.cs[i]; a.ce[p] = f.ce[i]; a.cm[p] = f.cm[i]; a.pid[p] = f.pid[i]; a.gidp[p] = f.gidp[i];
	a.sadj[p] = f.sadj[i]; a.rank[p] = f.rank[i]; a.sori[p] = f.sori[i]; a.sdom[p] = f.sdom[i]; a.pdom0[p] = f.pdom0[i]; a.pdom[p] = -1;
	a.cds[p] = f.cds[i]; a.full[p] = i; a.bits[p] = f.bits[i]; a.nex[p] = f.nex[i]; a.key[p] = f.key[i];
}

// shrink the active set to its unfiltered members (same order); pos = exclusive scan of the "still unfiltered" flags
__global__ void k_recompact_active(ActCols a, const uint32_t *__restrict__ pos, int32_t n, ActCols b, int32_t *old_of_new)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t p = pos[i];
	if (pos[i + 1] == p) return;
	b.ctg[p] = a.ctg[i]; b.cs[p] = a.cs[i]; b.ce[p] = a.ce[i]; b.cm[p] = a.cm[i]; b.pid[p] = a.pid[i]; b.gidp[p] = a.gidp[i];
	b.sadj[p] = a.sadj[i]; b.rank[p] = a.rank[i]; b.sori[p] = a.sori[i]; b.sdom[p] = a.sdom[i]; b.pdom0[p] = a.pdom0[i]; b.pdom[p] = a.pdom[i];
	b.cds[p] = a.cds[i]; b.full[p] = a.full[i]; b.bits[p] = a.bits[i]; b.nex[p] = a.nex[i]; b.key[p] = a.key[i];
	old_of_new[p] = i;
}
// the (contig, cm) permutation of the shrunken set: keep the survivors of the old permutation, renumbered
__global__ void k_flag_alive_cm(const uint32_t *__restrict__ pos, const uint32_t *__restrict__ by_cm, int32_t n, uint32_t *flag)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) { uint32_t i = by_cm[k]; flag[k] = pos[i + 1] != pos[i]? 1u : 0u; }
}
__global__ void k_renumber_cm(const uint32_t *__restrict__ pos, const uint32_t *__restrict__ by_cm, const uint32_t *__restrict__ pos_cm, int32_t n, uint32_t *out)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n && pos_cm[k + 1] != pos_cm[k]) out[pos_cm[k]] = pos[by_cm[k]];
}

__global__ void k_copy_exons(const int2 *__restrict__ ex_src, const uint32_t *__restrict__ eoff_src, const int32_t *__restrict__ full, const uint32_t *__restrict__ eoff_dst, int32_t n, int2 *ex_dst)
{
	int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	uint32_t s = eoff_src[full[p]], b = eoff_dst[p], e = eoff_dst[p + 1];
	for (uint32_t k = 0; k < e - b; ++k) ex_dst[b + k] = ex_src[s + k];
}

__global__ void k_make_cm_keys(const int32_t *__restrict__ ctg, const int32_t *__restrict__ cm, int32_t n, uint64_t *key)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) key[i] = (uint64_t)(uint32_t)ctg[i] << 32 | (uint32_t)cm[i];
}

// write the active set's state back into the full columns
__global__ void k_scatter_back(ActCols a, int32_t n, FullCols f)
{
	int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	int32_t i = a.full[p];
	f.bits[i] = a.bits[p]; f.pdom[i] = a.pdom[p]; f.sdom[i] = a.sdom[p];
}

// ------------------------------------------------------------------------------------------------
// vertex selection inputs (vertex.c:28-51)
// ------------------------------------------------------------------------------------------------
// t_dom[gene*G+g] = 1 if an unshadowed rank-0 unfiltered hit exists; t_first = smallest index of a shadowed one
__global__ void k_vtx_mark(const int32_t *__restrict__ ctg, const int32_t *__restrict__ ctg2g, const int32_t *__restrict__ gidp, const int32_t *__restrict__ rank,
	const uint32_t *__restrict__ bits, int32_t n, int32_t G, uint8_t *t_dom, int32_t *t_first)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t b = bits[i];
	if ((b & PG_HB_FLT) || rank[i] != 0) return;
	size_t k = (size_t)(gidp[i] >> 1) * G + ctg2g[ctg[i]];                              // gene-major: the host election reads rows
	if (b & PG_HB_SHADOW) atomicMin(&t_first[k], i); else t_dom[k] = 1;
}
// aux[k] = -2 (dominant) | gene that shadows it | -1 (absent)
__global__ void k_vtx_resolve(const uint8_t *__restrict__ t_dom, const int32_t *__restrict__ t_first, const int32_t *__restrict__ pdom, const int32_t *__restrict__ prot_gid, size_t n_tab, int32_t *aux, uint32_t *err)
{
	size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_tab) return;
	int32_t v = -1;
	if (t_dom[k]) v = -2;
	else if (t_first[k] != INT32_MAX) { int32_t pd = pdom[t_first[k]]; if (pd < 0) { atomicOr(err, 16u); pd = 0; } v = prot_gid[pd]; }   // vertex.c:38 assert
	aux[k] = v;
}

// What the greedy election (vertex.c:52-80) needs of the [gene][genome] table, compact: per gene the number of genomes where it is dominant /
// only shadowed (vertex.c:49-50), and one record per "shadowed" cell -- gene g, genome j, the gene t that shadows it there, and whether t is
// dominant in j (the only cells the election ever reads: an accepted gene claims the genes that shadow it, vertex.c:74-77).
struct VtxCell { int32_t g, j, t, t_dom; };
__global__ void k_vtx_cells(const int32_t *__restrict__ aux, size_t n_tab, int32_t G, int32_t genome_begin, int32_t *cnt, VtxCell *cells, uint32_t cap, uint32_t *n_cells)
{
	size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_tab) return;
	const int32_t v = aux[k];
	if (v == -1) return;
	const int32_t g = (int32_t)(k / G), j = (int32_t)(k % G);
	if (v == -2) { atomicAdd(&cnt[2 * g], 1); return; }
	atomicAdd(&cnt[2 * g + 1], 1);
	const uint32_t id = atomicAdd(n_cells, 1u);
	if (id < cap) { VtxCell c; c.g = g; c.j = genome_begin + j; c.t = v; c.t_dom = aux[(size_t)v * G + j] == -2? 1 : 0; cells[id] = c; }
}

// ------------------------------------------------------------------------------------------------
// arc generation (graph.c:87-177)
// ------------------------------------------------------------------------------------------------
__global__ void k_flag_survivor_cm(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ by_cm, int32_t n, uint32_t *flag)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) flag[k] = (bits[by_cm[k]] & (PG_HB_FLT | PG_HB_SHADOW))? 0u : 1u;       // graph.c:109
}
__global__ void k_flag_survivor_cs(const uint32_t *__restrict__ bits, int32_t n, uint32_t *flag)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) flag[k] = (bits[k] & (PG_HB_FLT | PG_HB_SHADOW))? 0u : 1u;              // branch.c:20
}
__global__ void k_compact_idx(const uint32_t *__restrict__ pos, const uint32_t *__restrict__ src, int32_t n, uint32_t *out)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n && pos[k + 1] != pos[k]) out[pos[k]] = src? src[k] : (uint32_t)k;
}

// (DevCounts, PGX_GS, ArcPartial, ArcDev, SegDev, GenePos, BranchOpt: pgx_graphtypes.cuh)
__global__ void k_copy_u32(uint32_t *dst, const uint32_t *src) { *dst = *src; }
__global__ void k_set_u32(uint32_t *dst, uint32_t v) { *dst = v; }

struct ArcEmit {
	const uint32_t *sv; int32_t cap;          // survivors in (contig, cm) order (indices into the active set); cap = capacity (active hits)
	const int32_t *ctg, *cm, *gidp, *sori, *sdom, *pdom0; const uint32_t *bits;
	const int32_t *g2s, *prot_gid, *ctg2g;
	int32_t use_ori, n_seg, gbits;         // n_seg: the segment count at vertex selection (fixed key/bitmap stride for the whole run)
};
__device__ __forceinline__ int32_t arc_score(const ArcEmit &e, uint32_t h)   // graph.c:82-85
{
	int32_t so = e.sori[h], sd = e.sdom[h], p0 = e.pdom0[h];
	return (e.use_ori || so > sd || p0 < 0 || e.g2s[e.prot_gid[p0]] >= 0)? so : sd;
}
// record t>0 with a predecessor on the same contig yields two keyed records (graph.c:115-120); others get the all-ones key
// key = v * 2n_seg + w;  payload = (local genome, dist, s1, s2).  Records are emitted in genome order and the sort is
// stable, so inside one key the records of a genome stay together -- no genome bits in the key, one radix pass fewer.
__global__ void k_arc_emit(ArcEmit e, const DevCounts *__restrict__ dc, uint64_t *key, int32_t *gen, int32_t *dist, int32_t *s1, int32_t *s2, uint32_t *seg_seen, int32_t *seg_ng, int32_t *seg_tot, uint32_t *tie_cnt, uint32_t *err)
{
	const int32_t m = (int32_t)dc->m_sv;
	PGX_GS(t, e.cap) {
		uint64_t k0 = ~0ull, k1 = ~0ull;
		int32_t d = 0, sa = 0, sp = 0, gl = 0;
		if (t < m) {
			uint32_t h = e.sv[t];
			int32_t sid = e.g2s[e.gidp[h] >> 1];
			if (sid < 0) { atomicOr(err, 32u); sid = 0; }                               // graph.c:111 assert
			int32_t c = e.ctg[h];
			gl = e.ctg2g[c];
			// seg occupancy (graph.c:113,125-126): first sighting of (genome, segment) bumps n_genome
			{
				size_t bit = (size_t)gl * e.n_seg + sid;
				uint32_t old = atomicOr(&seg_seen[bit >> 5], 1u << (bit & 31));
				if (!(old >> (bit & 31) & 1u)) atomicAdd(&seg_ng[sid], 1);
				atomicAdd(&seg_tot[sid], 1);
			}
			if (t > 0) {
				uint32_t hp = e.sv[t - 1];
				if (e.ctg[hp] == c) {
					uint32_t w = (uint32_t)sid << 1 | (e.bits[h] & PG_HB_REV);
					uint32_t v = (uint32_t)e.g2s[e.gidp[hp] >> 1] << 1 | (e.bits[hp] & PG_HB_REV);
					uint64_t nv = 2ull * e.n_seg;
					d = e.cm[h] - e.cm[hp];
					if (d == 0) atomicAdd(tie_cnt, 1u);
					sp = arc_score(e, hp); sa = arc_score(e, h);
					k0 = (uint64_t)v * nv + w;
					k1 = (uint64_t)(w ^ 1) * nv + (v ^ 1);
				}
			}
		}
		key[2 * (size_t)t] = k0; gen[2 * (size_t)t] = gl; dist[2 * (size_t)t] = d; s1[2 * (size_t)t] = sp; s2[2 * (size_t)t] = sa;
		key[2 * (size_t)t + 1] = k1; gen[2 * (size_t)t + 1] = gl; dist[2 * (size_t)t + 1] = d; s1[2 * (size_t)t + 1] = sa; s2[2 * (size_t)t + 1] = sp;
	}
}


__global__ void k_arc_heads(const uint64_t *__restrict__ key, int32_t n, uint32_t *flag)
{
	int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	uint64_t k = key[p];
	flag[p] = (k != ~0ull && (p == 0 || key[p - 1] != k))? 1u : 0u;
}

// one WARP per arc key: lanes stride over the key's records (sorted by key, then genome); a lane that sits on the first
// record of a (key, genome) run folds that run (graph.c:128-145: count, mean distance, max scores) and adds it to its
// private sums; a shuffle reduction finishes the key (the additive part of graph.c:153-163).
__global__ void k_arc_reduce(const uint64_t *__restrict__ key, const uint32_t *__restrict__ rec, const int32_t *__restrict__ gen, const int32_t *__restrict__ dist, const int32_t *__restrict__ s1, const int32_t *__restrict__ s2,
	int32_t n, int32_t n_seg, const uint32_t *__restrict__ start, const DevCounts *__restrict__ dc, ArcPartial *out, int32_t *arc_of_rec)
{
	const int32_t n_part = (int32_t)dc->n_part, lane = threadIdx.x & 31, n_warps = (gridDim.x * blockDim.x) >> 5;
	for (int32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_part; g += n_warps) {
		const int32_t b = start[g];
		int32_t e = g + 1 < n_part? (int32_t)start[g + 1] : n;
		const uint64_t xk = key[b];
		long long sum_dist = 0, sum_s1 = 0, sum_s2 = 0;
		int32_t n_genome = 0, tot = 0;
		for (int32_t q = b + lane; q < e; q += 32) {
			if (key[q] != xk) break;                            // only the last group can run into the "no arc" tail
			arc_of_rec[rec[q]] = g;                             // which arc a record went into (k_branch_hits, single-GPU)
			const int32_t gq = gen[rec[q]];
			if (q > b && gen[rec[q - 1]] == gq) continue;       // not the first record of its (key, genome) run
			unsigned long long dsum = 0; int32_t m1 = 0, m2 = 0, cnt = 0;
			for (int32_t r = q; r < e && key[r] == xk; ++r) {
				uint32_t o = rec[r];
				if (gen[o] != gq) break;
				dsum += (unsigned long long)(long long)dist[o];
				int32_t a1 = s1[o], a2 = s2[o];
				m1 = m1 > a1? m1 : a1; m2 = m2 > a2? m2 : a2;
				++cnt;
			}
			int32_t dg = (int32_t)((double)dsum / cnt + .499);                           // graph.c:141
			n_genome += 1; tot += cnt;
			sum_dist += (long long)((unsigned long long)(long long)dg * (unsigned long long)(long long)cnt);   // graph.c:160
			sum_s1 += m1; sum_s2 += m2;
		}
		for (int s = 16; s > 0; s >>= 1) {
			sum_dist += __shfl_xor_sync(0xffffffffu, sum_dist, s); sum_s1 += __shfl_xor_sync(0xffffffffu, sum_s1, s); sum_s2 += __shfl_xor_sync(0xffffffffu, sum_s2, s);
			n_genome += __shfl_xor_sync(0xffffffffu, n_genome, s); tot += __shfl_xor_sync(0xffffffffu, tot, s);
		}
		if (lane == 0) {
			ArcPartial a;
			uint64_t nv = 2ull * n_seg, v = xk / nv, w = xk % nv;
			a.x = v << 32 | w; a.sum_dist = sum_dist; a.sum_s1 = sum_s1; a.sum_s2 = sum_s2; a.n_genome = n_genome; a.tot_cnt = tot;
			out[g] = a;
		}
	}
}

// merge partials that share a key (multi-GPU: the same arc seen by several ranks), input sorted by x
__global__ void k_partial_heads(const ArcPartial *__restrict__ in, int32_t n, uint32_t *flag)
{
	int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n) flag[p] = (p == 0 || in[p - 1].x != in[p].x)? 1u : 0u;
}
__global__ void k_arc_finalize(const ArcPartial *__restrict__ in, int32_t n, const uint32_t *__restrict__ pos, ArcDev *arc)
{
	int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n || pos[p + 1] == pos[p]) return;
	ArcPartial a = in[p];
	for (int32_t q = p + 1; q < n && in[q].x == a.x; ++q) {
		a.sum_dist += in[q].sum_dist; a.sum_s1 += in[q].sum_s1; a.sum_s2 += in[q].sum_s2; a.n_genome += in[q].n_genome; a.tot_cnt += in[q].tot_cnt;
	}
	arc[pos[p]] = arc_finalize(a);
}

// single-GPU: every partial already is a complete arc (keys are unique after the reduction)
__global__ void k_arc_finalize_direct(const ArcPartial *__restrict__ in, const DevCounts *__restrict__ dc, ArcDev *arc)
{
	PGX_GS(p, dc->n_part) {
		arc[p] = arc_finalize(in[p]);
	}
}

// ------------------------------------------------------------------------------------------------
// small graph edits (graph.c:179-263)
// ------------------------------------------------------------------------------------------------

__global__ void k_arc_index(const ArcDev *__restrict__ arc, const DevCounts *__restrict__ dc, unsigned long long *idx)   // graph.c:202-211
{
	const int32_t n_arc = (int32_t)dc->n_arc;
	PGX_GS(p, n_arc) {
		uint32_t v = (uint32_t)(arc[p].x >> 32);
		if (p > 0 && (uint32_t)(arc[p - 1].x >> 32) == v) continue;
		int32_t q = p + 1;
		while (q < n_arc && (uint32_t)(arc[q].x >> 32) == v) ++q;
		idx[v] = (unsigned long long)p << 32 | (uint32_t)(q - p);
	}
}

__global__ void k_seg_store_occ(SegDev *seg, const DevCounts *__restrict__ dc, const int32_t *__restrict__ ng, const int32_t *__restrict__ tot)
{
	PGX_GS(s, dc->n_seg) { seg[s].n_genome = ng[s]; seg[s].tot_cnt = tot[s]; }
}

// graph.c:237-256; degrees are taken from the arc list itself (idx may be stale or absent, as in the reference). keep[] covers the
// capacity (segments at vertex selection); slots past the current count are 0
__global__ void k_seg_mark_del(SegDev *seg, int32_t cap, const DevCounts *__restrict__ dc, const ArcDev *__restrict__ arc, int32_t max_occ_total, int32_t max_degree, int32_t max_dist_loci, uint32_t *keep)
{
	const int32_t n_seg = (int32_t)dc->n_seg, n_arc = (int32_t)dc->n_arc;
	PGX_GS(s, cap) {
		if (s >= n_seg) { keep[s] = 0u; continue; }
		bool del = seg[s].tot_cnt > max_occ_total;
		for (int o = 0; o < 2 && !del; ++o) {
			uint64_t v = (uint64_t)(2 * s + o) << 32;
			int32_t lo = 0, hi = n_arc;                             // first arc with x >= v<<32
			while (lo < hi) { int32_t m = (lo + hi) >> 1; if (arc[m].x < v) lo = m + 1; else hi = m; }
			int32_t a = lo; hi = n_arc;
			uint64_t v1 = (uint64_t)(2 * s + o + 1) << 32;
			while (lo < hi) { int32_t m = (lo + hi) >> 1; if (arc[m].x < v1) lo = m + 1; else hi = m; }
			if (lo - a > max_degree) del = true;
		}
		int32_t m = seg[s].n_dist_loci[0] > seg[s].n_dist_loci[1]? seg[s].n_dist_loci[0] : seg[s].n_dist_loci[1];
		if (m > max_dist_loci) del = true;
		keep[s] = del? 0u : 1u;
	}
}
__global__ void k_seg_compact(const SegDev *__restrict__ in, const DevCounts *__restrict__ dc, const uint32_t *__restrict__ pos, SegDev *out)
{
	PGX_GS(s, dc->n_seg) if (pos[s + 1] != pos[s]) out[pos[s]] = in[s];
}
__global__ void k_g2s_fill(const SegDev *__restrict__ seg, const DevCounts *__restrict__ dc, int32_t *g2s)
{
	PGX_GS(s, dc->n_seg) g2s[seg[s].gid] = s;
}
__global__ void k_arc_keep_min_cnt(const ArcDev *__restrict__ arc, int32_t n_arc, int32_t min_cnt, uint32_t *keep)
{
	int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n_arc) keep[p] = arc[p].n_genome >= min_cnt? 1u : 0u;                       // graph.c:195
}
__global__ void k_arc_compact(const ArcDev *__restrict__ in, int32_t n_arc, const uint32_t *__restrict__ pos, ArcDev *out)
{
	int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n_arc && pos[p + 1] != pos[p]) out[pos[p]] = in[p];
}

// ------------------------------------------------------------------------------------------------
// branch filtering (branch.c)
// ------------------------------------------------------------------------------------------------
// last survivor of each gene per genome in (contig, cs) order (branch.c:6-29): atomicMax over the active index
__global__ void k_pos_last(const int32_t *__restrict__ ctg, const int32_t *__restrict__ ctg2g, const int32_t *__restrict__ gidp, const uint32_t *__restrict__ surv_pos, int32_t n, int32_t G, int32_t *t_last)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || surv_pos[i + 1] == surv_pos[i]) return;
	atomicMax(&t_last[(size_t)(gidp[i] >> 1) * G + ctg2g[ctg[i]]], i);                 // layout [gene][genome]: n_local reads are contiguous
}
__global__ void k_pos_fill(const int32_t *__restrict__ t_last, size_t n_tab, int32_t G, const int32_t *__restrict__ ctg, const int32_t *__restrict__ cm, const uint32_t *__restrict__ surv_pos,
	const int32_t *__restrict__ goff_act, GenePos *pos)
{
	size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_tab) return;
	int32_t i = t_last[k];
	GenePos g; g.cid = -1; g.ord = 0; g.cm = 0; g.pad = 0;
	if (i >= 0) { int32_t gl = (int32_t)(k % G); g.cid = ctg[i]; g.cm = cm[i]; g.ord = (int32_t)(surv_pos[i] - surv_pos[goff_act[gl]]); }
	pos[k] = g;
}

// Survivors that share one (contig, cs) with the survivor before them: counted (pgx_stats_t.tie[1]), listed
// (list[] = position t in the survivor list of the LATER member of each tied pair) and their genes marked in the position table
// (GenePos.pad) -- the reference numbers such hits in the order its unstable sort chain left them in (branch.c:18-25), which can
// differ from ours by a permutation inside the group.  long_run: a group of more than 9 hits (the sensitivity window of
// k_branch_nlocal assumes ordinals move by at most 8).
__global__ void k_cs_tie_mark(const int32_t *__restrict__ ctg, const int32_t *__restrict__ cs, const int32_t *__restrict__ gidp, const int32_t *__restrict__ ctg2g, int32_t G,
	const uint32_t *__restrict__ sv_cs, const DevCounts *__restrict__ dc, const int32_t *__restrict__ t_last, GenePos *pos, uint32_t *tie_cnt, uint32_t *round_cnt, uint32_t *list, uint32_t *long_run)
{
	PGX_GS(t, dc->m_cs) {
		if (t < 1) continue;
		const uint32_t a = sv_cs[t - 1], b = sv_cs[t];
		if (ctg[a] != ctg[b] || cs[a] != cs[b]) continue;
		atomicAdd(tie_cnt, 1u);
		list[atomicAdd(round_cnt, 1u)] = (uint32_t)t;
		const size_t ka = (size_t)(gidp[a] >> 1) * G + ctg2g[ctg[a]], kb = (size_t)(gidp[b] >> 1) * G + ctg2g[ctg[b]];
		if (t_last[ka] == (int32_t)a) pos[ka].pad = 1;
		if (t_last[kb] == (int32_t)b) pos[kb].pad = 1;
		if (t >= 9) { const uint32_t z = sv_cs[t - 9]; if (ctg[z] == ctg[b] && cs[z] == cs[b]) *long_run = 1u; }
	}
}

struct TieRec { int32_t t, a_prev, a, ctg, full_prev, full; };   // one tied pair: survivor-list position, active indices, contig, F indices
__global__ void k_tie_gather(const uint32_t *__restrict__ list, int32_t n, const uint32_t *__restrict__ sv_cs, const int32_t *__restrict__ ctg, const int32_t *__restrict__ full, TieRec *out)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	TieRec r;
	r.t = (int32_t)list[k]; r.a_prev = (int32_t)sv_cs[r.t - 1]; r.a = (int32_t)sv_cs[r.t];
	r.ctg = ctg[r.a]; r.full_prev = full[r.a_prev]; r.full = full[r.a];
	out[k] = r;
}

// positions t of the (contig, cm) survivor list whose survivor shares contig and cm with the one before it (Session::fix_cm_tied_survivors)
__global__ void k_cm_tie_list(const uint32_t *__restrict__ sv, const DevCounts *__restrict__ dc, const int32_t *__restrict__ ctg, const int32_t *__restrict__ cm, uint32_t *list, uint32_t *n)
{
	PGX_GS(t, dc->m_sv) {
		if (t < 1) continue;
		const uint32_t a = sv[t - 1], b = sv[t];
		if (ctg[a] == ctg[b] && cm[a] == cm[b]) list[atomicAdd(n, 1u)] = (uint32_t)t;
	}
}
__global__ void k_patch_u32(const int2 *__restrict__ patch, int32_t n, uint32_t *a)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) a[patch[k].x] = (uint32_t)patch[k].y;
}

// [lo, hi) of each listed contig in the (genome, contig, cs)-ordered full columns
__global__ void k_ctg_ranges(const int32_t *__restrict__ ids, int32_t n_ids, const int32_t *__restrict__ f_ctg, int32_t N, int2 *out)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_ids) return;
	const int32_t c = ids[k];
	int32_t lo = 0, hi = N;
	while (lo < hi) { const int32_t m = (lo + hi) >> 1; if (f_ctg[m] < c) lo = m + 1; else hi = m; }
	const int32_t b = lo;
	hi = N;
	while (lo < hi) { const int32_t m = (lo + hi) >> 1; if (f_ctg[m] <= c) lo = m + 1; else hi = m; }
	out[k] = make_int2(b, lo);
}

// ordinals of tied survivors, as the reference numbers them: only the hit that owns its gene's table entry is renumbered
__global__ void k_patch_ord(const int2 *__restrict__ patch, int32_t n, const int32_t *__restrict__ ctg, const int32_t *__restrict__ gidp, const int32_t *__restrict__ ctg2g, int32_t G,
	const int32_t *__restrict__ t_last, GenePos *pos)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const int32_t a = patch[k].x;
	const size_t key = (size_t)(gidp[a] >> 1) * G + ctg2g[ctg[a]];
	if (t_last[key] == a) pos[key].ord += patch[k].y;
}

// number of n_local values a vertex needs: all arc pairs (i <= k) of its out-list
__global__ void k_branch_count(const unsigned long long *__restrict__ idx, int32_t cap_vtx, const DevCounts *__restrict__ dc, uint32_t *cnt)
{
	const int32_t n_vtx = 2 * (int32_t)dc->n_seg;
	PGX_GS(v, cap_vtx) {
		uint32_t n = v < n_vtx? (uint32_t)idx[v] : 0u;
		cnt[v] = n >= 2? n * (n + 1) / 2 : 0u;
	}
}


__device__ __forceinline__ uint32_t tri_index(uint32_t i, uint32_t k, uint32_t n) { return i * n - i * (i - 1) / 2 + (k - i); }   // i <= k

// one WARP per table entry = one (vertex, arc pair): partial n_local (branch.c:31-46) over the LOCAL genomes.  (A warp per
// vertex leaves the few high-degree vertices -- 465 pairs at degree 30 -- as a serial tail: 264 us -> see profiles/README.md.)
__global__ void k_branch_nlocal(const unsigned long long *__restrict__ idx, const DevCounts *__restrict__ dc, const ArcDev *__restrict__ arc, const SegDev *__restrict__ seg, const GenePos *__restrict__ pos, int32_t G,
	const uint32_t *__restrict__ off, int32_t n_nl, BranchOpt o, int32_t *nl, uint32_t *ord_sensitive)
{
	const int32_t n_vtx = 2 * (int32_t)dc->n_seg, lane = threadIdx.x & 31, n_warps = (gridDim.x * blockDim.x) >> 5;
	for (int32_t ent = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ent < n_nl; ent += n_warps) {
		// vertex owning this entry: last v with off[v] <= ent (off is the exclusive scan of the per-vertex pair counts)
		int32_t lo = 0, hi = n_vtx;
		while (lo < hi) { int32_t m = (lo + hi) >> 1; if (off[m] <= (uint32_t)ent) lo = m + 1; else hi = m; }
		const int32_t v = lo - 1;
		const unsigned long long e = idx[v];
		const uint32_t n = (uint32_t)e;
		uint32_t r = (uint32_t)ent - off[v], i = 0;
		while (r >= n - i) { r -= n - i; ++i; }                 // row i of the upper triangle holds n - i entries (k = i .. n-1)
		const uint32_t k = i + r;
		const ArcDev *a = arc + (e >> 32);
		const GenePos *p1 = pos + (size_t)seg[(uint32_t)a[i].x >> 1].gid * G, *p2 = pos + (size_t)seg[(uint32_t)a[k].x >> 1].gid * G;
		int32_t c = 0;
		for (int32_t j = lane; j < G; j += 32) {
			GenePos x = p1[j], y = p2[j];
			if (x.cid < 0 || y.cid < 0) continue;
			if (!o.frag_mode && x.cid != y.cid) continue;
			long long dd = (long long)x.cm - y.cm; int32_t dcnt = x.ord - y.ord;
			if ((dd >= -o.local_dist && dd <= o.local_dist) || (dcnt >= -o.local_count && dcnt <= o.local_count)) ++c;
			// an ordinal of a cs-tied survivor (k_cs_tie_mark) decides here and is within reach of the threshold: the round is redone
			// with the reference's numbering of those hits (Session::fix_tied_ordinals)
			if ((x.pad | y.pad) && !(dd >= -o.local_dist && dd <= o.local_dist)) {
				const long long ac = dcnt < 0? -(long long)dcnt : dcnt;
				if (ac > (long long)o.local_count - 8 && ac <= (long long)o.local_count + 8) *ord_sensitive = 1u;
			}
		}
		for (int sft = 16; sft > 0; sft >>= 1) c += __shfl_xor_sync(0xffffffffu, c, sft);
		if (lane == 0) nl[ent] = c;
	}
}

// one thread per oriented vertex: branch.c:60-98 with the (summed) n_local table
__global__ void k_branch_decide(const unsigned long long *__restrict__ idx, const DevCounts *__restrict__ dc, ArcDev *arc, SegDev *seg, const uint32_t *__restrict__ off, const int32_t *__restrict__ nl, BranchOpt o, int32_t *grp_scratch)
{
	PGX_GS(v, 2 * (int32_t)dc->n_seg) {
	unsigned long long e = idx[v];
	int32_t n = (int32_t)(uint32_t)e;
	if (n < 2) continue;
	ArcDev *a = arc + (e >> 32);
	const int32_t *t = nl + off[v];
	int32_t best = 0;
	for (int32_t i = 0; i < n; ++i) best = best > a[i].s1? best : a[i].s1;
	for (int32_t i = 0; i < n; ++i) {
		double r = 1.0 - (double)a[i].s1 / best;
		if (r > o.branch_diff) {
			int32_t sum = 0;
			for (int32_t k = 0; k < n; ++k)
				if (a[k].s1 == best) sum += t[k < i? tri_index(k, i, n) : tri_index(i, k, n)];
			uint32_t wb = ((sum == 0 && r > o.branch_diff_dist) || r > o.branch_diff_cut)? 2u : 1u;   // branch.c:76-77
			a[i].bits = (a[i].bits & ~6u) | wb << 1;            // pg_arc_t: del:1, weak_br:2
		}
	}
	// branch.c:82-90: greedy grouping of the out-neighbours by locality; labels live in a scratch slice parallel to the arc list
	int32_t n_grp = 0;
	int32_t *grp = grp_scratch + (e >> 32);
	for (int32_t i = 0; i < n; ++i) grp[i] = 0;
	for (int32_t i = 0; i < n; ++i) {
		if (grp[i] == 0) grp[i] = ++n_grp;
		for (int32_t k = i + 1; k < n; ++k)
			if (grp[k] == 0 && t[tri_index(i, k, n)] > 0) grp[k] = grp[i];
	}
	seg[v >> 1].n_dist_loci[v & 1] = n_grp;
	}
}

// branch.c:108-145 + PG_SET_FILTER(weak_br == 2): survivors in (contig, cm) order look up the two arcs they sit on.
// Survivor t emitted records 2t (prev -> me) and 2t+1 (me^1 -> prev^1) in k_arc_emit; arc_of_rec (written by k_arc_reduce)
// says which arc each record was folded into, so no search is needed.  With arc_of_rec == nullptr (multi-GPU: arcs are
// re-numbered by the cross-rank merge) the arc is found by binary search on its key.
__device__ __forceinline__ uint32_t arc_weak(const ArcDev *__restrict__ arc, int32_t n_arc, uint64_t x)
{
	int32_t lo = 0, hi = n_arc;
	while (lo < hi) { int32_t m = (lo + hi) >> 1; if (arc[m].x < x) lo = m + 1; else hi = m; }
	return (lo < n_arc && arc[lo].x == x)? (arc[lo].bits >> 1 & 3u) : 0u;
}
__global__ void k_branch_hits(const uint32_t *__restrict__ sv, const DevCounts *__restrict__ dc, const int32_t *__restrict__ ctg, const int32_t *__restrict__ gidp, uint32_t *bits, const int32_t *__restrict__ g2s,
	const ArcDev *__restrict__ arc, const int32_t *__restrict__ arc_of_rec, uint32_t *n_changed)
{
	const int32_t m = (int32_t)dc->m_sv, n_arc = (int32_t)dc->n_arc;
	PGX_GS(t, m) {
		uint32_t h = sv[t], b = bits[h];
		const uint32_t b_in = b;
		int32_t c = ctg[h];
		uint32_t wb = (b & PG_HB_WBR_MASK) >> PG_HB_WBR_SHIFT;
		if (t > 0 && ctg[sv[t - 1]] == c) {                 // I am the later hit of (prev, me): arc (w^1 -> v^1), branch.c:130-132
			uint32_t x;
			if (arc_of_rec) x = arc[arc_of_rec[2 * (size_t)t + 1]].bits >> 1 & 3u;
			else {
				uint32_t hp = sv[t - 1];
				uint32_t w = (uint32_t)g2s[gidp[h] >> 1] << 1 | (b & PG_HB_REV), v = (uint32_t)g2s[gidp[hp] >> 1] << 1 | (bits[hp] & PG_HB_REV);
				x = arc_weak(arc, n_arc, (uint64_t)(w ^ 1) << 32 | (v ^ 1));
			}
			wb = wb > x? wb : x;
		}
		if (t + 1 < m && ctg[sv[t + 1]] == c) {             // I am the earlier hit of (me, next): arc (v -> w), branch.c:127-129
			uint32_t x;
			if (arc_of_rec) x = arc[arc_of_rec[2 * (size_t)(t + 1)]].bits >> 1 & 3u;
			else {
				uint32_t hn = sv[t + 1];
				uint32_t w = (uint32_t)g2s[gidp[h] >> 1] << 1 | (b & PG_HB_REV), wn = (uint32_t)g2s[gidp[hn] >> 1] << 1 | (bits[hn] & PG_HB_REV);
				x = arc_weak(arc, n_arc, (uint64_t)w << 32 | wn);
			}
			wb = wb > x? wb : x;
		}
		// only the weak_br bits (and flt) of my own word change; neighbours read rev only, which never changes
		b = (b & ~PG_HB_WBR_MASK) | wb << PG_HB_WBR_SHIFT;
		if (wb == 2) b |= PG_HB_FLT;                             // graph.c:309
		bits[h] = b;
		if (b != b_in) atomicAdd(n_changed, 1u);                 // (a round that changes no hit and deletes no segment leaves pg_gen_arc's input as it was)
	}
}

// ------------------------------------------------------------------------------------------------
// the way back: only what changed.  The host rebuilds pg_hit_t records in (contig, cs) order from its own copy.
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_out(const uint32_t *__restrict__ src, int32_t n, FullCols c, const int32_t *__restrict__ g2s, const int32_t *__restrict__ prot_rep, int32_t have_graph, OutHit *out)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	OutHit o;
	o.src = src[i]; o.rank = c.rank[i]; o.sdom = c.sdom[i]; o.pdom = c.pdom[i]; o.pdom0 = c.pdom0[i];
	uint32_t b = c.bits[i] & HB_ABI_MASK & ~(PG_HB_VTX | PG_HB_REP);
	if (prot_rep[c.pid[i]]) b |= PG_HB_REP;                                             // hit.c:219-224
	if (have_graph && g2s[c.gidp[i] >> 1] >= 0) b |= PG_HB_VTX;                         // graph.c:61-69 (last call)
	o.bits = b;
	out[i] = o;
}

// ------------------------------------------------------------------------------------------------
// walk tokens: what format.c:183-225 prints on the W lines, straight from the device-resident state.  The unfiltered hits in
// (genome, contig, cm) order -- by_cm already is that order of the active set, ties in (contig, cs) order -- each as (contig, gene << 1 | strand, lof).
// ------------------------------------------------------------------------------------------------
struct WalkToken { int32_t ctg; uint32_t gene_strand; int32_t lof; };                       // 12 bytes; ctg: the rank-local contig key (coff[genome] + cid)
__global__ void k_flag_walk_cm(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ by_cm, int32_t n, uint32_t *flag)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) flag[k] = (bits[by_cm[k]] & PG_HB_FLT)? 0u : 1u;                            // format.c:205,213
}
__global__ void k_walk_tokens(const uint32_t *__restrict__ pos, const uint32_t *__restrict__ by_cm, int32_t n, ActCols a, const uint32_t *__restrict__ src,
	const RawHit *__restrict__ raw, WalkToken *out)
{
	int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n || pos[k + 1] == pos[k]) return;
	const uint32_t p = by_cm[k];
	WalkToken t;
	t.ctg = a.ctg[p]; t.gene_strand = (uint32_t)(a.gidp[p] >> 1) << 1 | (a.bits[p] & PG_HB_REV);
	t.lof = raw[src[a.full[p]]].lof;
	out[pos[k]] = t;
}

// ------------------------------------------------------------------------------------------------
// odds and ends
// ------------------------------------------------------------------------------------------------
// (contig << 32 | coordinate) -> (contig << coord_bits | coordinate): same order, fewer significant bits, fewer radix passes
__global__ void k_pack_sort_key(uint64_t *key, int32_t n, int coord_bits)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { uint64_t k = key[i]; key[i] = (k >> 32) << coord_bits | (k & 0xffffffffull); }
}
__global__ void k_fill_i32(int32_t *a, int64_t n, int32_t v)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) a[i] = v;
}
__global__ void k_gather_u32(const uint32_t *__restrict__ table, const uint32_t *__restrict__ index, int32_t n, uint32_t *out)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = table[index[i]];
}
__global__ void k_seg_zero_loci(SegDev *seg, const DevCounts *__restrict__ dc)   // branch.c:56-57
{
	PGX_GS(s, dc->n_seg) seg[s].n_dist_loci[0] = seg[s].n_dist_loci[1] = 0;
}
__global__ void k_partial_keys(const ArcPartial *__restrict__ in, int32_t n, int32_t n_seg, uint64_t *key)
{
	int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n) key[p] = (in[p].x >> 32) * (2ull * n_seg) + (uint32_t)in[p].x;
}
__global__ void k_partial_gather(const ArcPartial *__restrict__ in, const uint32_t *__restrict__ perm, int32_t n, ArcPartial *out)
{
	int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n) out[p] = in[perm[p]];
}
// algorithmic bytes of one sweep launch: 52 + 8 n_exon per unfiltered hit, 4 per filtered hit
__global__ void k_sweep_bytes(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ eoff, int32_t n, unsigned long long *total)
{
	int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long b = 0;
	if (i < n) b = (bits[i] & PG_HB_FLT)? 4ull : 52ull + 8ull * (eoff[i + 1] - eoff[i]);
	for (int s = 16; s > 0; s >>= 1) b += __shfl_xor_sync(0xffffffffu, b, s);
	if ((threadIdx.x & 31) == 0 && b) atomicAdd(total, b);
}

} // namespace pgx
