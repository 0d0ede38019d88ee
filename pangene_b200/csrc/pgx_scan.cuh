// pgx_scan.cuh -- device-wide prefix sums and the segmented running maximum (hand-written, no CUB).
//
// Three launches per scan: (1) every 2048-element tile is scanned on its own and leaves one aggregate,
// (2) ONE block walks the aggregates (<= a few thousand) and turns them into tile prefixes, (3) tiles add
// their prefix.  Bandwidth: reads n, writes n, reads+writes n again (phase 3) -- these scans feed stream
// compactions that are a small part of a round, see DESIGN.md.
#pragma once
#include "pgx_util.cuh"
#include <cstring>

namespace pgx {

// [emu-begin scan]  (tests/emu/build.py compiles the kernels up to [emu-end scan] for the host: tests/test_prim_emu.py)
constexpr int SCAN_B = 256, SCAN_I = 8, SCAN_TILE = SCAN_B * SCAN_I;

// ---- exclusive sum of uint32 -------------------------------------------------------------------
// out has n_in+1 entries: out[i] = sum(in[0..i)), out[n_in] = grand total. in may alias out.
__global__ void k_scan_sum_tiles(const uint32_t *__restrict__ in, uint32_t *out, uint32_t *agg, int64_t n_in)
{
	__shared__ uint32_t sh[SCAN_B];
	const int tid = threadIdx.x;
	const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)tid * SCAN_I;
	uint32_t v[SCAN_I], sum = 0;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) { int64_t i = base + k; v[k] = i < n_in? in[i] : 0u; sum += v[k]; }
	sh[tid] = sum;
	__syncthreads();
	for (int d = 1; d < SCAN_B; d <<= 1) {
		uint32_t t = tid >= d? sh[tid - d] : 0u;
		__syncthreads();
		sh[tid] += t;
		__syncthreads();
	}
	uint32_t run = sh[tid] - sum;
	if (tid == SCAN_B - 1) agg[blockIdx.x] = sh[tid];
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) { int64_t i = base + k; if (i <= n_in) out[i] = run; run += v[k]; }
}

// one block: agg[i] <- exclusive prefix of agg (in place)
__global__ void k_scan_sum_agg(uint32_t *agg, int n)
{
	__shared__ uint32_t sh[SCAN_B];
	__shared__ uint32_t carry;
	const int tid = threadIdx.x;
	if (tid == 0) carry = 0;
	__syncthreads();
	for (int c0 = 0; c0 < n; c0 += SCAN_B) {
		int i = c0 + tid;
		uint32_t x = i < n? agg[i] : 0u;
		sh[tid] = x;
		__syncthreads();
		for (int d = 1; d < SCAN_B; d <<= 1) {
			uint32_t t = tid >= d? sh[tid - d] : 0u;
			__syncthreads();
			sh[tid] += t;
			__syncthreads();
		}
		uint32_t c = carry;
		if (i < n) agg[i] = c + sh[tid] - x;
		__syncthreads();
		if (tid == SCAN_B - 1) carry = c + sh[tid];
		__syncthreads();
	}
}

__global__ void k_scan_sum_add(uint32_t *out, const uint32_t *__restrict__ agg, int64_t n_out)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_out) out[i] += agg[i / SCAN_TILE];
}

// whole scan in ONE block (n_in up to a few hundred thousand): the many small per-round scans (segments, vertices, arc
// groups) are launch-latency bound, three launches each hurt more than the serial chunk loop
constexpr int SCAN_SMALL_B = 1024, SCAN_SMALL_MAX = 1 << 16;
__global__ void __launch_bounds__(SCAN_SMALL_B) k_scan_sum_small(const uint32_t *__restrict__ in, uint32_t *out, int32_t n_in)
{
	__shared__ uint32_t wsum[32];
	__shared__ uint32_t carry;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) carry = 0;
	__syncthreads();
	for (int32_t c0 = 0; c0 <= n_in; c0 += SCAN_SMALL_B) {
		int32_t i = c0 + tid;
		uint32_t x = i < n_in? in[i] : 0u, v = x;
		for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
		if (lane == 31) wsum[warp] = v;
		__syncthreads();
		if (warp == 0) {
			uint32_t w = wsum[lane], ws = w;
			for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, ws, d); if (lane >= d) ws += t; }
			wsum[lane] = ws - w;                       // exclusive prefix of the warp totals
		}
		__syncthreads();
		uint32_t base = carry + wsum[warp];
		if (i <= n_in) out[i] = base + v - x;
		__syncthreads();
		if (tid == SCAN_SMALL_B - 1) carry = base + v;
		__syncthreads();
	}
}

// ---- single-pass variant (decoupled look-back): one launch instead of three -------------------------------------------
// Every tile publishes (flag, value) in a 64-bit status word: AGG = "my own sum", PREFIX = "sum of everything up to and
// including me".  A tile adds up the words of its predecessors, 32 at a time (one per lane), until it meets a PREFIX.  Tile
// numbers come from an atomic ticket, so a tile only ever waits for tiles that are already running.
constexpr unsigned long long LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_MASK = 3ull << 62;
__global__ void __launch_bounds__(SCAN_B) k_scan_sum_lookback(const uint32_t *__restrict__ in, uint32_t *out, int64_t n_in, unsigned long long *status, uint32_t *ticket)
{
	__shared__ uint32_t sh[SCAN_B];
	__shared__ uint32_t s_tile, s_prefix;
	const int tid = threadIdx.x;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)tid * SCAN_I;
	uint32_t v[SCAN_I], sum = 0;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) { int64_t i = base + k; v[k] = i < n_in? in[i] : 0u; sum += v[k]; }
	sh[tid] = sum;
	__syncthreads();
	for (int d = 1; d < SCAN_B; d <<= 1) {
		uint32_t t = tid >= d? sh[tid - d] : 0u;
		__syncthreads();
		sh[tid] += t;
		__syncthreads();
	}
	const uint32_t tile_sum = sh[SCAN_B - 1];
	if (tid < 32) {
		if (tid == 0) { __threadfence(); atomicExch(&status[tile], (tile == 0? LB_PREFIX : LB_AGG) | tile_sum); }
		uint32_t prefix = 0;
		if (tile > 0) {
			int64_t look = (int64_t)tile - 1;
			for (;;) {
				const int64_t p = look - tid;
				unsigned long long w = LB_PREFIX;                    // lanes before the first tile: a zero-valued closed prefix
				if (p >= 0) { do { w = *(volatile unsigned long long*)&status[p]; } while ((w & LB_MASK) == 0); }
				const uint32_t closed = __ballot_sync(0xffffffffu, (w & LB_MASK) == LB_PREFIX);
				const int first = closed? __ffs(closed) - 1 : 32;   // nearest predecessor that already knows its prefix
				uint32_t add = tid <= first? (uint32_t)(w & 0xffffffffull) : 0u;
				for (int sft = 16; sft > 0; sft >>= 1) add += __shfl_xor_sync(0xffffffffu, add, sft);
				prefix += add;
				if (closed) break;
				look -= 32;
			}
			if (tid == 0) { __threadfence(); atomicExch(&status[tile], LB_PREFIX | (unsigned long long)(prefix + tile_sum)); }
		}
		if (tid == 0) s_prefix = prefix;
	}
	__syncthreads();
	uint32_t run = s_prefix + sh[tid] - sum;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) { int64_t i = base + k; if (i <= n_in) out[i] = run; run += v[k]; }
}

// [emu-end scan]
// scratch for the look-back variant: one status word per tile + the ticket, as uint32 elements
static inline size_t scan_scratch_elems(int64_t n_in) { return (size_t)((n_in + 1 + SCAN_TILE - 1) / SCAN_TILE) * 2 + 4; }

static void exclusive_sum_u32(const uint32_t *in, uint32_t *out, int64_t n_in, uint32_t *agg_scratch, cudaStream_t s)
{
	static const bool three_pass = []() { const char *e = getenv("PGX_SCAN"); return e && strcmp(e, "3pass") == 0; }();
	int64_t n_out = n_in + 1;
	if (n_in <= SCAN_SMALL_MAX) { PGX_LAUNCH(k_scan_sum_small, 1, SCAN_SMALL_B, 0, s, in, out, (int32_t)n_in); return; }
	int tiles = (int)((n_out + SCAN_TILE - 1) / SCAN_TILE);
	if (!three_pass) {
		// status words must start 8-byte aligned: scratch comes from 256-byte aligned arena/cache blocks
		unsigned long long *status = (unsigned long long*)agg_scratch;
		uint32_t *ticket = agg_scratch + (size_t)tiles * 2;
		PGX_CHECK(cudaMemsetAsync(agg_scratch, 0, ((size_t)tiles * 2 + 4) * 4, s));
		PGX_LAUNCH(k_scan_sum_lookback, tiles, SCAN_B, 0, s, in, out, n_in, status, ticket);
		return;
	}
	PGX_LAUNCH(k_scan_sum_tiles, tiles, SCAN_B, 0, s, in, out, agg_scratch, n_in);
	if (tiles > 1) {
		PGX_LAUNCH(k_scan_sum_agg, 1, SCAN_B, 0, s, agg_scratch, tiles);
		PGX_LAUNCH(k_scan_sum_add, grid_for(n_out, 256), 256, 0, s, out, agg_scratch, n_out);
	}
}

// ---- segmented inclusive running max -----------------------------------------------------------
// out[i] = max(val[k] : k <= i, seg[k] == seg[i]);  seg must be non-decreasing (sorted contig keys)
// [emu-begin segmax]
struct SegVal { int32_t seg, val; };
__device__ __forceinline__ SegVal seg_op(SegVal a, SegVal b) { SegVal r; r.seg = b.seg; r.val = (a.seg == b.seg && a.val > b.val)? a.val : b.val; return r; }

__global__ void k_segmax_tiles(const int32_t *__restrict__ seg, const int32_t *__restrict__ val, int32_t *out, SegVal *agg, int64_t n)
{
	__shared__ SegVal sh[SCAN_B];
	const int tid = threadIdx.x;
	const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)tid * SCAN_I;
	SegVal it[SCAN_I], tot;
	tot.seg = -1; tot.val = INT32_MIN;
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) {
		int64_t i = base + k;
		if (i < n) { it[k].seg = seg[i]; it[k].val = val[i]; } else { it[k].seg = -1; it[k].val = INT32_MIN; }
		tot = seg_op(tot, it[k]);
	}
	sh[tid] = tot;
	__syncthreads();
	for (int d = 1; d < SCAN_B; d <<= 1) {
		SegVal t, me = sh[tid];
		if (tid >= d) t = sh[tid - d];
		__syncthreads();
		if (tid >= d) sh[tid] = seg_op(t, me);
		__syncthreads();
	}
	SegVal run;
	if (tid > 0) run = sh[tid - 1]; else { run.seg = -1; run.val = INT32_MIN; }
	if (tid == SCAN_B - 1) agg[blockIdx.x] = sh[tid];
#pragma unroll
	for (int k = 0; k < SCAN_I; ++k) { int64_t i = base + k; run = seg_op(run, it[k]); if (i < n) out[i] = run.val; }
}

// one block: agg[i] <- INCLUSIVE scan of agg under seg_op (in place)
__global__ void k_segmax_agg(SegVal *agg, int n)
{
	__shared__ SegVal sh[SCAN_B];
	__shared__ SegVal carry;
	const int tid = threadIdx.x;
	if (tid == 0) { carry.seg = -1; carry.val = INT32_MIN; }
	__syncthreads();
	for (int c0 = 0; c0 < n; c0 += SCAN_B) {
		int i = c0 + tid;
		SegVal x; x.seg = -1; x.val = INT32_MIN;
		if (i < n) x = agg[i];
		// out-of-range lanes sit to the right of every valid one, so their garbage never flows into a valid result
		sh[tid] = x;
		__syncthreads();
		for (int d = 1; d < SCAN_B; d <<= 1) {
			SegVal t, me = sh[tid];
			if (tid >= d) t = sh[tid - d];
			__syncthreads();
			if (tid >= d) sh[tid] = seg_op(t, me);
			__syncthreads();
		}
		SegVal c = carry, r = seg_op(c, sh[tid]);
		if (i < n) agg[i] = r;
		__syncthreads();
		int last = (n - c0 < SCAN_B? n - c0 : SCAN_B) - 1;
		if (tid == last) carry = r;
		__syncthreads();
	}
}

__global__ void k_segmax_apply(const int32_t *__restrict__ seg, int32_t *out, const SegVal *__restrict__ agg, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int64_t t = i / SCAN_TILE;
	if (t == 0) return;
	SegVal p = agg[t - 1];
	if (p.seg == seg[i] && p.val > out[i]) out[i] = p.val;
}

// [emu-end segmax]
static void segmented_running_max(const int32_t *seg, const int32_t *val, int32_t *out, int64_t n, SegVal *agg_scratch, cudaStream_t s)
{
	if (n <= 0) return;
	int tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
	PGX_LAUNCH(k_segmax_tiles, tiles, SCAN_B, 0, s, seg, val, out, agg_scratch, n);
	if (tiles > 1) {
		PGX_LAUNCH(k_segmax_agg, 1, SCAN_B, 0, s, agg_scratch, tiles);
		PGX_LAUNCH(k_segmax_apply, grid_for(n, 256), 256, 0, s, seg, out, agg_scratch, n);
	}
}

} // namespace pgx
