// pgx_sort.cuh -- stable LSD radix sort of (uint64 key, uint32 value) pairs, 8 bits per pass, hand-written.
//
// Used for: hits by (genome:contig:cs) once, active hits by (genome:contig:cm) once, arc records by
// (v:w:genome) once per arc round.  STABILITY IS PART OF THE CONTRACT: ties keep their incoming order
// (DESIGN.md "tie order").
//
// Per pass: k_rs_hist (per-tile digit histogram) -> exclusive scan over [digit][tile] -> k_rs_scatter
// (re-reads the tile, ranks every key among equal digits in tile order, scatters).  A tile is 8 warps x 512
// consecutive keys; each warp ranks its chunk with __match_any_sync + a private 256-counter row in shared
// memory, so the rank of a key is (#equal digits in earlier warps' chunks) + (#equal digits earlier in its own
// chunk) -- i.e. exactly its position among equal digits in input order.
#pragma once
#include "pgx_util.cuh"
#include "pgx_scan.cuh"

namespace pgx {

// [emu-begin sort]  (tests/emu/build.py compiles the two kernels for the host: tests/test_prim_emu.py)
constexpr int RS_WARPS = 8, RS_BLOCK = RS_WARPS * 32, RS_ITEMS = 8, RS_TILE = RS_BLOCK * RS_ITEMS;   // 2048 keys per tile: 34 KB of shared memory, six CTAs per SM (4096 keys: 59 KB, three CTAs, 1.37 TB/s)

__device__ __forceinline__ uint32_t rs_digit(uint64_t k, int shift) { return (uint32_t)(k >> shift) & 0xffu; }

__global__ void __launch_bounds__(RS_BLOCK) k_rs_hist(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t *hist, int n_tiles)
{
	__shared__ uint32_t h[256];
	const int tid = threadIdx.x;
	h[tid] = 0;
	__syncthreads();
	const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
	for (int r = 0; r < RS_ITEMS; ++r) {
		int64_t i = base + (int64_t)r * RS_BLOCK + tid;
		if (i < n) atomicAdd(&h[rs_digit(keys[i], shift)], 1u);
	}
	__syncthreads();
	hist[(size_t)tid * n_tiles + blockIdx.x] = h[tid];
}

// vals_in == nullptr means "the value of element i is i" (first pass over fresh keys).
// After ranking, the tile is first permuted INSIDE shared memory (digit-major, stable), then written out: consecutive threads
// store consecutive elements of a digit run, so every 32-byte sector written is fully used (a direct scatter of 8-byte keys to
// 256 destinations touches one sector per key: ~4x write amplification, measured 96 us per pass on 3.6 M records).
constexpr size_t RS_SMEM = (size_t)RS_TILE * 8 + (size_t)RS_TILE * 4 + (size_t)RS_WARPS * 256 * 4 + 256 * 4 * 2;
__global__ void __launch_bounds__(RS_BLOCK, 5) k_rs_scatter(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals_in, int64_t n, int shift,
	const uint32_t *__restrict__ hist_scan, int n_tiles, uint64_t *keys_out, uint32_t *vals_out)
{
	extern __shared__ __align__(16) unsigned char rs_smem[];
	uint64_t *skey = (uint64_t*)rs_smem;                                   // [RS_TILE] tile, digit-major
	uint32_t *sval = (uint32_t*)(rs_smem + (size_t)RS_TILE * 8);           // [RS_TILE]
	uint32_t (*wcnt)[256] = (uint32_t (*)[256])(rs_smem + (size_t)RS_TILE * 12);   // [RS_WARPS][256]
	uint32_t *blk_off = (uint32_t*)(rs_smem + (size_t)RS_TILE * 12 + (size_t)RS_WARPS * 1024);   // [256] exclusive digit offsets inside the tile
	uint32_t *gbase = blk_off + 256;                                       // [256] global base of (digit, tile)
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int k = tid; k < RS_WARPS * 256; k += RS_BLOCK) (&wcnt[0][0])[k] = 0;
	__syncthreads();
	const int64_t tbase = (int64_t)blockIdx.x * RS_TILE, wbase = tbase + (int64_t)warp * (32 * RS_ITEMS);
	uint64_t key[RS_ITEMS];
	uint32_t rank[RS_ITEMS];
	const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
	for (int r = 0; r < RS_ITEMS; ++r) {
		int64_t i = wbase + r * 32 + lane;
		bool ok = i < n;
		key[r] = ok? keys[i] : ~0ull;
		uint32_t d = ok? rs_digit(key[r], shift) : 256u;    // 256 = "not an element": its own match group
		uint32_t peers = __match_any_sync(0xffffffffu, d);
		uint32_t before = ok? wcnt[warp][d] : 0u;
		__syncwarp();
		if (ok && (peers & lt) == 0) wcnt[warp][d] = before + __popc(peers);   // lowest lane of the group bumps the row
		__syncwarp();
		rank[r] = before + __popc(peers & lt);
	}
	__syncthreads();
	// digit `tid`: per-warp counts -> per-warp exclusive bases inside the digit; digit total -> blk_off (scanned below)
	{
		uint32_t run = 0;
#pragma unroll
		for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = wcnt[w][tid]; wcnt[w][tid] = run; run += c; }
		blk_off[tid] = run;
		gbase[tid] = hist_scan[(size_t)tid * n_tiles + blockIdx.x];
	}
	__syncthreads();
	if (warp == 0) {   // exclusive scan of the 256 digit totals: 8 per lane + a warp scan
		uint32_t v[8], sum = 0;
#pragma unroll
		for (int k = 0; k < 8; ++k) { v[k] = blk_off[lane * 8 + k]; sum += v[k]; }
		uint32_t inc = sum;
		for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
		uint32_t run = inc - sum;
#pragma unroll
		for (int k = 0; k < 8; ++k) { blk_off[lane * 8 + k] = run; run += v[k]; }
	}
	__syncthreads();
#pragma unroll
	for (int r = 0; r < RS_ITEMS; ++r) {
		int64_t i = wbase + r * 32 + lane;
		if (i < n) {
			uint32_t d = rs_digit(key[r], shift);
			uint32_t lp = blk_off[d] + wcnt[warp][d] + rank[r];
			skey[lp] = key[r];
			sval[lp] = vals_in? vals_in[i] : (uint32_t)i;
		}
	}
	__syncthreads();
	const int n_tile = (int)((n - tbase) < RS_TILE? (n - tbase) : RS_TILE);
	for (int j = tid; j < n_tile; j += RS_BLOCK) {
		uint64_t k = skey[j];
		uint32_t d = rs_digit(k, shift);
		uint32_t pos = gbase[d] + ((uint32_t)j - blk_off[d]);
		keys_out[pos] = k;
		vals_out[pos] = sval[j];
	}
}

// [emu-end sort]
struct RadixSortBuffers {          // caller-provided scratch (no allocation inside the sort)
	uint64_t *keys_alt; uint32_t *vals_alt; uint32_t *hist; uint32_t *scan_agg;
	static size_t hist_elems(int64_t n) { int64_t t = (n + RS_TILE - 1) / RS_TILE; if (t < 1) t = 1; return (size_t)(256 * t + 1); }
};

// Sorts by key bits [0, n_bits). Result lands in *keys_res / *vals_res (either the primary or the alternate
// buffers). vals may be "implicit iota" on entry (iota_vals = true: vals need not be initialised).
static void radix_sort_pairs(uint64_t *keys, uint32_t *vals, int64_t n, int n_bits, bool iota_vals, const RadixSortBuffers &b,
	uint64_t **keys_res, uint32_t **vals_res, cudaStream_t s)
{
	static int attr_dev = -1;                  // per device
	int cur_dev = 0; PGX_CHECK(cudaGetDevice(&cur_dev));
	if (attr_dev != cur_dev) {
		PGX_CHECK(cudaFuncSetAttribute(k_rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM));
		PGX_CHECK(cudaFuncSetAttribute(k_rs_scatter, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));   // five 35 KB CTAs per SM (ncu: the default carve-out left room for three)
		attr_dev = cur_dev;
	}
	uint64_t *kin = keys, *kout = b.keys_alt;
	uint32_t *vin = vals, *vout = b.vals_alt;
	if (n > 0) {
		int n_tiles = (int)((n + RS_TILE - 1) / RS_TILE);
		int64_t hn = 256ll * n_tiles;
		bool first = true;
		if (n_bits < 1) n_bits = 1;
		for (int shift = 0; shift < n_bits; shift += 8) {
			PGX_LAUNCH(k_rs_hist, n_tiles, RS_BLOCK, 0, s, kin, n, shift, b.hist, n_tiles);
			exclusive_sum_u32(b.hist, b.hist, hn, b.scan_agg, s);
			PGX_LAUNCH(k_rs_scatter, n_tiles, RS_BLOCK, RS_SMEM, s, kin, (first && iota_vals)? (const uint32_t*)nullptr : vin, n, shift, b.hist, n_tiles, kout, vout);
			uint64_t *tk = kin; kin = kout; kout = tk;
			uint32_t *tv = vin; vin = vout; vout = tv;
			first = false;
		}
	}
	*keys_res = kin; *vals_res = vin;
}

static inline int bits_for(uint64_t max_value) { int b = 0; while (max_value) { ++b; max_value >>= 1; } return b > 0? b : 1; }

} // namespace pgx
