// pgx_graphtypes.cuh -- device-side records shared by the graph kernels (pgx_kernels.cuh) and the sync-free round body (pgx_round.cuh).
// Plain structs and two macros; the text between the markers is also compiled for the host by tests/emu/build.py.
#pragma once
#include <cstdint>

namespace pgx {
// [emu-begin graphtypes]
// Logical sizes that change from round to round live ON THE DEVICE: kernels read them through this struct and are launched
// with capacity-sized (or fixed, grid-stride) grids, so a whole arc/branch round is enqueued without a host round-trip.
struct DevCounts {
	uint32_t m_sv, n_part, n_arc, n_seg, n_nl, m_cs;
	uint32_t n_ovf, n_l2, n_gpart, fallback;      // pgx_round.cuh: overflow records, their (arc, genome) groups, merged arcs, sticky "redo this run carefully"
	uint32_t pad[6];
};
#define PGX_GS(i, n) for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x, gs_n_ = (int32_t)(n), gs_s_ = gridDim.x * blockDim.x; i < gs_n_; i += gs_s_)

// partial arc = sums over the LOCAL genomes of one key (graph.c:128-145 per genome, then the additive part of 153-163)
struct ArcPartial { uint64_t x; int64_t sum_dist, sum_s1, sum_s2; int32_t n_genome, tot_cnt; };
struct ArcDev { uint64_t x; int32_t n_genome, tot_cnt, avg_dist, s1, s2; uint32_t bits; };   // == pg_arc_t
struct SegDev { int32_t gid, n_dom, n_sub, n_genome, tot_cnt; uint32_t del; int32_t n_dist_loci[2]; };   // == pg_seg_t
struct GenePos { int32_t cid, ord, cm, pad; };                  // cid < 0: absent
struct BranchOpt { int32_t local_dist, local_count, frag_mode; double branch_diff, branch_diff_dist, branch_diff_cut; };

// graph.c:153-175, the three divisions: integer sums first (exact for any split over genomes / ranks), IEEE double division last
__device__ __forceinline__ ArcDev arc_finalize(const ArcPartial &a)
{
	ArcDev o;
	o.x = a.x; o.n_genome = a.n_genome; o.tot_cnt = a.tot_cnt; o.bits = 0;
	o.avg_dist = (int32_t)(long long)((double)a.sum_dist / a.tot_cnt + .499);          // graph.c:170
	o.s1 = (int32_t)((double)a.sum_s1 / a.n_genome + .499);                            // graph.c:171
	o.s2 = (int32_t)((double)a.sum_s2 / a.n_genome + .499);                            // graph.c:172
	return o;
}
// [emu-end graphtypes]
} // namespace pgx
