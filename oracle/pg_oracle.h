/*
 * pg_oracle.h -- TEST INFRASTRUCTURE ONLY: the CPU restatement of the hot path (see pg_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load libpgoracle.so.
 */
#ifndef PG_ORACLE_H
#define PG_ORACLE_H
#include "pangene_abi.h"
#ifdef __cplusplus
extern "C" {
#endif

/* same contracts as the reference entry points (pangene.h:133-137); pgo_post_process additionally runs
 * the per-genome block of read.c:243-260 first, because our reader does not (include/pgio.h). */
void pgo_post_process(const pg_opt_t *opt, pg_data_t *d);
pg_graph_t *pgo_graph_init(pg_data_t *d);
void pgo_graph_gen(const pg_opt_t *opt, pg_graph_t *q);
void pgo_graph_destroy(pg_graph_t *q);

/* exact replay of the permutation produced by the reference's in-place MSD byte radix sort
 * (ksort.h:52-87) on 16-byte records keyed by .x */
void pgo_radix_replay128(pg128_t *a, int64_t n);
void pgo_test_chain_order(const pg_genome_t *g_in_file_order, int32_t n_pairs, int32_t by_cm, int32_t *order_out);

/* tie-order detector (DESIGN.md "tie order"): out[0] = survivor pairs with equal (contig, cm),
 * out[1] = unfiltered pairs with equal (contig, cs), out[2] = duplicate-dominator ties in the load-time
 * sweep, out[3] = equal-score first-wins ties in the sub-optimal isoform filter. All zero means the result
 * does not depend on the reference's history-dependent tie order. Counters reset by pgo_post_process. */
void pgo_tie_report(int64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
