/*
 * pgx.h -- C ABI of libpgx.so, the B200 (sm_100a) implementation of pangene's hot path.
 *
 * DROP-IN BOUNDARY.  The four pg_* functions below have exactly the signatures and the post-conditions of
 * the reference's entry points; a pangene host links libpgx instead of graph.o/vertex.o/branch.o/hit.o/
 * overlap.o and calls them unchanged (INTEGRATION.md shows the two-line Makefile change and the one-line
 * reader change).  Everything is plain C: pointers, int32/int64/double, the POD structs of pangene_abi.h.
 * Results are returned in libc-malloc'd memory that the reference's own destroy functions may free().
 * CUDA / NCCL failures print to stderr and abort(), matching the reference's fail-fast style.  There is no
 * CPU fallback: without a usable CUDA device every entry point aborts with a message.
 */
#ifndef PGX_H
#define PGX_H

#include "pangene_abi.h"
#include "pgx_walk.h"
#include "pgx_packed.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference: `extern int pg_verbose` (pangene.h:126, option.c:4) */
extern int pg_verbose;

/* Limits of this build: hit coordinates (cs, cm, ce) must lie in [0, 2^31) -- the reference keeps them in int64, libpgx packs them into
 * int32 columns and aborts with a message on upload otherwise (no contig of any sequenced genome reaches 2^31 bp); at most 2^31 / 11 hits and
 * 2^32 exons per rank.
 *
 * reference: pg_post_process(), pangene.h:133 / graph.c:7-32.  ADDITIONALLY runs, first, the per-genome
 * preprocessing block that the reference executes at the tail of pg_read_paf() (read.c:243-260) -- so `d` must
 * hold RAW parsed hits (as left by pgio_read_paf(), or by the reference reader with that block removed). On
 * return every genome's hit[] is a fresh malloc'd array sorted by (cid, cs) with all flags of that stage set
 * (unless lazy materialisation is on, see pgx_set_lazy). */
void pg_post_process(const pg_opt_t *opt, pg_data_t *d);

/* reference: pg_graph_init(), pangene.h:135 / graph.c:34-41 */
pg_graph_t *pg_graph_init(pg_data_t *d);

/* reference: pg_graph_gen(), pangene.h:136 / graph.c:280-322.  Continues on the device-resident state left by
 * pg_post_process(q->d). Fills q->seg/arc/g2s/idx (malloc'd) and every hit flag the writers print. */
void pg_graph_gen(const pg_opt_t *opt, pg_graph_t *q);

/* reference: pg_graph_destroy(), pangene.h:137 / graph.c:43-47 */
void pg_graph_destroy(pg_graph_t *g);

/* ---- engine control (no reference counterpart) ------------------------------------------------- */

/* Frees the device state pg_post_process(d) keeps for the pg_graph_gen() that follows; only needed when that call never comes
 * (e.g. `--bed=raw` stops after pg_post_process, main.c:122-123).  Safe to call at any time; a no-op for unknown pointers. */
void pgx_data_release(const pg_data_t *d);

int  pgx_device_count(void);
void pgx_set_device(int dev);                 /* default: current CUDA device */

/* 1: pg_post_process() leaves the hits on the device and only updates d->prot/d->gene; pg_graph_gen()
 * materialises everything once.  0 (default): strict reference post-conditions after each call.
 * 2: as 1, and pg_graph_gen() returns the graph only -- d->genome[].hit keeps the parsed records and the per-hit results stay on the
 * device, to be read through pgx_walk_tokens() (what the GFA writer needs) or pgx_fetch_hits() (everything); pgx_data_release(d) ends it. */
void pgx_set_lazy(int lazy);

/* The walk tokens of the final state, read off the device (lazy mode 2, after pg_graph_gen(); collective under pgx_comm_init: every rank gets
 * the tokens of its own genomes).  *tokens: malloc'd (the caller frees it), genome-major; inside a genome contig-major, inside a contig by
 * cm with ties in (cs, file) order -- the order pgio_write_gfa() prints.  genome_off: d->n_genome + 1 offsets into *tokens, caller-provided.
 * Returns the number of tokens. */
int64_t pgx_walk_tokens(const pg_data_t *d, pgx_walk_token_t **tokens, int64_t *genome_off);

/* pg_post_process() on DEVICE-READY input (pgx_packed.h): the hits of genome j are pk->hit[j][0 .. pk->n_hit[j]) -- the
 * twelve words pg_post_process() would pack out of every pg_hit_t -- in file order; d supplies everything else (n_genome, contigs, exon
 * lists, proteins, genes) and its genome[].hit arrays are not read.  Needs pgx_set_lazy(2): results come back as graph (pg_graph_gen) and
 * walk tokens (pgx_walk_tokens) only; d->prot / d->gene are updated as by pg_post_process(). */
void pgx_post_process_packed(const pg_opt_t *opt, pg_data_t *d, const pgx_packed_t *pk);

/* lazy mode 2: materialise the hits after all (what pg_graph_gen() does in modes 0 and 1); the device state stays until pgx_data_release(). */
void pgx_fetch_hits(pg_data_t *d);

/* Optional: create the CUDA context, the pinned staging ring and the host thread pool now instead of inside the first
 * pg_post_process().  May be called from a side thread while the caller parses its input (join before pg_post_process).
 * Returns 1 if a device was initialised, 0 otherwise (never aborts). */
int  pgx_warmup(void);

/* Multi-GPU, one process per GPU.  Genomes are partitioned across ranks (this rank's pg_data_t holds genomes
 * [genome_begin, genome_begin + d->n_genome) of n_genome_total); gene/protein ids must be global.  The 128-byte
 * id comes from pgx_comm_unique_id() on rank 0 and is broadcast by the launcher (torch.distributed, MPI, ...).
 * After pgx_comm_init() the pg_* calls are collective: every rank calls them in the same order. The graph
 * (seg/arc) comes back replicated on every rank; hits stay with their owner. */
int  pgx_comm_unique_id(void *out128);
int  pgx_comm_init(int rank, int nranks, const void *id128, int32_t n_genome_total, int32_t genome_begin);
void pgx_comm_destroy(void);
/* The same sharded path over the LOOPBACK transport (pgx_comm.cuh): collectives staged through a POSIX shared-memory segment and reduced on
 * the host, so that the ranks may be processes sharing ONE device.  For testing the partition / merge logic on a single-GPU box only. */
int  pgx_comm_init_loopback(int rank, int nranks, const char *segment_name, int32_t n_genome_total, int32_t genome_begin);

/* ---- staged API: the same path with the host<->device boundary made explicit (bench.py, tests) -- */
typedef struct pgx_session_s pgx_session_t;

pgx_session_t *pgx_upload(const pg_data_t *d);                       /* host AoS -> HBM */
void pgx_run(pgx_session_t *s, const pg_opt_t *opt, int stage);      /* stage 1: load-time block + pg_post_process; 2: pg_graph_gen; 3: both */
void pgx_download(pgx_session_t *s, pg_data_t *d, pg_graph_t *q);    /* HBM -> malloc'd host structs; q may be NULL after stage 1 */
void pgx_session_free(pgx_session_t *s);

typedef struct {
	uint64_t launches;            /* kernels launched by this library since the last reset */
	double   sweep_ms;            /* device time of the FIRST (load-time) overlap sweep of the last run, CUDA events */
	double   sweep_bytes;         /* its algorithmic bytes: sum_unfiltered(52 + 8 n_exon) + 4 n_filtered; counted by an
	                               * extra kernel + read-back, only in runs made while pgx_set_sweep_stat(1) is in force (else kept) */
	double   run_ms[3];           /* device time of stage 1, stage 2, total of the last pgx_run */
	int64_t  n_hit, n_active;     /* hits uploaded / hits that survived the load-time filters */
	int64_t  tie[4];              /* tie-order detector, same meaning as oracle pgo_tie_report() */
	int32_t  n_seg, n_arc;
	int64_t  h2d_bytes, d2h_bytes;
	int32_t  fast_rounds;         /* pg_gen_arc calls of the last runs that took the sync-free path (pgx_round.cuh) */
	int32_t  redo;                /* runs redone on the careful path (table overflow, or a round that needs the reference's tie order) */
} pgx_stats_t;
void pgx_get_stats(pgx_stats_t *out);
void pgx_reset_stats(void);
void pgx_set_sweep_stat(int on);              /* count sweep_bytes in the following runs (default off: it costs a pass over the hits) */
void pgx_trim(void);                          /* give cached device memory back to the driver */

/* test hooks for the device primitives (tests/test_gpu_primitives.py) */
void pgx_test_radix_sort(uint64_t *keys, uint32_t *vals, int64_t n, int n_bits);
void pgx_test_exclusive_sum(const uint32_t *in, uint32_t *out, int64_t n);
void pgx_test_segmented_max(const int32_t *seg, const int32_t *val, int32_t *out, int64_t n);
void pgx_test_flag_sort_replay(pg128_t *a, int64_t n);
int  pgx_test_choose_representatives(const unsigned long long *z, const int32_t *prot_gid, int32_t n_prot, int32_t n_gene, int32_t force_replay, int32_t *gene_rep_pid);

#ifdef __cplusplus
}
#endif
#endif
