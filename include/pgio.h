/*
 * pgio.h -- host-side I/O around the hot path (C ABI, no CUDA): PAF reader, GFA/BED writers, option
 * defaults and the deterministic synthetic-PAF generator used by tests and bench.py.
 *
 * These are the SURVEY.md §8(f) "next" rows (callers/data formats either side of the path), written
 * from scratch; each entry cites the reference interface it stands in for.  The one deliberate
 * difference from the reference reader: pgio_read_paf() stops after parsing -- the per-genome
 * preprocessing block that the reference runs at the tail of pg_read_paf() (read.c:243-260) is
 * hot-path work and is executed by pg_post_process() of libpgx instead (SURVEY.md §3.2 [verified]
 * output-neutral).
 */
#ifndef PGIO_H
#define PGIO_H

#include "pangene_abi.h"
#include "pgx_walk.h"
#include "pgx_packed.h"

#ifdef __cplusplus
extern "C" {
#endif

extern int pgio_verbose;

/* option.c:6-26 -- reference defaults */
void pgio_opt_init(pg_opt_t *opt);

/* read.c:10-32 */
pg_data_t *pgio_data_init(void);
void pgio_data_destroy(pg_data_t *d);

/* read.c:107-241 (everything before the postprocessing block). gzip or plain text, "-" = stdin.
 * Returns 0, or -1 if the file cannot be opened (then no genome is added), like the reference. */
int32_t pgio_read_paf(const pg_opt_t *opt, pg_data_t *d, const char *fn);

/* All files of a command line at once: each file is parsed by a worker thread with file-local ids, then the files are merged in
 * argument order -- gene/protein ids, names, ranks and every hit come out exactly as from calling pgio_read_paf() file by file
 * (tests/test_io.py).  n_threads 0 = hardware concurrency.  Returns the number of files opened. */
int32_t pgio_read_pafs(const pg_opt_t *opt, pg_data_t *d, const char *const *fns, int32_t n_files, int32_t n_threads);

/* The same parse, emitting DEVICE-READY records (pgx_packed.h): every kept alignment becomes a 48-byte pgx_packed_hit_t in pk->hit[genome]
 * instead of an 88-byte pg_hit_t -- same ids, ranks, scores, exon lists, cm and lof as pgio_read_pafs() computes (tests/test_io.py packs the
 * one and compares with the other).  d gets everything else (names, contigs, exons, proteins, genes); d->genome[j].hit stays NULL / n_hit 0.
 * pk: zero-initialised by the caller, released with pgio_packed_destroy().  A coordinate >= 2^31 aborts with a message. */
int32_t pgio_read_pafs_packed(const pg_opt_t *opt, pg_data_t *d, pgx_packed_t *pk, const char *const *fns, int32_t n_files, int32_t n_threads);
void pgio_packed_destroy(pgx_packed_t *pk);
int64_t pgio_packed_total_hits(const pgx_packed_t *pk);

/* read.c:305-318: name set for -X/-I/-P from "a,b,c" or "@file" */
void *pgio_name_set(const char *spec);
void pgio_name_set_destroy(void *set);

/* format.c:153 + format.c:183 + format.c:113. path "-" = stdout. flags: bit0 = write S/L lines,
 * bit1 = write W lines.  pgio_write_gfa re-sorts nothing in place (the reference's pg_write_walk
 * sorts every genome by cm and back; we sort an index instead). */
int32_t pgio_write_gfa(const pg_graph_t *g, const char *path, int32_t flags);
int32_t pgio_write_bed(const pg_data_t *d, int32_t is_walk, const char *path);

/* The same GFA from walk tokens (pgx_walk_tokens(), pgx.h) instead of materialised hits: S/L lines from the graph (bit0), W lines (bit1,
 * format.c:183-225) from tok[genome_off[j] .. genome_off[j+1]) of genome j -- already in printing order.  d supplies names only
 * (contigs, genes, genome labels); d->genome[].hit is not read.  Byte-identical to pgio_write_gfa() on the materialised data. */
int32_t pgio_write_gfa_tokens(const pg_graph_t *g, const pgx_walk_token_t *tok, const int64_t *genome_off, const char *path, int32_t flags);

/* pangene.js:1168-1246 without the GFA in between: the gene x assembly table straight from the graph (rows: segments, in S-line order) and
 * the walk tokens (columns: sample#hap of the contigs that have a W line, in W-line order; a token counts for the segment of its gene).
 * Same TSV as pgio_gfa2matrix() prints from the GFA text of the same run.  genome_begin: index of the first local genome (names unlabeled
 * genomes, format.c:207-208).  Returns #rows or -1. */
int32_t pgio_matrix_from_tokens(const pg_graph_t *g, const pgx_walk_token_t *tok, const int64_t *genome_off, const char *out_path, int32_t copy_number);

/* pangene.js:1168-1246 (gfa2matrix, without -d): gene x assembly presence/absence (or counts with
 * copy_number != 0) computed from a GFA file; written as TSV to out_path. Returns #rows or -1. */
int32_t pgio_gfa2matrix(const char *gfa_path, const char *out_path, int32_t copy_number);

/* ---- deterministic synthetic input ------------------------------------------- */
typedef struct {
	int32_t shape;            /* 0 = eukaryote (configs 2-4), 1 = bacterial (config 5) */
	int32_t n_genome;         /* genomes to generate in THIS call: global indices [genome_begin, genome_begin+n_genome) */
	int32_t genome_begin;     /* first global genome index (rank sharding); ids/names are global */
	int32_t n_genome_total;   /* population size (only used for naming/frequencies) */
	int32_t n_gene;           /* gene (cluster) count */
	int32_t iso_per_gene;     /* isoforms per gene: 1 or 4 */
	int32_t n_ctg;            /* contigs per genome (eukaryote) / max contigs (bacterial) */
	int32_t hits_per_genome;  /* target hits per genome (approximate) */
	uint64_t seed;
	int32_t n_threads;        /* 0 = hardware concurrency */
} pgio_synth_t;

void pgio_synth_default(pgio_synth_t *p, int32_t config);   /* config = 2,3,5 (BASELINE.json configs) */
/* Build pg_data_t in memory, exactly what pgio_read_paf() would produce from the PAF text of the same
 * records (tests/test_io.py checks the round trip). opt supplies the reader-side filters/score coef. */
pg_data_t *pgio_synth(const pg_opt_t *opt, const pgio_synth_t *p);
/* Write genome j of a synthetic run as miniprot-dialect PAF text (plain, not gz). */
int32_t pgio_synth_write_paf(const pg_opt_t *opt, const pgio_synth_t *p, int32_t j_global, const char *path);

/* deep copy (hits, exons, ctg tables share name pointers with the source dictionaries) */
pg_data_t *pgio_data_clone(const pg_data_t *d);
void pgio_data_clone_destroy(pg_data_t *d);

int64_t pgio_total_hits(const pg_data_t *d);

#ifdef __cplusplus
}
#endif
#endif
