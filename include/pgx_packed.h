/*
 * pgx_packed.h -- device-ready hit records ("emit packed columns directly so the H2D copy needs no AoS pass").
 *
 * pg_post_process() of libpgx packs the twelve words the device needs out of every 88-byte pg_hit_t (pangene.h:61-72) on its way into the
 * pinned staging buffers.  A reader that knows it feeds libpgx can emit those words itself: libpgio's pgio_read_pafs_packed() parses PAF text
 * straight into pgx_packed_hit_t arrays (the pg_hit_t records never exist), and pgx_post_process_packed() uploads them as they are.
 * Only the GFA route (graph + walk tokens, pgx_walk.h) is available on packed input: there is no pg_hit_t record to hand results back in.
 */
#ifndef PGX_PACKED_H
#define PGX_PACKED_H
#include <stdint.h>
typedef struct {          /* 48 bytes; field meaning as in pg_hit_t (pangene.h:61-72), coordinates narrowed to int32 (checked by the producer) */
	int32_t pid, cid, rank, score_ori, score_adj, n_exon, off_exon, rev, cs, cm, ce, lof;
} pgx_packed_hit_t;
typedef struct {          /* the hits of a pg_data_t whose genome[j].hit arrays are empty; exons, contigs and the name tables stay in the pg_data_t */
	int32_t n_genome, m_genome;
	int32_t *n_hit, *m_hit;          /* per genome */
	pgx_packed_hit_t **hit;          /* per genome, malloc'd, file order */
} pgx_packed_t;
#endif
