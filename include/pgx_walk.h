/*
 * pgx_walk.h -- the walk token: one printed hit of a W line (reference: pg_write_walk(), format.c:183-225 prints the unfiltered hits of a
 * contig in CDS-mid-point order, each as strand + gene name, followed by the lf:B:i list).  libpgx reads these 12 bytes per printed hit off
 * its device-resident final state (pgx_walk_tokens(), pgx.h); libpgio formats W lines and the gfa2matrix table from them (pgio.h) -- the
 * 88-byte pg_hit_t records of ALL hits never have to come back for a GFA.
 */
#ifndef PGX_WALK_H
#define PGX_WALK_H
#include <stdint.h>
typedef struct {
	int32_t  cid;             /* contig of the hit, genome-local (pg_hit_t.cid) */
	uint32_t gene_strand;     /* prot[pid].gid << 1 | rev */
	int32_t  lof;             /* pg_hit_t.lof */
} pgx_walk_token_t;
#endif
