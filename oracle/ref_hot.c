/*
 * ref_hot.c -- TEST INFRASTRUCTURE ONLY.  A thin driver (ours) that is linked INTO oracle/_ref/libpgref.so
 * next to the unmodified reference objects.  It runs the reference's own hot path on a pg_data_t that
 * holds raw parsed hits (as produced by libpgio, i.e. WITHOUT the per-genome block of read.c:243-260):
 *
 *     for each genome: the block of read.c:243-260, call for call      (hoist verified output-neutral,
 *     pg_post_process()  graph.c:7                                       SURVEY.md §3.2)
 *     pg_graph_init()    graph.c:34
 *     pg_graph_gen()     graph.c:280
 *
 * and prints results through the reference's own writers (format.c) into a file.  Only the reference's
 * headers are included (from $(REF) via -I); no reference source is copied.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include "pgpriv.h"

/* t[0] = per-genome block, t[1] = pg_post_process, t[2] = pg_graph_init + pg_graph_gen (seconds) */
pg_graph_t *pgref_hot_path(const pg_opt_t *opt, pg_data_t *d, double *t, int stop_after_post)
{
	int32_t j, i, check_strand = !!(opt->flag & PG_F_CHECK_STRAND);
	double t0 = pg_realtime(), t1, t2, t3;
	pg_graph_t *q = 0;
	for (j = 0; j < d->n_genome; ++j) {
		pg_genome_t *g = &d->genome[j];
		pg_flag_pseudo(d->prot, g);
		for (i = 0; i < g->n_hit; ++i)
			if (g->hit[i].pseudo) g->hit[i].flt = 1;
		pg_hit_sort(g, 0);
		pg_shadow(opt, d, j, 1, check_strand);
		for (i = 0; i < g->n_hit; ++i) {
			pg_hit_t *a = &g->hit[i];
			a->pid_dom0 = a->pid_dom;
			a->pid_dom = -1, a->shadow = 0;
		}
		pg_flt_ov_isoform(opt, d, j, check_strand);
		pg_flt_chain_shadow(d->prot, d->n_prot, g);
		pg_flt_subopt_isoform(d->prot, d->n_gene, g);
	}
	t1 = pg_realtime();
	pg_post_process(opt, d);
	t2 = pg_realtime();
	if (!stop_after_post) {
		q = pg_graph_init(d);
		pg_graph_gen(opt, q);
	}
	t3 = pg_realtime();
	if (t) t[0] = t1 - t0, t[1] = t2 - t1, t[2] = t3 - t2;
	return q;
}

void pgref_graph_destroy(pg_graph_t *q) { if (q) pg_graph_destroy(q); }

void pgref_set_verbose(int v) { pg_verbose = v; }

/* mode 0: GFA S+L+W; 1: GFA S+L only; 2: BED all hits (raw/flag); 3: BED walk.  The reference writers
 * print to stdout, so fd 1 is pointed at `path` for the duration of the call. */
int pgref_write(pg_graph_t *q, pg_data_t *d, int mode, const char *path)
{
	int fd, saved;
	fflush(stdout);
	fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
	if (fd < 0) return -1;
	saved = dup(1);
	dup2(fd, 1);
	close(fd);
	if (mode == 0 || mode == 1) {
		pg_write_graph(q);
		if (mode == 0) pg_write_walk(q);
	} else pg_write_bed(d, mode == 3);
	fflush(stdout);
	dup2(saved, 1);
	close(saved);
	return 0;
}
