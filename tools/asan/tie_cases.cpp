#include "pgio.h"
extern "C" {
#include "pg_oracle.h"
}
#include <cstdio>
struct Case { int cfg, n_genome, n_gene, iso, hits, n_ctg; unsigned long long seed; int strand, T, r, a; double p, f; int ld, lc; };
int main()
{
	Case cs[] = {
		{2, 8, 1904, 1, 6585, 1, 1047424420575ull, 0, 15, 3, 1, 0.01, 0.5, -1, -1},
		{5, 79, 852, 1, 751, 24, 204228014272ull, 1, 7, 1, 1, 0.05, 0.5, 1, 2},
		{5, 69, 696, 1, 446, 2, 155451220174ull, 1, 7, 1, 1, 0.05, 0.5, -1, -1},
		{5, 75, 742, 1, 628, 1, 770633651119ull, 1, 15, 3, 2, 0.05, 0.2, 1, 1},
	};
	for (const Case &c : cs) {
		pg_opt_t o; pgio_opt_init(&o);
		if (c.strand) o.flag |= PG_F_CHECK_STRAND;
		o.n_branch_flt = c.T; o.max_dist_loci = c.r; o.min_arc_cnt = c.a; o.min_vertex_ratio = c.p; o.min_ov_ratio = c.f;
		if (c.ld >= 0) { o.local_dist = c.ld; o.local_count = c.lc; }
		pgio_synth_t p; pgio_synth_default(&p, c.cfg);
		p.n_genome = c.n_genome; p.n_genome_total = c.n_genome; p.n_gene = c.n_gene; p.iso_per_gene = c.iso; p.hits_per_genome = c.hits; p.n_ctg = c.n_ctg; p.seed = c.seed;
		pg_data_t *d = pgio_synth(&o, &p);
		pgo_post_process(&o, d);
		pg_graph_t *g = pgo_graph_init(d);
		pgo_graph_gen(&o, g);
		int64_t tie[4]; pgo_tie_report(tie);
		pgio_write_gfa(g, "/tmp/asan/out2.gfa", 3);
		printf("%ld hits, %d segs %d arcs, ties %ld %ld %ld %ld\n", (long)pgio_total_hits(d), g->n_seg, g->n_arc, (long)tie[0], (long)tie[1], (long)tie[2], (long)tie[3]);
		pgo_graph_destroy(g); pgio_data_destroy(d);
	}
	return 0;
}
