#include "pgio.h"
extern "C" {
#include "pg_oracle.h"
}
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
int main(int argc, char **argv)
{
	pg_opt_t o; pgio_opt_init(&o);
	for (int cfg : {2, 5}) {
		pgio_synth_t p; pgio_synth_default(&p, cfg);
		p.n_genome = cfg == 2? 4 : 30; p.n_genome_total = p.n_genome; p.n_gene = cfg == 2? 700 : 400; p.iso_per_gene = cfg == 2? 4 : 1; p.hits_per_genome = cfg == 2? 6000 : 350;
		std::vector<std::string> files; std::vector<const char*> fp;
		for (int j = 0; j < p.n_genome; ++j) { files.push_back("/tmp/asan/g" + std::to_string(cfg) + "_" + std::to_string(j) + ".paf"); pgio_synth_write_paf(&o, &p, j, files.back().c_str()); }
		for (auto &f : files) fp.push_back(f.c_str());
		fp.push_back("/tmp/asan/missing.paf");
		pg_data_t *d = pgio_data_init();
		int n = pgio_read_pafs(&o, d, fp.data(), (int)fp.size(), 4);
		pg_data_t *d1 = pgio_data_init();
		for (auto f : fp) pgio_read_paf(&o, d1, f);
		printf("cfg %d: %d files, %ld hits (sequential %ld)\n", cfg, n, (long)pgio_total_hits(d), (long)pgio_total_hits(d1));
		pgo_post_process(&o, d);
		pg_graph_t *g = pgo_graph_init(d);
		pgo_graph_gen(&o, g);
		pgio_write_gfa(g, "/tmp/asan/out.gfa", 3);
		pgio_write_bed(d, 0, "/tmp/asan/out.bed"); pgio_write_bed(d, 1, "/tmp/asan/out.walk.bed");
		int rows = pgio_gfa2matrix("/tmp/asan/out.gfa", "/tmp/asan/out.tsv", 0);
		int64_t tie[4]; pgo_tie_report(tie);
		printf("  %d segs %d arcs, matrix rows %d, ties %ld %ld %ld %ld\n", g->n_seg, g->n_arc, rows, (long)tie[0], (long)tie[1], (long)tie[2], (long)tie[3]);
		{   // the packed reader (pgx_packed.h) and the token-fed writers (pgx_walk.h): tokens built from the oracle's hits
			pg_data_t *d2 = pgio_data_init();
			pgx_packed_t pk = {};
			int n2 = pgio_read_pafs_packed(&o, d2, &pk, fp.data(), (int)fp.size(), 3);
			long same = 0;
			for (int j = 0; j < d2->n_genome; ++j) for (int i = 0; i < pk.n_hit[j]; ++i) same += pk.hit[j][i].cs >= 0;
			std::vector<pgx_walk_token_t> tok; std::vector<int64_t> off(1, 0);
			for (int j = 0; j < d->n_genome; ++j) {
				std::vector<int> ord;
				for (int i = 0; i < d->genome[j].n_hit; ++i) if (!d->genome[j].hit[i].flt) ord.push_back(i);
				std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { const pg_hit_t &x = d->genome[j].hit[a], &y = d->genome[j].hit[b]; return x.cid != y.cid? x.cid < y.cid : x.cm < y.cm; });
				for (int i : ord) { const pg_hit_t &h = d->genome[j].hit[i]; tok.push_back(pgx_walk_token_t{h.cid, (uint32_t)d->prot[h.pid].gid << 1 | h.rev, h.lof}); }
				off.push_back((int64_t)tok.size());
			}
			pgio_write_gfa_tokens(g, tok.data(), off.data(), "/tmp/asan/out_tok.gfa", 3);
			int rows2 = pgio_matrix_from_tokens(g, tok.data(), off.data(), "/tmp/asan/out_tok.tsv", 1);
			printf("  packed reader: %d files, %ld records (%ld checked); token writers: %zu tokens, matrix rows %d, GFA %s\n", n2, (long)pgio_packed_total_hits(&pk), same, tok.size(), rows2,
				system("cmp -s /tmp/asan/out.gfa /tmp/asan/out_tok.gfa") == 0? "identical" : "DIFFERENT");
			pgio_packed_destroy(&pk); pgio_data_destroy(d2);
		}
		pgo_graph_destroy(g); pgio_data_destroy(d); pgio_data_destroy(d1);
	}
	void *s = pgio_name_set("a,b,c"); pgio_name_set_destroy(s);
	return 0;
}
