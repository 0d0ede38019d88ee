// main.cpp -- the `pangene` command line over libpgio (reader/writers) and libpgx (the CUDA hot path).
// Same options and output as the reference CLI (main.c:12-152); the call sequence in main() is the
// reference's own: read every PAF, pg_post_process, pg_graph_init, pg_graph_gen, write.
// Attribution: the usage text and the option letters / their mapping onto pg_opt_t fields are the command-line CONTRACT of
// lh3/pangene (main.c:12-113, MIT licence, (c) Heng Li) and are reproduced as such; everything around them is this project's.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <getopt.h>
#include <thread>
#include <vector>
#include <sys/time.h>
#include <sys/resource.h>
#include "pgio.h"
#include "pgx.h"

#define PGX_VERSION "1.1-r231+b200"

static double now() { struct timeval tp; gettimeofday(&tp, 0); return tp.tv_sec + tp.tv_usec * 1e-6; }

static int usage(FILE *fp, const pg_opt_t *o)
{
	fprintf(fp, "Usage: pangene [options] <in.paf> [...]\n");
	fprintf(fp, "Options:\n  Input preprocessing:\n");
	fprintf(fp, "    -d CHAR       gene-protein delimiter [%c]\n", o->gene_delim);
	fprintf(fp, "    -X STR/@FILE  exclude genes in STR list or in @FILE []\n");
	fprintf(fp, "    -I STR/@FILE  include genes in the output graph []\n");
	fprintf(fp, "    -P STR/@FILE  prioritize genes in the output graph []\n");
	fprintf(fp, "    -e FLOAT      drop an alignment if its identity <FLOAT [%g]\n", o->min_prot_iden);
	fprintf(fp, "    -l FLOAT      drop an alignment if <FLOAT fraction of the protein aligned [%g]\n", o->min_prot_ratio);
	fprintf(fp, "    -m FLOAT      score adjustment coefficient [%g]\n", o->score_adj_coef);
	fprintf(fp, "  Graph construction:\n");
	fprintf(fp, "    -f FLOAT      min overlap fraction [%g]\n", o->min_ov_ratio);
	fprintf(fp, "    -J            don't filter pseudogenes across samples\n");
	fprintf(fp, "    -E            ignore genes that are single-exon in all genomes\n");
	fprintf(fp, "    -p FLOAT      gene considered if dominant in FLOAT fraction of genes [%g]\n", o->min_vertex_ratio);
	fprintf(fp, "    -c INT        drop a gene if average occurrence is >INT [%d]\n", o->max_avg_occ);
	fprintf(fp, "    -g INT        drop a gene if its in- or out-degree >INT [%d]\n", o->max_degree);
	fprintf(fp, "    -r INT        drop a gene if it connects >INT distant loci [%d]\n", o->max_dist_loci);
	fprintf(fp, "    -b FLOAT      demote a branching arc if weaker than the best by FLOAT [%g]\n", o->branch_diff);
	fprintf(fp, "    -B FLOAT      cut a branching arc if weaker by FLOAT [%g]\n", o->branch_diff_cut);
	fprintf(fp, "    -y FLOAT      cut a distant branching arc if weaker by FLOAT [%g]\n", o->branch_diff_dist);
	fprintf(fp, "    -T INT        apply branch cutting for INT times [%d]\n", o->n_branch_flt);
	fprintf(fp, "    -F            don't consider genes on different contigs as distant\n");
	fprintf(fp, "    -a INT        prune an arc if it is supported by <INT genomes [%d]\n", o->min_arc_cnt);
	fprintf(fp, "  Output:\n");
	fprintf(fp, "    -w            Suppress walk lines (W-lines)\n");
	fprintf(fp, "    --bed[=STR]   output 12-column BED where STR is walk, raw or flag [walk]\n");
	fprintf(fp, "    --version     print version number\n");
	return fp == stdout? 0 : 1;
}

static int64_t parse_num(const char *s)   // main.c:45-60
{
	char *p;
	double x = strtod(s, &p);
	if (*p == 'G' || *p == 'g') x *= 1e9; else if (*p == 'M' || *p == 'm') x *= 1e6; else if (*p == 'K' || *p == 'k') x *= 1e3;
	return (int64_t)(x + .499);
}

int main(int argc, char *argv[])
{
	static struct option lopts[] = { {"bed", optional_argument, 0, 301}, {"ori-sc", no_argument, 0, 302}, {"version", no_argument, 0, 401}, {0, 0, 0, 0} };
	pg_opt_t opt;
	int c;
	double t0 = now();
	pgio_opt_init(&opt);
	while ((c = getopt_long(argc, argv, "d:e:l:f:g:p:b:B:y:Fr:c:a:wv:GD:C:T:X:I:P:m:JOSE", lopts, 0)) >= 0) {
		if (c == 'd') opt.gene_delim = *optarg;
		else if (c == 'X') opt.excl = pgio_name_set(optarg);
		else if (c == 'I') opt.incl = pgio_name_set(optarg);
		else if (c == 'P') opt.preferred = pgio_name_set(optarg);
		else if (c == 'e') opt.min_prot_iden = atof(optarg);
		else if (c == 'l') opt.min_prot_ratio = atof(optarg);
		else if (c == 'm') opt.score_adj_coef = atof(optarg);
		else if (c == 'f') opt.min_ov_ratio = atof(optarg);
		else if (c == 'p') opt.min_vertex_ratio = atof(optarg);
		else if (c == 'c') opt.max_avg_occ = atoi(optarg);
		else if (c == 'g') opt.max_degree = atoi(optarg);
		else if (c == 'r') opt.max_dist_loci = atoi(optarg);
		else if (c == 'J') opt.flag |= PG_F_NO_JOINT_PSEUDO;
		else if (c == 'E') opt.flag |= PG_F_DROP_SGL_EXON;
		else if (c == 'b') opt.branch_diff = atof(optarg);
		else if (c == 'B') opt.branch_diff_cut = atof(optarg);
		else if (c == 'y') opt.branch_diff_dist = atof(optarg);
		else if (c == 'T') opt.n_branch_flt = (int32_t)atof(optarg);
		else if (c == 'a') opt.min_arc_cnt = atoi(optarg);
		else if (c == 'F') opt.flag |= PG_F_FRAG_MODE;
		else if (c == 'D') opt.local_dist = (int32_t)parse_num(optarg);
		else if (c == 'C') opt.local_count = atoi(optarg);
		else if (c == 'S') opt.flag |= PG_F_CHECK_STRAND;
		else if (c == 'w') opt.flag |= PG_F_WRITE_NO_WALK;
		else if (c == 'G') opt.flag |= PG_F_WRITE_VTX_SEL;
		else if (c == 'v') pg_verbose = pgio_verbose = atoi(optarg);
		else if (c == 301) {
			if (optarg == 0 || strcmp(optarg, "walk") == 0) opt.flag |= PG_F_WRITE_BED_WALK;
			else if (strcmp(optarg, "raw") == 0) opt.flag |= PG_F_WRITE_BED_RAW;
			else if (strcmp(optarg, "flag") == 0) opt.flag |= PG_F_WRITE_BED_FLAG;
			else { fprintf(stderr, "ERROR: unrecognized --bed argument. Should be 'raw' or 'walk'.\n"); return 1; }
		} else if (c == 302) opt.flag |= PG_F_ORI_FOR_BRANCH;
		else if (c == 401) { puts(PGX_VERSION); return 0; }
	}
	if (argc - optind < 1) return usage(stderr, &opt);

	pg_data_t *d = pgio_data_init();
	std::thread warm([]() { pgx_warmup(); });   // CUDA context + pinned staging ring come up while the files are parsed
	// GFA output (the default) never needs a pg_hit_t record: the reader emits the device-ready 48-byte records (pgx_packed.h), libpgx uploads them
	// as they are and hands back the graph and the walk tokens of the W lines.  The BED outputs print every field of every hit: pg_hit_t route.
	const bool gfa_route = !(opt.flag & (PG_F_WRITE_BED_RAW | PG_F_WRITE_BED_WALK | PG_F_WRITE_BED_FLAG));
	pgx_packed_t pk;
	memset(&pk, 0, sizeof(pk));
	if (gfa_route) pgio_read_pafs_packed(&opt, d, &pk, (const char *const*)&argv[optind], argc - optind, 0);
	else pgio_read_pafs(&opt, d, (const char *const*)&argv[optind], argc - optind, 0);   // parallel parse, ids as in sequential order
	warm.join();
	double t1 = now();
	const long n_hits_in = gfa_route? (long)pgio_packed_total_hits(&pk) : (long)pgio_total_hits(d);
	// --bed=raw: strict post-conditions after pg_post_process; other BED outputs: the hits come back once, after pg_graph_gen; GFA (the
	// default): no hit record comes back at all -- the W lines are formatted from the walk tokens libpgx reads off its final state
	const bool want_bed = (opt.flag & (PG_F_WRITE_BED_RAW | PG_F_WRITE_BED_WALK | PG_F_WRITE_BED_FLAG)) != 0;
	pgx_set_lazy((opt.flag & PG_F_WRITE_BED_RAW)? 0 : want_bed? 1 : 2);
	if (gfa_route) { pgx_post_process_packed(&opt, d, &pk); pgio_packed_destroy(&pk); }   // (the records are on the device now)
	else pg_post_process(&opt, d);
	if (opt.flag & PG_F_WRITE_BED_RAW) {
		pgio_write_bed(d, 0, "-");
		pgx_data_release(d);                          // no pg_graph_gen() follows: hand the device state back
	} else {
		pg_graph_t *g = pg_graph_init(d);
		pg_graph_gen(&opt, g);
		double t2 = now();
		if (opt.flag & PG_F_WRITE_BED_WALK) pgio_write_bed(d, 1, "-");
		else if (opt.flag & PG_F_WRITE_BED_FLAG) pgio_write_bed(d, 0, "-");
		else if (opt.flag & PG_F_WRITE_NO_WALK) { pgio_write_gfa_tokens(g, 0, 0, "-", 1); pgx_data_release(d); }
		else {
			pgx_walk_token_t *tok = 0;
			std::vector<int64_t> off((size_t)d->n_genome + 1);
			pgx_walk_tokens(d, &tok, off.data());
			pgx_data_release(d);
			pgio_write_gfa_tokens(g, tok, off.data(), "-", 3);
			free(tok);
		}
		if (pg_verbose >= 3) fprintf(stderr, "[M::%s] parse %.3f s; hot path (incl. transfers) %.3f s; %ld hits\n", __func__, t1 - t0, t2 - t1, n_hits_in);
		pg_graph_destroy(g);
	}
	pgio_data_destroy(d);
	if (opt.excl) pgio_name_set_destroy(opt.excl);
	if (opt.incl) pgio_name_set_destroy(opt.incl);
	if (opt.preferred) pgio_name_set_destroy(opt.preferred);
	if (pg_verbose >= 3) {
		struct rusage r;
		getrusage(RUSAGE_SELF, &r);
		fprintf(stderr, "[M::%s] Version: %s\n[M::%s] CMD:", __func__, PGX_VERSION, __func__);
		for (int i = 0; i < argc; ++i) fprintf(stderr, " %s", argv[i]);
		fprintf(stderr, "\n[M::%s] Real time: %.3f sec; CPU: %.3f sec; Peak RSS: %.3f GB\n", __func__, now() - t0,
			r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec), r.ru_maxrss / 1024.0 / 1024.0);
	}
	return 0;
}
