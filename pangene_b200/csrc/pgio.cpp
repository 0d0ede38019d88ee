// pgio.cpp -- host-side I/O around the hot path: PAF reader, GFA/BED writers, gfa2matrix, option
// defaults.  Plain C++17 + zlib, exported with a C ABI (include/pgio.h).  No CUDA here.
//
// Reference behaviour being matched (cited per function): read.c:107-241 (reader, WITHOUT the
// postprocessing block at read.c:243-260, which is hot-path work done by libpgx), option.c:6-26,
// format.c:78-225 (writers), pangene.js:1168-1246 (gfa2matrix).  Written from scratch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cassert>
#include <string>
#include <string_view>
#include <vector>
#include <deque>
#include <unordered_map>
#include <algorithm>
#include <thread>
#include <atomic>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <time.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include "pgio.h"
#include "pgio_priv.hpp"

int pgio_verbose = 3;

// ------------------------------------------------------------------------------------------------
// options
// ------------------------------------------------------------------------------------------------
extern "C" void pgio_opt_init(pg_opt_t *o)   // option.c:6-26
{
	memset(o, 0, sizeof(*o));
	o->gene_delim = ':';
	o->min_prot_ratio = 0.5;  o->min_prot_iden = 0.5;   o->score_adj_coef = 2.0;
	o->min_ov_ratio = 0.5;    o->min_vertex_ratio = 0.05;
	o->branch_diff = 0.02;    o->branch_diff_dist = 0.05; o->branch_diff_cut = 0.5;
	o->max_avg_occ = 10;      o->max_degree = 15;       o->max_dist_loci = 3;
	o->n_branch_flt = 15;     o->min_arc_cnt = 1;
	o->local_dist = 2000000;  o->local_count = 10;
}

// ------------------------------------------------------------------------------------------------
// containers
// ------------------------------------------------------------------------------------------------
uint64_t pgio::NameTable::hash_of(std::string_view s)   // 8 bytes at a time, multiply-fold (names are 5-40 bytes)
{
	const char *p = s.data();
	size_t n = s.size();
	uint64_t h = 0x9e3779b97f4a7c15ull ^ (uint64_t)n * 0xff51afd7ed558ccdull;
	while (n >= 8) { uint64_t w; memcpy(&w, p, 8); h = (h ^ w) * 0x9fb21c651e98df25ull; h ^= h >> 29; p += 8; n -= 8; }
	if (n) { uint64_t w = 0; memcpy(&w, p, n); h = (h ^ w) * 0x9fb21c651e98df25ull; h ^= h >> 29; }
	h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 32;
	return h;
}
void pgio::NameTable::grow()
{
	const size_t cap = slots.empty()? 1024 : slots.size() * 2;
	std::vector<Slot> fresh(cap, Slot{0, -1});
	for (const Slot &o : slots) {
		if (o.id < 0) continue;
		size_t k = o.hash & (cap - 1);
		while (fresh[k].id >= 0) k = (k + 1) & (cap - 1);
		fresh[k] = o;
	}
	slots.swap(fresh);
}
const char *pgio::NameTable::intern(std::string_view s, int32_t *id, bool *absent)
{
	if ((names.size() + 1) * 2 > slots.size()) grow();   // load factor <= 1/2
	const uint64_t h = hash_of(s);
	const size_t mask = slots.size() - 1;
	size_t k = h & mask;
	for (; slots[k].id >= 0; k = (k + 1) & mask) {
		const int32_t j = slots[k].id;
		if (slots[k].hash == h && lens[j] == s.size() && memcmp(names[j], s.data(), s.size()) == 0) {
			if (id) *id = j;
			if (absent) *absent = false;
			return names[j];
		}
	}
	char *p = (char*)malloc(s.size() + 1);
	memcpy(p, s.data(), s.size()); p[s.size()] = 0;
	const int32_t j = (int32_t)names.size();
	names.push_back(p); lens.push_back((uint32_t)s.size());
	slots[k] = Slot{h, j};
	if (id) *id = j;
	if (absent) *absent = true;
	return p;
}
int32_t pgio::NameTable::find(std::string_view s) const
{
	if (slots.empty()) return -1;
	const uint64_t h = hash_of(s);
	const size_t mask = slots.size() - 1;
	for (size_t k = h & mask; slots[k].id >= 0; k = (k + 1) & mask) {
		const int32_t j = slots[k].id;
		if (slots[k].hash == h && lens[j] == s.size() && memcmp(names[j], s.data(), s.size()) == 0) return j;
	}
	return -1;
}
pgio::NameTable::~NameTable() { for (char *p : names) free(p); }

extern "C" pg_data_t *pgio_data_init(void)   // read.c:10-17
{
	pg_data_t *d = (pg_data_t*)calloc(1, sizeof(pg_data_t));
	pgio::DataPriv *pv = new pgio::DataPriv();
	d->d_ctg = pv;           // all three dictionaries live in one private object
	d->d_gene = d->d_prot = 0;
	return d;
}

extern "C" void pgio_data_destroy(pg_data_t *d)   // read.c:20-32
{
	if (d == 0) return;
	for (int32_t i = 0; i < d->n_genome; ++i) {
		pg_genome_t *g = &d->genome[i];
		free(g->ctg); free(g->hit); free(g->exon); free(g->label);
	}
	free(d->genome); free(d->gene); free(d->prot);
	delete (pgio::DataPriv*)d->d_ctg;
	free(d);
}

extern "C" int64_t pgio_total_hits(const pg_data_t *d)
{
	int64_t n = 0;
	for (int32_t j = 0; j < d->n_genome; ++j) n += d->genome[j].n_hit;
	return n;
}

extern "C" pg_data_t *pgio_data_clone(const pg_data_t *s)
{
	pg_data_t *d = (pg_data_t*)calloc(1, sizeof(pg_data_t));
	*d = *s;
	d->m_genome = d->n_genome; d->m_gene = d->n_gene; d->m_prot = d->n_prot;
	d->genome = (pg_genome_t*)malloc(sizeof(pg_genome_t) * (d->n_genome > 0? d->n_genome : 1));
	d->gene = (pg_gene_t*)malloc(sizeof(pg_gene_t) * (d->n_gene > 0? d->n_gene : 1));
	d->prot = (pg_prot_t*)malloc(sizeof(pg_prot_t) * (d->n_prot > 0? d->n_prot : 1));
	memcpy(d->gene, s->gene, sizeof(pg_gene_t) * d->n_gene);
	memcpy(d->prot, s->prot, sizeof(pg_prot_t) * d->n_prot);
	for (int32_t j = 0; j < d->n_genome; ++j) {
		const pg_genome_t *a = &s->genome[j];
		pg_genome_t *b = &d->genome[j];
		*b = *a;
		b->m_ctg = a->n_ctg; b->m_hit = a->n_hit; b->m_exon = a->n_exon;
		b->ctg = (pg_ctg_t*)malloc(sizeof(pg_ctg_t) * (a->n_ctg > 0? a->n_ctg : 1));
		b->hit = (pg_hit_t*)malloc(sizeof(pg_hit_t) * (a->n_hit > 0? a->n_hit : 1));
		b->exon = (pg_exon_t*)malloc(sizeof(pg_exon_t) * (a->n_exon > 0? a->n_exon : 1));
		memcpy(b->ctg, a->ctg, sizeof(pg_ctg_t) * a->n_ctg);
		memcpy(b->hit, a->hit, sizeof(pg_hit_t) * a->n_hit);
		memcpy(b->exon, a->exon, sizeof(pg_exon_t) * a->n_exon);
		b->label = a->label? strdup(a->label) : 0;
	}
	d->d_ctg = d->d_gene = d->d_prot = 0;   // names stay owned by the source
	return d;
}

extern "C" void pgio_data_clone_destroy(pg_data_t *d)
{
	if (d == 0) return;
	for (int32_t i = 0; i < d->n_genome; ++i) {
		pg_genome_t *g = &d->genome[i];
		free(g->ctg); free(g->hit); free(g->exon); free(g->label);
	}
	free(d->genome); free(d->gene); free(d->prot);
	free(d);
}

// ------------------------------------------------------------------------------------------------
// name sets for -X / -I / -P  (read.c:264-318)
// ------------------------------------------------------------------------------------------------
static bool read_whole(const char *fn, std::string &out)
{
	out.clear();
	if (fn && strcmp(fn, "-")) {   // a regular, uncompressed file goes straight into the buffer (no zlib pass-through copy)
		int fd = open(fn, O_RDONLY);
		if (fd < 0) return false;
		struct stat sb;
		unsigned char magic[2] = {0, 0};
		if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size >= 2 && pread(fd, magic, 2, 0) == 2 && !(magic[0] == 0x1f && magic[1] == 0x8b)) {
			out.resize((size_t)sb.st_size);
			size_t got = 0;
			while (got < out.size()) {
				ssize_t r = read(fd, &out[got], out.size() - got);
				if (r <= 0) break;
				got += (size_t)r;
			}
			out.resize(got);
			close(fd);
			return true;
		}
		close(fd);
	}
	gzFile fp = (fn && strcmp(fn, "-"))? gzopen(fn, "r") : gzdopen(0, "r");
	if (fp == 0) return false;
	gzbuffer(fp, 1 << 20);
	std::vector<char> buf(1 << 20);
	int n;
	while ((n = gzread(fp, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), n);
	gzclose(fp);
	return true;
}

// line iteration with the semantics of the reference's line reader (kseq KS_SEP_LINE): '\n' ends a
// line, one trailing '\r' is dropped, a last line without '\n' still counts, no phantom empty line.
template<class F> static void for_each_line(std::string &buf, F f)
{
	size_t n = buf.size(), st = 0;
	while (st < n) {
		const char *nl = (const char*)memchr(buf.data() + st, '\n', n - st);
		size_t en = nl? (size_t)(nl - buf.data()) : n;
		size_t e2 = en;
		if (e2 > st && buf[e2 - 1] == '\r') --e2;
		char saved = buf[e2]; // buf has a terminating 0 at buf[n] (std::string guarantee), so e2 <= n is fine
		buf[e2] = 0;
		f(&buf[st], e2 - st);
		if (e2 < n) buf[e2] = saved;
		st = en + 1;
	}
}

extern "C" void *pgio_name_set(const char *o)
{
	pgio::NameTable *t = new pgio::NameTable();
	if (*o != '@') {
		const char *q = o;
		for (const char *p = q;; ++p) {
			if (*p == ',' || *p == ' ' || *p == '\t' || *p == 0) {
				if (p > q) t->intern(std::string_view(q, p - q), 0, 0);
				if (*p == 0) break;
				q = p + 1;
			}
		}
	} else {
		std::string buf;
		if (!read_whole(o + 1, buf)) { delete t; return 0; }
		for_each_line(buf, [&](char *s, size_t) {
			char *p = s;
			while (*p && !isspace((unsigned char)*p)) ++p;
			t->intern(std::string_view(s, p - s), 0, 0);
		});
	}
	return t;
}

extern "C" void pgio_name_set_destroy(void *s) { delete (pgio::NameTable*)s; }

// ------------------------------------------------------------------------------------------------
// reader
// ------------------------------------------------------------------------------------------------
// strtol(p, &end, 10) for the common case "digits"; anything else (sign, blanks, 19+ digits) goes to strtol itself
static inline long fast_l(const char *p, const char **end = 0)
{
	const char *q = p;
	unsigned long u = 0;                                 // (unsigned: a run of 20+ digits wraps harmlessly before strtol takes over)
	while ((unsigned)(*q - '0') <= 9u) u = u * 10 + (unsigned long)(*q++ - '0');
	long v = (long)u;
	if (q == p || q - p > 18) { char *e; v = strtol(p, &e, 10); q = e; }
	if (end) *end = q;
	return v;
}

template<class T> static inline void grow0(T *&p, int64_t i, int32_t &m)   // pgpriv.h:28-36
{
	if (i >= m) {
		int64_t old = m, nm = i + 1;
		nm += (nm >> 1) + 16;
		p = (T*)realloc(p, nm * sizeof(T));
		memset((void*)(p + old), 0, (nm - old) * sizeof(T));
		m = (int32_t)nm;
	}
}

// mid-point of the CDS in contig coordinates (hit.c:14-27)
int64_t pgio::cds_mid(const pg_hit_t *a, const pg_exon_t *e)
{
	int32_t tot = 0;
	for (int32_t i = 0; i < a->n_exon; ++i) tot += e[i].oe - e[i].os;
	int32_t half = tot >> 1, acc = 0;
	for (int32_t i = 0; i < a->n_exon; ++i) {
		int32_t l = e[i].oe - e[i].os;
		if (acc <= half && half < acc + l) return a->cs + e[i].os + (half - acc);
		acc += l;
	}
	fprintf(stderr, "[pgio] ERROR: alignment with empty CDS\n");
	abort();
}

// CIGAR -> exon list (read.c:47-90). Returns the reference-consumed length x; exons land in `out`
// in contig orientation; *n_fs = number of F/G operations.
int64_t pgio::cigar_to_exons(const char *cg, bool rev, std::vector<pg_exon_t> &tmp, std::vector<pg_exon_t> &out, int32_t *n_fs)
{
	int64_t x = 0;
	int32_t fs = 0;
	size_t n = 1;                                    // exons so far; tmp[n-1] is the open one
	if (tmp.size() < 64) tmp.resize(64);
	tmp[0] = pg_exon_t{0, 0};
	const char *p = cg;
	while (*p) {
		uint64_t lu = 0;
		const char *r = p;
		while ((unsigned)(*r - '0') <= 9u) lu = lu * 10 + (uint64_t)(*r++ - '0');
		int64_t l = (int64_t)lu;
		if (r == p || r - p > 18) l = fast_l(p, &r);   // sign, blanks, overflow: strtol's own answer
		const char op = *r;
		if (op == 'M' || op == 'X' || op == '=' || op == 'D') {
			x += l * 3;
		} else if (op == 'N' || op == 'U' || op == 'V') {
			int64_t st, en;
			if (op == 'N') st = x, en = x + l;
			else if (op == 'U') st = x + 1, en = x + l - 2;
			else st = x + 2, en = x + l - 1;
			tmp[n - 1].oe = (int32_t)st;
			if (n == tmp.size()) tmp.resize(n * 2);
			tmp[n++] = pg_exon_t{(int32_t)en, (int32_t)en};
			x += l;
		} else if (op == 'F' || op == 'G') {
			x += l, ++fs;
		}
		if (op == 0) break;   // malformed tail; the reference would run off the string here
		p = r + 1;
	}
	tmp[n - 1].oe = (int32_t)x;
	out.resize(n);
	if (!rev) memcpy(out.data(), tmp.data(), n * sizeof(pg_exon_t));
	else for (size_t i = 0; i < n; ++i) out[i] = pg_exon_t{(int32_t)(x - tmp[n - 1 - i].oe), (int32_t)(x - tmp[n - 1 - i].os)};
	*n_fs = fs;
	return x;
}

int32_t pgio::adjusted_score(const pg_opt_t *opt, int32_t ms, int32_t mlen, int32_t blen, int32_t qs, int32_t qe, int32_t plen)
{   // read.c:212-216: (int32_t)(ms * expl(x) + .499), evaluated in long double
	double div = 1.0 - (double)mlen / blen;
	double uncov = 1.0 - (double)(qe - qs) / plen;
	const double x = -opt->score_adj_coef * (div + uncov);
	// expl() costs ~250 cycles per line.  exp() is within 1 ulp (1.2e-16 relative), so for |ms * exp(x)| < 1e8 the double value is within
	// ~5e-8 of the long-double one; unless it lands within 1e-6 of an integer, truncation gives the same int32.
	if (ms > -4194304 && ms < 4194304 && x == x && x < 40.0 && x > -700.0) {
		const double v = ms * exp(x) + .499;
		const double r = nearbyint(v);
		if (fabs(v - r) > 1e-6 && fabs(v) < 1e8) return (int32_t)v;
	}
	return (int32_t)(ms * expl(x) + .499);
}

static char *label_from_path(const char *fn)   // read.c:92-105
{
	int32_t len = (int32_t)strlen(fn), en = len, st;
	int32_t i;
	for (i = len - 1; i >= 0 && fn[i] != '/'; --i) {}
	st = i + 1;
	if (en >= 3 && strncmp(&fn[en - 3], ".gz", 3) == 0) en -= 3;
	if (en >= 4 && strncmp(&fn[en - 4], ".paf", 4) == 0) en -= 4;
	if (st >= en) return 0;
	char *s = (char*)calloc(en - st + 1, 1);
	memcpy(s, &fn[st], en - st);
	return s;
}

void pgio::push_hit(pg_genome_t *g, pg_hit_t &hit, const std::vector<pg_exon_t> &ex, int32_t cg_fs, int32_t n_fs, int32_t n_stop)
{
	grow0(g->exon, (int64_t)g->n_exon + (int64_t)ex.size() - 1, g->m_exon);
	memcpy(&g->exon[g->n_exon], ex.data(), ex.size() * sizeof(pg_exon_t));
	hit.n_exon = (int32_t)ex.size();
	hit.off_exon = g->n_exon;
	g->n_exon += (int32_t)ex.size();
	int32_t lof = (n_fs > 0? n_fs : 0) + (n_stop > 0? n_stop : 0);   // read.c:229-230
	hit.lof = cg_fs > lof? cg_fs : lof;
	hit.cm = pgio::cds_mid(&hit, &g->exon[hit.off_exon]);
	grow0(g->hit, g->n_hit, g->m_hit);
	g->hit[g->n_hit++] = hit;
}

void pgio::push_hit_packed(pg_genome_t *g, pgx_packed_hit_t *&pk, int32_t &n, int32_t &m, pg_hit_t &hit, const std::vector<pg_exon_t> &ex, int32_t cg_fs, int32_t n_fs, int32_t n_stop)
{
	grow0(g->exon, (int64_t)g->n_exon + (int64_t)ex.size() - 1, g->m_exon);
	memcpy(&g->exon[g->n_exon], ex.data(), ex.size() * sizeof(pg_exon_t));
	hit.n_exon = (int32_t)ex.size();
	hit.off_exon = g->n_exon;
	g->n_exon += (int32_t)ex.size();
	int32_t lof = (n_fs > 0? n_fs : 0) + (n_stop > 0? n_stop : 0);   // read.c:229-230
	hit.lof = cg_fs > lof? cg_fs : lof;
	hit.cm = pgio::cds_mid(&hit, &g->exon[hit.off_exon]);
	if (hit.cs < 0 || hit.ce > 0x7fffffffll || hit.cm < 0 || hit.cm > 0x7fffffffll || hit.cs > hit.ce) {
		fprintf(stderr, "[pgio] ERROR: a hit coordinate is outside [0, 2^31) or cs > ce -- device-ready records keep coordinates in int32\n"); abort();
	}
	grow0(pk, n, m);
	pgx_packed_hit_t &o = pk[n++];
	o.pid = hit.pid; o.cid = hit.cid; o.rank = hit.rank; o.score_ori = hit.score_ori; o.score_adj = hit.score_adj; o.n_exon = hit.n_exon; o.off_exon = hit.off_exon;
	o.rev = hit.rev; o.cs = (int32_t)hit.cs; o.cm = (int32_t)hit.cm; o.ce = (int32_t)hit.ce; o.lof = hit.lof;
}

// tab-separated fields of one NUL-terminated line, split in place (the tabs become terminators); like the reference's column
// loop (read.c:135-137) the line ends at its first NUL byte
static inline void split_tabs(char *s, size_t len, std::vector<char*> &fld)
{
	fld.clear();
	fld.push_back(s);
	if (const void *z = memchr(s, 0, len)) len = (size_t)((const char*)z - s);
	size_t i = 0;
#if defined(__SSE2__)
	const __m128i tab = _mm_set1_epi8('\t');
	for (; i + 16 <= len; i += 16) {
		unsigned m = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(s + i)), tab));
		while (m) { const unsigned b = (unsigned)__builtin_ctz(m); m &= m - 1; s[i + b] = 0; fld.push_back(s + i + b + 1); }
	}
#endif
	for (; i < len; ++i) if (s[i] == '\t') { s[i] = 0; fld.push_back(s + i + 1); }
}

// One PAF file parsed on its own, with FILE-LOCAL gene/protein ids (first-seen order inside the file).  Files are
// independent except for the global id assignment (ids = dictionary size at first sight, read.c:151,161; hit.rank is a
// per-file count, read.c:170), so they can be parsed in parallel and merged in command-line order afterwards
// (SURVEY.md §8f rank 1): merging the per-file first-seen lists in file order reproduces the sequential ids exactly.
namespace {
struct LocalFile {
	bool ok = false;
	pg_genome_t g;
	pgio::NameTable gene, prot, ctg;                  // local dictionaries (own their strings until the merge)
	std::vector<int32_t> prot_gid, prot_len;          // per local protein: local gene id, length as left by its LAST line (read.c:166,174)
	std::vector<int32_t> gene_len;                    // per local gene: max protein length seen (read.c:175)
	std::vector<uint8_t> gene_pref, gene_incl;
	int64_t n_lines = 0;
	bool packed = false;                              // emit device-ready records (pgx_packed.h) instead of pg_hit_t
	pgx_packed_hit_t *ph = nullptr; int32_t n_ph = 0, m_ph = 0;
	LocalFile() { memset(&g, 0, sizeof(g)); }
};

std::string &file_buffer() { static thread_local std::string b; return b; }

void parse_file(const pg_opt_t *opt, const char *fn, LocalFile &F)
{
	std::string &buf = file_buffer();                              // one per worker thread: the pages of the previous file are reused
	if (!read_whole(fn, buf)) return;                              // read.c:117-118
	F.ok = true;
	const pgio::NameTable *excl = (const pgio::NameTable*)opt->excl, *incl = (const pgio::NameTable*)opt->incl, *pref = (const pgio::NameTable*)opt->preferred;
	pg_genome_t *g = &F.g;
	g->label = fn && strcmp(fn, "-")? label_from_path(fn) : label_from_path("-");
	{   // size the hit and exon arrays once from a sample of the text (calloc: untouched pages cost nothing; PG_GROW0-style growth stays as the fallback)
		const size_t smp = buf.size() < ((size_t)1 << 16)? buf.size() : ((size_t)1 << 16);
		size_t nl = 0, nn = 0;
		for (size_t i = 0; i < smp; ++i) { nl += buf[i] == '\n'; nn += buf[i] == 'N'; }
		if (nl > 0 && buf.size() / smp < 100000) {
			const double lines = (double)buf.size() / smp * (double)nl;
			const int64_t eh = (int64_t)(lines * 1.03) + 64, ee = (int64_t)(lines * 1.05 * (1.0 + (double)nn / nl)) + 64;
			if (eh < 0x7fffffff && ee < 0x7fffffff) {
				if (F.packed) { F.ph = (pgx_packed_hit_t*)calloc((size_t)eh, sizeof(pgx_packed_hit_t)); F.m_ph = F.ph? (int32_t)eh : 0; }
				else { g->hit = (pg_hit_t*)calloc((size_t)eh, sizeof(pg_hit_t)); g->m_hit = g->hit? (int32_t)eh : 0; }
				g->exon = (pg_exon_t*)calloc((size_t)ee, sizeof(pg_exon_t)); g->m_exon = g->exon? (int32_t)ee : 0;
			}
		}
	}
	std::vector<int32_t> seen_cnt;                                // per-protein line count in THIS file (rank, read.c:170)
	std::vector<pg_exon_t> tmp, ex;
	std::vector<char*> fld;
	int32_t last_pid = -1, last_gid = -1;
	std::string_view last_pname, last_gname;                      // views into `buf` (earlier lines stay intact)

	for_each_line(buf, [&](char *s, size_t len) {
		++F.n_lines;
		split_tabs(s, len, fld);
		pg_hit_t hit;
		memset(&hit, 0, sizeof(hit));
		hit.pid = hit.pid_dom = hit.cid = hit.off_exon = hit.n_exon = -1;
		int32_t n_fs = -1, n_stop = -1, cg_fs = 0, pid = -1, gid = -1;
		bool have_exons = false;
		const int nf = (int)fld.size();
		do {
			// ---- column 0: "gene<delim>protein" (read.c:139-171)
			char *q = fld[0], *r = q;
			if (excl && excl->find(q) >= 0) break;
			while (*r && *r != opt->gene_delim) ++r;
			std::string_view gname(q, r - q);
			if (excl && excl->find(gname) >= 0) break;
			bool absent;
			// (alignments of one protein, and of the isoforms of one gene, tend to sit on consecutive lines: remember the last of each)
			const std::string_view pname(q);
			if (last_pid >= 0 && pname == last_pname) pid = last_pid, gid = last_gid;
			else {
				if (last_gid >= 0 && gname == last_gname) gid = last_gid;
				else {
					F.gene.intern(gname, &gid, &absent);
					if (absent) { F.gene_len.push_back(0); F.gene_pref.push_back(pref && pref->find(gname) >= 0); F.gene_incl.push_back(incl && incl->find(gname) >= 0); }
					last_gid = gid; last_gname = gname;
				}
				F.prot.intern(pname, &pid, &absent);
				if (absent) { F.prot_gid.push_back(gid); F.prot_len.push_back(0); }
				last_pid = pid; last_pname = pname;
			}
			F.prot_gid[pid] = gid; F.prot_len[pid] = 0;
			hit.pid = pid;
			if ((size_t)pid >= seen_cnt.size()) seen_cnt.resize(pid + 1 + (pid >> 1), -1);
			hit.rank = ++seen_cnt[pid];
			if (nf < 2) break;
			// ---- columns 1..10 (read.c:172-211)
			int32_t len = (int32_t)fast_l(fld[1]);
			F.prot_len[pid] = len;
			if (F.gene_len[gid] < len) F.gene_len[gid] = len;
			if (nf < 3) break;
			hit.qs = (int32_t)fast_l(fld[2]);
			if (nf < 4) break;
			hit.qe = (int32_t)fast_l(fld[3]);
			if (hit.qe - hit.qs < len * opt->min_prot_ratio) break;
			if (nf < 5) break;
			if (fld[4][0] != '+' && fld[4][0] != '-') break;
			hit.rev = fld[4][0] == '+'? 0 : 1;
			if (nf < 6) break;
			{
				int32_t cid; bool cabs;
				const char *cn = F.ctg.intern(fld[5], &cid, &cabs);
				if (cabs) { grow0(g->ctg, g->n_ctg, g->m_ctg); g->ctg[g->n_ctg++].name = cn; }
				hit.cid = cid;
			}
			if (nf < 7) break;
			g->ctg[hit.cid].len = fast_l(fld[6]);
			if (nf < 8) break;
			hit.cs = fast_l(fld[7]);
			if (nf < 9) break;
			hit.ce = fast_l(fld[8]);
			if (nf < 10) break;
			hit.mlen = (int32_t)fast_l(fld[9]);
			if (nf < 11) break;
			hit.blen = (int32_t)fast_l(fld[10]);
			if (hit.mlen < hit.blen * opt->min_prot_iden) break;
			// ---- tags (read.c:212-224)
			for (int i = 12; i < nf; ++i) {
				const char *t = fld[i];
				if (t[0] == 0 || t[1] == 0 || t[2] != ':' || t[3] == 0 || t[4] != ':') continue;   // none of ms:i: fs:i: st:i: cg:Z: (never reads past the field's terminator)
				if (t[0] == 'm' && t[1] == 's' && t[3] == 'i') {
					hit.score_ori = (int32_t)fast_l(t + 5);
					hit.score_adj = pgio::adjusted_score(opt, hit.score_ori, hit.mlen, hit.blen, hit.qs, hit.qe, len);
				} else if (t[0] == 'f' && t[1] == 's' && t[3] == 'i') n_fs = (int32_t)fast_l(t + 5);
				else if (t[0] == 's' && t[1] == 't' && t[3] == 'i') n_stop = (int32_t)fast_l(t + 5);
				else if (t[0] == 'c' && t[1] == 'g' && t[3] == 'Z') {
					int64_t x = pgio::cigar_to_exons(t + 5, hit.rev, tmp, ex, &cg_fs);
					if (x != hit.ce - hit.cs) { fprintf(stderr, "[pgio] ERROR: CIGAR length %ld != ce-cs %ld\n", (long)x, (long)(hit.ce - hit.cs)); abort(); } // read.c:75
					have_exons = true;
				}
			}
		} while (0);
		if (have_exons) { if (F.packed) pgio::push_hit_packed(g, F.ph, F.n_ph, F.m_ph, hit, ex, cg_fs, n_fs, n_stop); else pgio::push_hit(g, hit, ex, cg_fs, n_fs, n_stop); }
	});
}

// fold one parsed file into the global tables, in command-line order (serial, cheap: names only), then rewrite its hits' ids
void merge_file(pg_data_t *d, LocalFile &F, pgx_packed_t *pk = nullptr)
{
	pgio::DataPriv *pv = (pgio::DataPriv*)d->d_ctg;
	std::vector<int32_t> g2g(F.gene.names.size()), p2p(F.prot.names.size());
	for (size_t lg = 0; lg < F.gene.names.size(); ++lg) {
		int32_t gid; bool absent;
		const char *nm = pv->gene.intern(F.gene.names[lg], &gid, &absent);
		if (absent) { d->n_gene++; grow0(d->gene, gid, d->m_gene); }
		d->gene[gid].name = nm; d->gene[gid].preferred = F.gene_pref[lg]; d->gene[gid].included = F.gene_incl[lg];
		if ((int32_t)d->gene[gid].len < F.gene_len[lg]) d->gene[gid].len = F.gene_len[lg];
		g2g[lg] = gid;
	}
	for (size_t lp = 0; lp < F.prot.names.size(); ++lp) {
		int32_t pid; bool absent;
		const char *nm = pv->prot.intern(F.prot.names[lp], &pid, &absent);
		if (absent) { d->n_prot++; grow0(d->prot, pid, d->m_prot); }
		d->prot[pid].name = nm; d->prot[pid].gid = g2g[F.prot_gid[lp]]; d->prot[pid].len = F.prot_len[lp];
		p2p[lp] = pid;
	}
	grow0(d->genome, d->n_genome, d->m_genome);
	pg_genome_t *g = &d->genome[d->n_genome++];
	*g = F.g;
	for (int32_t c = 0; c < g->n_ctg; ++c) g->ctg[c].name = pv->ctg.intern(g->ctg[c].name, 0, 0);   // global, stable storage
	for (int32_t i = 0; i < g->n_hit; ++i) g->hit[i].pid = p2p[g->hit[i].pid];
	if (pk) {   // device-ready records: the genome's hit array stays empty, its records go to pk->hit[genome]
		for (int32_t i = 0; i < F.n_ph; ++i) F.ph[i].pid = p2p[F.ph[i].pid];
		const int32_t j = d->n_genome - 1;
		if (j >= pk->m_genome) {
			const int32_t m = pk->m_genome? pk->m_genome * 2 : 16;
			pk->n_hit = (int32_t*)realloc(pk->n_hit, sizeof(int32_t) * m); pk->m_hit = (int32_t*)realloc(pk->m_hit, sizeof(int32_t) * m);
			pk->hit = (pgx_packed_hit_t**)realloc(pk->hit, sizeof(pgx_packed_hit_t*) * m);
			pk->m_genome = m;
		}
		pk->hit[j] = F.ph; pk->n_hit[j] = F.n_ph; pk->m_hit[j] = F.m_ph; pk->n_genome = j + 1;
		F.ph = nullptr; F.n_ph = F.m_ph = 0;
	}
	memset(&F.g, 0, sizeof(F.g));
	if (pgio_verbose >= 4)
		fprintf(stderr, "[M::pgio_read_paf] [%d] %s: %ld lines parsed, %d hits kept\n", d->n_genome - 1, g->label? g->label : "*", (long)F.n_lines, g->n_hit);
}
} // namespace

extern "C" int32_t pgio_read_paf(const pg_opt_t *opt, pg_data_t *d, const char *fn)
{
	LocalFile F;
	parse_file(opt, fn, F);
	if (!F.ok) return -1;
	merge_file(d, F);
	if (file_buffer().capacity() > ((size_t)256 << 20)) std::string().swap(file_buffer());   // keep a modest buffer for the next file, not a huge one
	return 0;
}

static double wall_now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }

static int32_t read_pafs_impl(const pg_opt_t *opt, pg_data_t *d, pgx_packed_t *pk, const char *const *fns, int32_t n_files, int32_t n_threads);

// All files of a command line: parsed by n_threads workers (0 = hardware concurrency, capped at the file count), merged in
// order.  Returns the number of files that could be opened (the others are skipped, like main.c:120 ignores the error).
extern "C" int32_t pgio_read_pafs(const pg_opt_t *opt, pg_data_t *d, const char *const *fns, int32_t n_files, int32_t n_threads)
{
	return read_pafs_impl(opt, d, nullptr, fns, n_files, n_threads);
}

extern "C" int32_t pgio_read_pafs_packed(const pg_opt_t *opt, pg_data_t *d, pgx_packed_t *pk, const char *const *fns, int32_t n_files, int32_t n_threads)
{
	if (pk->n_genome != d->n_genome) { fprintf(stderr, "[pgio] ERROR: pgio_read_pafs_packed(): the packed set and the pg_data_t must grow together\n"); abort(); }
	return read_pafs_impl(opt, d, pk, fns, n_files, n_threads);
}

extern "C" void pgio_packed_destroy(pgx_packed_t *pk)
{
	for (int32_t j = 0; j < pk->n_genome; ++j) free(pk->hit[j]);
	free(pk->hit); free(pk->n_hit); free(pk->m_hit);
	memset(pk, 0, sizeof(*pk));
}

extern "C" int64_t pgio_packed_total_hits(const pgx_packed_t *pk)
{
	int64_t n = 0;
	for (int32_t j = 0; j < pk->n_genome; ++j) n += pk->n_hit[j];
	return n;
}

static int32_t read_pafs_impl(const pg_opt_t *opt, pg_data_t *d, pgx_packed_t *pk, const char *const *fns, int32_t n_files, int32_t n_threads)
{
	if (n_threads <= 0) n_threads = (int32_t)std::thread::hardware_concurrency();
	if (n_threads > n_files) n_threads = n_files;
	if (n_threads < 1) n_threads = 1;
	std::vector<LocalFile> files(n_files > 0? n_files : 0);
	for (auto &f : files) f.packed = pk != nullptr;
	std::atomic<int32_t> next(0);
	const double t0 = wall_now();
	auto worker = [&]() { for (int32_t i; (i = next.fetch_add(1)) < n_files;) parse_file(opt, fns[i], files[i]); };
	std::vector<std::thread> th;
	for (int32_t t = 1; t < n_threads; ++t) th.emplace_back(worker);
	worker();
	for (auto &t : th) t.join();
	std::string().swap(file_buffer());   // the calling thread's read buffer (the workers' went with their threads)
	const double t1 = wall_now();
	int32_t n_ok = 0;
	for (int32_t i = 0; i < n_files; ++i) if (files[i].ok) { merge_file(d, files[i], pk); ++n_ok; }
	for (auto &f : files) free(f.ph);                     // (files that could not be opened hold nothing)
	if (getenv("PGIO_TIMING")) fprintf(stderr, "[pgio timing] %d files on %d threads: parse %.3f s, merge %.3f s\n", n_files, n_threads, t1 - t0, wall_now() - t1);
	return n_ok;
}

// ------------------------------------------------------------------------------------------------
// writers
// ------------------------------------------------------------------------------------------------
namespace {
struct Out {
	FILE *fp; bool close_it; std::string s;
	explicit Out(const char *path) {
		if (path == 0 || strcmp(path, "-") == 0) fp = stdout, close_it = false;
		else fp = fopen(path, "w"), close_it = true;
	}
	~Out() { flush(); if (fp && close_it) fclose(fp); }
	void flush() { if (fp && !s.empty()) fwrite(s.data(), 1, s.size(), fp); s.clear(); }
	void maybe_flush() { if (s.size() > (1u << 20)) flush(); }
	void num(long v) { char b[24]; int l = 0; bool neg = v < 0; unsigned long x = neg? 0ul - (unsigned long)v : (unsigned long)v;
		do { b[l++] = '0' + x % 10; x /= 10; } while (x); if (neg) b[l++] = '-'; while (l) s.push_back(b[--l]); }
	void str(const char *p) { s.append(p); }
	void ch(char c) { s.push_back(c); }
};
}

static void write_seg_lines(const pg_graph_t *g, Out &o)   // format.c:120-135
{
	const pg_data_t *d = g->d;
	for (int32_t i = 0; i < g->n_seg; ++i) {
		const pg_seg_t *s = &g->seg[i];
		int32_t pid = d->gene[s->gid].rep_pid;
		o.str("S\t"); o.str(d->gene[s->gid].name); o.str("\t*\tLN:i:"); o.num(d->prot[pid].len);
		o.str("\tng:i:"); o.num(s->n_genome); o.str("\tnc:i:"); o.num(s->tot_cnt);
		o.str("\tc1:i:"); o.num(s->n_dom); o.str("\tc2:i:"); o.num(s->n_sub);
		o.str("\tpp:Z:"); o.str(d->prot[pid].name); o.ch('\n');
		o.maybe_flush();
	}
}

static void write_arc_lines(const pg_graph_t *g, Out &o)   // format.c:137-151
{
	const pg_data_t *d = g->d;
	for (int32_t i = 0; i < g->n_arc; ++i) {
		const pg_arc_t *a = &g->arc[i];
		uint32_t v = (uint32_t)(a->x >> 32), w = (uint32_t)a->x;
		o.str("L\t"); o.str(d->gene[g->seg[v >> 1].gid].name); o.ch('\t'); o.ch("+-"[v & 1]); o.ch('\t');
		o.str(d->gene[g->seg[w >> 1].gid].name); o.ch('\t'); o.ch("+-"[w & 1]); o.str("\t0M\tng:i:"); o.num(a->n_genome);
		o.str("\tnc:i:"); o.num(a->tot_cnt); o.str("\tad:i:"); o.num(a->avg_dist);
		o.str("\ts1:i:"); o.num(a->s1); o.str("\ts2:i:"); o.num(a->s2); o.ch('\n');
		o.maybe_flush();
	}
}

// "sample#hap#contig" (format.c:159-181): returns hap >= 0 and the sample in `buf`, or -1
static int32_t parse_sample(std::string &buf, const char *name)
{
	const char *h1 = strchr(name, '#');
	if (h1 == 0) return -1;
	if (h1 == name) return -1;
	buf.assign(name, h1 - name);
	const char *q = h1 + 1, *p = q;
	while (*p && *p != '#') ++p;
	char *r;
	long hap = strtol(q, &r, 10);
	return (r == p && hap >= 0)? (int32_t)hap : -1;
}

// Per-genome text produced by worker threads, written in genome order: the W lines and the BED lines of one genome depend
// on that genome alone (format.c:183-225, 78-118), and at 100 x 200 k hits formatting them single-threaded costs an order of
// magnitude more than the whole hot path.  Batches of genomes bound the memory held in strings.
template<class F> static void for_each_genome_text(int32_t n_genome, Out &o, F f)
{
	int32_t n_thr = (int32_t)std::thread::hardware_concurrency();
	if (const char *e = getenv("PGIO_THREADS")) n_thr = atoi(e);
	if (n_thr > n_genome) n_thr = n_genome;
	if (n_thr < 1) n_thr = 1;
	const int32_t batch = n_thr * 2;
	std::vector<std::string> text(batch);
	for (int32_t j0 = 0; j0 < n_genome; j0 += batch) {
		const int32_t nb = n_genome - j0 < batch? n_genome - j0 : batch;
		std::atomic<int32_t> next(0);
		auto worker = [&]() { for (int32_t k; (k = next.fetch_add(1)) < nb;) { text[k].clear(); f(j0 + k, text[k]); } };
		std::vector<std::thread> th;
		for (int32_t t = 1; t < n_thr && t < nb; ++t) th.emplace_back(worker);
		worker();
		for (auto &t : th) t.join();
		o.flush();
		for (int32_t k = 0; k < nb; ++k) if (!text[k].empty()) fwrite(text[k].data(), 1, text[k].size(), o.fp);
	}
}

static inline void put_num(std::string &s, long v)
{
	char b[24]; int l = 0; bool neg = v < 0; unsigned long x = neg? 0ul - (unsigned long)v : (unsigned long)v;
	do { b[l++] = '0' + x % 10; x /= 10; } while (x);
	if (neg) b[l++] = '-';
	while (l) s.push_back(b[--l]);
}

static void walk_lines_of_genome(const pg_data_t *d, int32_t j, std::string &s)   // format.c:183-225
{
	const pg_genome_t *g = &d->genome[j];
	std::string sample;
	// only unfiltered hits are printed (format.c:205,213): order THEM contig-major, then by CDS mid-point; ties keep the incoming
	// (cid,cs) order (see DESIGN.md "tie order") -- a stable sort of the kept hits equals the stable sort of all hits with the rest removed
	std::vector<int32_t> ord;
	for (int32_t i = 0; i < g->n_hit; ++i) if (!g->hit[i].flt) ord.push_back(i);
	std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) {
		const pg_hit_t &x = g->hit[a], &y = g->hit[b];
		return x.cid != y.cid? x.cid < y.cid : x.cm < y.cm; });
	const size_t n = ord.size();
	for (size_t i0 = 0, i = 1; i <= n; ++i) {
		if (i < n && g->hit[ord[i]].cid == g->hit[ord[i0]].cid) continue;
		const int32_t cid = g->hit[ord[i0]].cid;
		const int32_t hap = parse_sample(sample, g->ctg[cid].name);
		s += "W\t";
		if (hap >= 0) { s += sample; s.push_back('\t'); put_num(s, hap); }
		else if (g->label) { s += g->label; s += "\t0"; }
		else { put_num(s, j); s += "\t0"; }
		s.push_back('\t'); s += g->ctg[cid].name; s += "\t*\t*\t";
		for (size_t k = i0; k < i; ++k) {
			const pg_hit_t *a = &g->hit[ord[k]];
			s.push_back("><"[a->rev]); s += d->gene[d->prot[a->pid].gid].name;
		}
		s += "\tlf:B:i";
		for (size_t k = i0; k < i; ++k) { s.push_back(','); put_num(s, g->hit[ord[k]].lof); }
		s.push_back('\n');
		i0 = i;
	}
}

extern "C" int32_t pgio_write_gfa(const pg_graph_t *g, const char *path, int32_t flags)
{
	Out o(path);
	if (o.fp == 0) return -1;
	if (flags & 1) { write_seg_lines(g, o); write_arc_lines(g, o); }
	if (flags & 2) for_each_genome_text(g->d->n_genome, o, [&](int32_t j, std::string &s) { walk_lines_of_genome(g->d, j, s); });
	return 0;
}

// W lines of genome j from walk tokens (already contig-major, cm order inside a contig): format.c:183-225 without the two sorts
static void walk_lines_from_tokens(const pg_data_t *d, int32_t j, const pgx_walk_token_t *t, int64_t n, std::string &s)
{
	const pg_genome_t *g = &d->genome[j];
	std::string sample;
	for (int64_t i0 = 0, i = 1; i <= n; ++i) {
		if (i < n && t[i].cid == t[i0].cid) continue;
		const int32_t cid = t[i0].cid;
		const int32_t hap = parse_sample(sample, g->ctg[cid].name);
		s += "W\t";
		if (hap >= 0) { s += sample; s.push_back('\t'); put_num(s, hap); }
		else if (g->label) { s += g->label; s += "\t0"; }
		else { put_num(s, j); s += "\t0"; }
		s.push_back('\t'); s += g->ctg[cid].name; s += "\t*\t*\t";
		for (int64_t k = i0; k < i; ++k) { s.push_back("><"[t[k].gene_strand & 1]); s += d->gene[t[k].gene_strand >> 1].name; }
		s += "\tlf:B:i";
		for (int64_t k = i0; k < i; ++k) { s.push_back(','); put_num(s, t[k].lof); }
		s.push_back('\n');
		i0 = i;
	}
}

extern "C" int32_t pgio_write_gfa_tokens(const pg_graph_t *g, const pgx_walk_token_t *tok, const int64_t *genome_off, const char *path, int32_t flags)
{
	Out o(path);
	if (o.fp == 0) return -1;
	if (flags & 1) { write_seg_lines(g, o); write_arc_lines(g, o); }
	if (flags & 2) for_each_genome_text(g->d->n_genome, o, [&](int32_t j, std::string &s) { walk_lines_from_tokens(g->d, j, tok + genome_off[j], genome_off[j + 1] - genome_off[j], s); });
	return 0;
}

// gfa2matrix (pangene.js:1168-1246) without the GFA text: rows = segments in S-line order, columns = the "sample#hap" of every W line in
// order of first appearance, cell = number of walk tokens of that assembly whose gene is the row's segment
extern "C" int32_t pgio_matrix_from_tokens(const pg_graph_t *g, const pgx_walk_token_t *tok, const int64_t *genome_off, const char *out_path, int32_t copy_number)
{
	const pg_data_t *d = g->d;
	std::unordered_map<std::string, int32_t> asm_id;
	std::vector<std::string> asm_name;
	std::vector<int32_t> asm_of_run;                       // one entry per W line (genome, contig run)
	std::string sample, a;
	// pass 1: the assemblies, in W-line order
	for (int32_t j = 0; j < d->n_genome; ++j) {
		const pgx_walk_token_t *t = tok + genome_off[j];
		const int64_t n = genome_off[j + 1] - genome_off[j];
		for (int64_t i0 = 0, i = 1; i <= n; ++i) {
			if (i < n && t[i].cid == t[i0].cid) continue;
			const int32_t hap = parse_sample(sample, d->genome[j].ctg[t[i0].cid].name);
			a.clear();
			if (hap >= 0) { a = sample; a.push_back('#'); put_num(a, hap); }
			else if (d->genome[j].label) { a = d->genome[j].label; a += "#0"; }
			else { put_num(a, j); a += "#0"; }
			auto it = asm_id.find(a);
			int32_t ai;
			if (it == asm_id.end()) { ai = (int32_t)asm_name.size(); asm_id.emplace(a, ai); asm_name.push_back(a); } else ai = it->second;
			asm_of_run.push_back(ai);
			i0 = i;
		}
	}
	const size_t n_asm = asm_name.size();
	std::vector<int32_t> mat((size_t)g->n_seg * (n_asm > 0? n_asm : 1), 0);
	size_t run = 0;
	for (int32_t j = 0; j < d->n_genome; ++j) {
		const pgx_walk_token_t *t = tok + genome_off[j];
		const int64_t n = genome_off[j + 1] - genome_off[j];
		for (int64_t i0 = 0, i = 1; i <= n; ++i) {
			if (i < n && t[i].cid == t[i0].cid) continue;
			const int32_t ai = asm_of_run[run++];
			for (int64_t k = i0; k < i; ++k) {
				const int32_t sid = g->g2s[t[k].gene_strand >> 1];
				if (sid >= 0) ++mat[(size_t)sid * n_asm + ai];
			}
			i0 = i;
		}
	}
	Out o(out_path);
	if (o.fp == 0) return -1;
	o.str("Gene");
	for (auto &x : asm_name) { o.ch('\t'); o.str(x.c_str()); }
	o.ch('\n');
	for (int32_t i = 0; i < g->n_seg; ++i) {
		o.str(d->gene[g->seg[i].gid].name);
		for (size_t k = 0; k < n_asm; ++k) { const int32_t v = mat[(size_t)i * n_asm + k]; o.ch('\t'); o.num(copy_number? v : (v > 1? 1 : v)); }
		o.ch('\n'); o.maybe_flush();
	}
	return g->n_seg;
}

static void bed_lines_of_genome(const pg_data_t *d, int32_t j, int32_t is_walk, std::string &s)   // format.c:78-118
{
	const pg_genome_t *g = &d->genome[j];
	for (int32_t i = 0; i < g->n_hit; ++i) {
		const pg_hit_t *a = &g->hit[i];
		if (is_walk && a->flt) continue;
		char idbuf[16];
		// the reference's %ld formatter narrows through int (format.c:43-47); coordinates < 2^31 print identically
		s += g->ctg[a->cid].name; s.push_back('\t'); put_num(s, (int)a->cs); s.push_back('\t'); put_num(s, (int)a->ce); s.push_back('\t');
		s += d->prot[a->pid].name; s.push_back('\t'); put_num(s, a->score_ori); s.push_back('\t'); s.push_back("+-"[a->rev]); s.push_back('\t');
		put_num(s, (int)a->cs); s.push_back('\t'); put_num(s, (int)a->ce); s += "\t0\t"; put_num(s, a->n_exon); s.push_back('\t');
		for (int32_t k = 0; k < a->n_exon; ++k) { put_num(s, g->exon[a->off_exon + k].oe - g->exon[a->off_exon + k].os); s.push_back(','); }
		s.push_back('\t');
		for (int32_t k = 0; k < a->n_exon; ++k) { put_num(s, g->exon[a->off_exon + k].os); s.push_back(','); }
		snprintf(idbuf, 15, "%.4f", (double)a->mlen / a->blen);
		s += "\tft:i:"; put_num(s, a->flt);
		s += "\tpf:Z:"; put_num(s, a->pseudo); put_num(s, a->flt_iso_ov); put_num(s, a->flt_chain); put_num(s, a->flt_iso_sub_self);
		s += "\trk:i:"; put_num(s, a->rank); s += "\trp:i:"; put_num(s, a->rep); s += "\tsd:i:"; put_num(s, a->shadow);
		s += "\tvt:i:"; put_num(s, a->vtx); s += "\tbr:i:"; put_num(s, a->weak_br); s += "\tcm:i:"; put_num(s, (int)a->cm);
		s += "\tid:f:"; s += idbuf; s += "\tdm:Z:"; s += a->pid_dom0 < 0? "*" : d->prot[a->pid_dom0].name; s.push_back('\n');
	}
}

extern "C" int32_t pgio_write_bed(const pg_data_t *d, int32_t is_walk, const char *path)
{
	Out o(path);
	if (o.fp == 0) return -1;
	for_each_genome_text(d->n_genome, o, [&](int32_t j, std::string &s) { bed_lines_of_genome(d, j, is_walk, s); });
	return 0;
}

// ------------------------------------------------------------------------------------------------
// gfa2matrix (pangene.js:1168-1246, options -c only; -d/-p CD-HIT merging is not reproduced)
// ------------------------------------------------------------------------------------------------
extern "C" int32_t pgio_gfa2matrix(const char *gfa_path, const char *out_path, int32_t copy_number)
{
	std::string buf;
	if (!read_whole(gfa_path, buf)) return -1;
	std::unordered_map<std::string, int32_t> seg_id, asm_id;
	std::vector<std::string> seg_name, asm_name;
	struct Walk { int32_t asm_i; std::vector<int32_t> segs; };
	std::vector<Walk> walks;
	auto reg_seg = [&](const std::string &n) { auto it = seg_id.find(n); if (it != seg_id.end()) return it->second;
		int32_t k = (int32_t)seg_name.size(); seg_id.emplace(n, k); seg_name.push_back(n); return k; };
	for_each_line(buf, [&](char *s, size_t) {
		std::vector<char*> f;
		for (char *p = s, *q = s;; ++p) if (*p == '\t' || *p == 0) { int c = *p; *p = 0; f.push_back(q); q = p + 1; if (c == 0) break; }
		if (f.empty()) return;
		if (f[0][0] == 'S' && f.size() >= 2) reg_seg(f[1]);
		else if (f[0][0] == 'L' && f.size() >= 5) { reg_seg(f[1]); reg_seg(f[3]); }
		else if (f[0][0] == 'W' && f.size() >= 7) {
			std::string a = std::string(f[1]) + "#" + f[2];
			auto it = asm_id.find(a);
			int32_t ai;
			if (it == asm_id.end()) { ai = (int32_t)asm_name.size(); asm_id.emplace(a, ai); asm_name.push_back(a); } else ai = it->second;
			Walk w; w.asm_i = ai;
			for (const char *p = f[6]; *p;) {
				if (*p == '>' || *p == '<') {
					const char *q = p + 1;
					while (*q && *q != '>' && *q != '<' && !isspace((unsigned char)*q)) ++q;
					auto jt = seg_id.find(std::string(p + 1, q - p - 1));
					if (jt != seg_id.end()) w.segs.push_back(jt->second);
					p = q;
				} else ++p;
			}
			walks.push_back(std::move(w));
		}
	});
	std::vector<std::vector<int32_t>> mat(seg_name.size(), std::vector<int32_t>(asm_name.size(), 0));
	for (const Walk &w : walks) for (int32_t s : w.segs) ++mat[s][w.asm_i];
	Out o(out_path);
	if (o.fp == 0) return -1;
	o.str("Gene");
	for (auto &a : asm_name) { o.ch('\t'); o.str(a.c_str()); }
	o.ch('\n');
	for (size_t i = 0; i < seg_name.size(); ++i) {
		o.str(seg_name[i].c_str());
		for (size_t j = 0; j < asm_name.size(); ++j) { o.ch('\t'); o.num(copy_number? mat[i][j] : (mat[i][j] > 1? 1 : mat[i][j])); }
		o.ch('\n'); o.maybe_flush();
	}
	return (int32_t)seg_name.size();
}
