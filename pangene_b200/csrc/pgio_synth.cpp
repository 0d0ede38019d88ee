// pgio_synth.cpp -- deterministic synthetic miniprot-PAF generator (SURVEY.md §8d "Synthetic inputs").
//
// One record model, two sinks: (a) straight into pg_data_t (what pgio_read_paf() would build from the
// text of the same records; used by bench.py at full size, threads over genomes) and (b) PAF text
// (so the unmodified reference binary can eat the same data in the parity tests).
// Eukaryote shape (BASELINE configs 2-4): genes every 40 kb on n_ctg contigs, 12 % single-exon, paralog
// families (65 % singletons / 23 % pairs / 12 % size 3-5, 60 % tandem), per-genome deletions, tandem
// duplications and inversions, isoforms dropping an internal exon with a start shift of 3*((5k) mod 7) bp,
// processed-pseudogene-like single-exon partial hits at random loci.
// Bacterial shape (config 5): single-exon genes, core/accessory cluster frequencies, local
// rearrangements, 1-30 contigs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include <algorithm>
#include "pgio.h"
#include "pgio_priv.hpp"

namespace {

struct Rng {
	uint64_t s;
	explicit Rng(uint64_t seed) : s(seed) {}
	uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
	double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
	int32_t range(int32_t lo, int32_t hi) { return lo + (int32_t)(next() % (uint64_t)(hi - lo + 1)); } // inclusive
	bool flip(double p) { return uni() < p; }
};
static inline uint64_t mix(uint64_t a, uint64_t b) { Rng r(a * 0x9E3779B97F4A7C15ull + b + 0x1234567ull); r.next(); return r.next(); }

struct Op { char op; int32_t len; };

struct Rec {                   // one PAF line
	int32_t pid, qlen, qs, qe, ctg; bool rev;
	int64_t cs, ce;
	int32_t mlen, blen, ms, fs, st;
	std::vector<Op> cg;
};

struct Pop {                   // population-level model, a pure function of (params, seed)
	pgio_synth_t p;
	int32_t n_prot;
	// eukaryote
	std::vector<int32_t> fam_of, strand, slot_of;        // per gene
	std::vector<std::vector<int32_t>> fam_members, fam_exon_aa;
	std::vector<std::vector<int32_t>> intron;            // per gene: intron bp between consecutive exons
	std::vector<int32_t> slot_gene, ctg_first_slot;       // layout
	std::vector<int64_t> ctg_len;
	std::vector<std::vector<int32_t>> prot_exons;        // per protein: kept exon indices
	std::vector<int32_t> prot_shift_aa, prot_len;
	double avg_fam;
	// bacterial
	std::vector<double> freq;
};

const int32_t SLOT = 40000, SLOT0 = 100000;

static void build_pop_euk(Pop &P)
{
	const pgio_synth_t &p = P.p;
	Rng r(mix(p.seed, 11));
	int32_t G = p.n_gene;
	P.fam_of.assign(G, -1); P.strand.resize(G); P.slot_of.resize(G); P.intron.resize(G);
	std::vector<std::vector<int32_t>> units;   // layout units: tandem blocks or single genes
	int32_t g = 0;
	while (g < G) {
		double u = r.uni();
		int32_t sz = u < 0.65? 1 : (u < 0.88? 2 : r.range(3, 5));
		if (g + sz > G) sz = G - g;
		int32_t f = (int32_t)P.fam_members.size();
		std::vector<int32_t> mem;
		for (int32_t k = 0; k < sz; ++k) mem.push_back(g + k), P.fam_of[g + k] = f;
		std::vector<int32_t> ex;
		int32_t ne = r.flip(0.12)? 1 : r.range(2, 14);
		for (int32_t k = 0; k < ne; ++k) ex.push_back(r.range(20, 200));
		P.fam_members.push_back(mem); P.fam_exon_aa.push_back(ex);
		bool tandem = sz > 1 && r.flip(0.6);
		if (tandem) units.push_back(mem);
		else for (int32_t m : mem) units.push_back(std::vector<int32_t>{m});
		for (int32_t m : mem) {
			P.strand[m] = r.flip(0.5);
			for (int32_t k = 0; k + 1 < ne; ++k) P.intron[m].push_back(r.range(80, 2500));
		}
		g += sz;
	}
	for (size_t i = units.size(); i > 1; --i) std::swap(units[i - 1], units[r.next() % i]);   // Fisher-Yates
	for (auto &u : units) for (int32_t m : u) { P.slot_of[m] = (int32_t)P.slot_gene.size(); P.slot_gene.push_back(m); }
	int32_t C = p.n_ctg;
	P.ctg_first_slot.resize(C + 1);
	for (int32_t c = 0; c <= C; ++c) P.ctg_first_slot[c] = (int32_t)((int64_t)c * G / C);
	P.ctg_len.resize(C);
	for (int32_t c = 0; c < C; ++c) P.ctg_len[c] = 2 * (int64_t)SLOT0 + (int64_t)(P.ctg_first_slot[c + 1] - P.ctg_first_slot[c]) * SLOT;
	// proteins
	int32_t K = p.iso_per_gene;
	P.n_prot = G * K;
	P.prot_exons.resize(P.n_prot); P.prot_shift_aa.assign(P.n_prot, 0); P.prot_len.assign(P.n_prot, 0);
	double famsum = 0;
	for (int32_t gg = 0; gg < G; ++gg) {
		const std::vector<int32_t> &ex = P.fam_exon_aa[P.fam_of[gg]];
		int32_t ne = (int32_t)ex.size();
		famsum += (double)P.fam_members[P.fam_of[gg]].size();
		for (int32_t k = 0; k < K; ++k) {
			int32_t pid = gg * K + k, drop = -1;
			if (k > 0 && ne >= 3) drop = 1 + (k - 1) % (ne - 2);
			for (int32_t e = 0; e < ne; ++e) if (e != drop) P.prot_exons[pid].push_back(e);
			int32_t sh = (5 * k) % 7;
			if (ex[P.prot_exons[pid][0]] < sh + 8) sh = 0;
			P.prot_shift_aa[pid] = sh;
			int32_t len = -sh;
			for (int32_t e : P.prot_exons[pid]) len += ex[e];
			P.prot_len[pid] = len;
		}
	}
	P.avg_fam = famsum / G;
}

static void build_pop_bact(Pop &P)
{
	const pgio_synth_t &p = P.p;
	Rng r(mix(p.seed, 13));
	int32_t G = p.n_gene;
	P.n_prot = G;
	P.strand.resize(G); P.prot_len.resize(G); P.freq.resize(G);
	for (int32_t g = 0; g < G; ++g) {
		P.strand[g] = r.flip(0.5);
		P.prot_len[g] = r.range(100, 600);
		if (r.flip(0.6)) P.freq[g] = 0.99 + 0.01 * r.uni();
		else { // Beta(0.3,0.3) by Joehnk's method
			for (;;) { double x = pow(r.uni(), 1.0 / 0.3), y = pow(r.uni(), 1.0 / 0.3); if (x + y <= 1.0 && x + y > 0) { P.freq[g] = x / (x + y); break; } }
		}
	}
}

static int32_t score_of(int32_t naa, double iden) { double f = 1.0 - 3.0 * (1.0 - iden); int32_t s = (int32_t)(5.2 * naa * f); return s > 1? s : 1; }

// records of one eukaryote genome, in file order (proteins in pid order, hits by descending score)
static void gen_genome_euk(const Pop &P, int32_t j, std::vector<Rec> &out)
{
	const pgio_synth_t &p = P.p;
	Rng r(mix(p.seed, 1000 + (uint64_t)j));
	int32_t G = p.n_gene, K = p.iso_per_gene, C = p.n_ctg;
	std::vector<uint8_t> present(G, 1), dup(G, 0), flip(G, 0);
	std::vector<int32_t> phys(G);   // layout slot -> physical slot
	for (int32_t s = 0; s < G; ++s) phys[s] = s;
	if (j > 0) {
		for (int32_t g = 0; g < G; ++g) { if (r.flip(0.01)) present[g] = 0; else if (r.flip(0.01)) dup[g] = 1; }
		int32_t n_inv = G / 2000 + 1;
		for (int32_t t = 0; t < n_inv; ++t) {
			int32_t c = r.range(0, C - 1), a0 = P.ctg_first_slot[c], n = P.ctg_first_slot[c + 1] - a0, len = r.range(3, 20);
			if (n < len + 2) continue;
			int32_t a = a0 + r.range(0, n - len);
			// invert layout slots [a, a+len): the gene at a+i moves to the physical place of a+len-1-i and flips
			std::vector<int32_t> ph(phys.begin() + a, phys.begin() + a + len);
			for (int32_t i = 0; i < len; ++i) phys[a + i] = ph[len - 1 - i], flip[P.slot_gene[a + i]] ^= 1;
		}
	} else for (int32_t g = 0; g < G; ++g) if (r.flip(0.01)) dup[g] = 1;
	auto locus = [&](int32_t g, int32_t *ctg, int64_t *pos0, int64_t *span, bool *rev) {
		int32_t ps = phys[P.slot_of[g]];
		int32_t c = (int32_t)(std::upper_bound(P.ctg_first_slot.begin(), P.ctg_first_slot.end(), ps) - P.ctg_first_slot.begin()) - 1;
		*ctg = c; *pos0 = SLOT0 + (int64_t)(ps - P.ctg_first_slot[c]) * SLOT;
		const std::vector<int32_t> &ex = P.fam_exon_aa[P.fam_of[g]];
		int64_t sp = 0;
		for (size_t e = 0; e < ex.size(); ++e) sp += 3 * ex[e] + (e + 1 < ex.size()? P.intron[g][e] : 0);
		*span = sp; *rev = (P.strand[g] ^ flip[g]) != 0;
	};
	auto hit_at = [&](int32_t pid, int32_t b, int64_t pos0, int64_t span, bool rev, int32_t ctg, double iden, Rec &rec) {
		const std::vector<int32_t> &ex = P.fam_exon_aa[P.fam_of[b]];
		const std::vector<int32_t> &keep = P.prot_exons[pid];
		std::vector<int64_t> T(ex.size());
		int64_t t = 0;
		for (size_t e = 0; e < ex.size(); ++e) { T[e] = t; t += 3 * ex[e] + (e + 1 < ex.size()? P.intron[b][e] : 0); }
		int32_t sh = P.prot_shift_aa[pid];
		int64_t ts = T[keep[0]] + 3 * sh, te = T[keep.back()] + 3 * ex[keep.back()];
		rec.cg.clear();
		rec.cg.push_back(Op{'M', ex[keep[0]] - sh});
		for (size_t i = 1; i < keep.size(); ++i) {
			int64_t gap = T[keep[i]] - (T[keep[i - 1]] + 3 * ex[keep[i - 1]]);
			rec.cg.push_back(Op{'N', (int32_t)gap});
			rec.cg.push_back(Op{'M', ex[keep[i]]});
		}
		rec.pid = pid; rec.qlen = P.prot_len[pid]; rec.qs = 0; rec.qe = rec.qlen; rec.ctg = ctg; rec.rev = rev;
		if (!rev) rec.cs = pos0 + ts, rec.ce = pos0 + te; else rec.cs = pos0 + span - te, rec.ce = pos0 + span - ts;
		rec.blen = 3 * rec.qlen; rec.mlen = (int32_t)(rec.blen * iden); rec.ms = score_of(rec.qlen, iden);
		rec.fs = r.flip(0.005); rec.st = r.flip(0.005);
	};
	double target = (double)p.hits_per_genome / P.n_prot, extra = target - P.avg_fam;
	if (extra < 0) extra = 0;
	std::vector<Rec> pr;
	for (int32_t g = 0; g < G; ++g) {
		const std::vector<int32_t> &mem = P.fam_members[P.fam_of[g]];
		for (int32_t k = 0; k < K; ++k) {
			int32_t pid = g * K + k;
			pr.clear();
			for (int32_t b : mem) {
				if (!present[b]) continue;
				int32_t ctg; int64_t pos0, span; bool rev;
				locus(b, &ctg, &pos0, &span, &rev);
				Rec rec;
				double iden = b == g? 0.985 + 0.015 * r.uni() : 0.90 + 0.085 * r.uni();
				hit_at(pid, b, pos0, span, rev, ctg, iden, rec);
				pr.push_back(rec);
				if (dup[b] && 2 * span + 1000 < SLOT) {
					iden = b == g? 0.985 + 0.015 * r.uni() : 0.90 + 0.085 * r.uni();
					hit_at(pid, b, pos0 + span + 1000, span, rev, ctg, iden, rec);
					pr.push_back(rec);
				}
			}
			int32_t nx = (int32_t)extra + (r.uni() < extra - (int32_t)extra);
			for (int32_t t = 0; t < nx; ++t) {
				Rec rec;
				int32_t plen = P.prot_len[pid], naa = (int32_t)(plen * (0.55 + 0.40 * r.uni()));
				if (naa < 10) naa = plen < 10? plen : 10;
				rec.pid = pid; rec.qlen = plen; rec.qs = r.range(0, plen - naa); rec.qe = rec.qs + naa;
				rec.ctg = r.range(0, C - 1); rec.rev = r.flip(0.5);
				rec.cs = 1000 + (int64_t)(r.uni() * (double)(P.ctg_len[rec.ctg] - 3 * naa - 2000));
				rec.ce = rec.cs + 3 * naa;
				double iden = 0.70 + 0.25 * r.uni();
				rec.blen = 3 * naa; rec.mlen = (int32_t)(rec.blen * iden); rec.ms = score_of(naa, iden);
				rec.fs = r.flip(0.005); rec.st = r.flip(0.005);
				rec.cg.assign(1, Op{'M', naa});
				pr.push_back(rec);
			}
			std::stable_sort(pr.begin(), pr.end(), [](const Rec &a, const Rec &b) { return a.ms > b.ms; });
			for (Rec &x : pr) out.push_back(std::move(x));
		}
	}
}

static void gen_genome_bact(const Pop &P, int32_t j, std::vector<Rec> &out, std::vector<int64_t> &ctg_len)
{
	const pgio_synth_t &p = P.p;
	Rng r(mix(p.seed, 2000 + (uint64_t)j));
	int32_t G = p.n_gene;
	std::vector<int32_t> order;
	std::vector<uint8_t> flip(G, 0);
	for (int32_t g = 0; g < G; ++g) if (j == 0 || r.uni() < P.freq[g]) order.push_back(g);
	int32_t n = (int32_t)order.size();
	int32_t n_re = j == 0? 0 : r.range(0, 6);
	for (int32_t t = 0; t < n_re && n > 60; ++t) {
		int32_t len = r.range(5, 50), a = r.range(0, n - len);
		std::reverse(order.begin() + a, order.begin() + a + len);
		for (int32_t i = a; i < a + len; ++i) flip[order[i]] ^= 1;
	}
	double u = r.uni();
	int32_t C = 1 + (int32_t)((p.n_ctg - 1) * u * u * u);
	if (C > n) C = n > 0? n : 1;
	std::vector<int32_t> brk;   // first gene index of each contig
	brk.push_back(0);
	for (int32_t c = 1; c < C; ++c) brk.push_back(r.range(1, n - 1));
	std::sort(brk.begin(), brk.end());
	brk.erase(std::unique(brk.begin(), brk.end()), brk.end());
	C = (int32_t)brk.size();
	brk.push_back(n);
	ctg_len.assign(C, 0);
	std::vector<Rec> recs(G);          // own-locus record per cluster (pid == cluster id)
	std::vector<uint8_t> has(G, 0);
	for (int32_t c = 0; c < C; ++c) {
		int64_t pos = 500;
		for (int32_t i = brk[c]; i < brk[c + 1]; ++i) {
			int32_t g = order[i], naa = P.prot_len[g];
			Rec &rec = recs[g];
			double iden = 0.90 + 0.10 * r.uni();
			rec.pid = g; rec.qlen = naa; rec.qs = 0; rec.qe = naa; rec.ctg = c; rec.rev = (P.strand[g] ^ flip[g]) != 0;
			rec.cs = pos; rec.ce = pos + 3 * naa; rec.blen = 3 * naa; rec.mlen = (int32_t)(rec.blen * iden); rec.ms = score_of(naa, iden);
			rec.fs = r.flip(0.005); rec.st = r.flip(0.005);
			rec.cg.assign(1, Op{'M', naa});
			has[g] = 1;
			pos = rec.ce + r.range(50, 300);
		}
		ctg_len[c] = pos + 500;
	}
	for (int32_t g = 0; g < G; ++g) {   // file order = protein order; extras are partial hits at random places
		std::vector<Rec> pr;
		if (has[g]) pr.push_back(recs[g]);
		// (the rate of these stray hits shrinks with the population: every one of them makes two adjacencies nobody else has, and at 4 % per
		// gene and genome a 1000-genome set gives every gene ~80 one-off neighbours -- all vertices go over the degree cap and the graph is empty)
		if (r.flip(0.04 * (p.n_genome_total > 25? 25.0 / p.n_genome_total : 1.0)) && n > 0) {
			Rec rec;
			int32_t plen = P.prot_len[g], naa = (int32_t)(plen * (0.55 + 0.40 * r.uni()));
			rec.pid = g; rec.qlen = plen; rec.qs = r.range(0, plen - naa); rec.qe = rec.qs + naa;
			rec.ctg = r.range(0, C - 1); rec.rev = r.flip(0.5);
			int64_t room = ctg_len[rec.ctg] - 3 * naa - 20;
			rec.cs = 10 + (room > 0? (int64_t)(r.uni() * (double)room) : 0);
			rec.ce = rec.cs + 3 * naa;
			if (rec.ce > ctg_len[rec.ctg]) ctg_len[rec.ctg] = rec.ce + 10;
			double iden = 0.70 + 0.25 * r.uni();
			rec.blen = 3 * naa; rec.mlen = (int32_t)(rec.blen * iden); rec.ms = score_of(naa, iden);
			rec.fs = rec.st = 0;
			rec.cg.assign(1, Op{'M', naa});
			pr.push_back(rec);
		}
		std::stable_sort(pr.begin(), pr.end(), [](const Rec &a, const Rec &b) { return a.ms > b.ms; });
		for (Rec &x : pr) out.push_back(std::move(x));
	}
}

static void build_pop(Pop &P, const pgio_synth_t *p)
{
	P.p = *p;
	if (p->shape == 0) build_pop_euk(P); else build_pop_bact(P);
}

static void gen_genome(const Pop &P, int32_t j, std::vector<Rec> &recs, std::vector<int64_t> &ctg_len)
{
	recs.clear();
	if (P.p.shape == 0) { gen_genome_euk(P, j, recs); ctg_len = P.ctg_len; }
	else gen_genome_bact(P, j, recs, ctg_len);
}

static void names_of(const Pop &P, int32_t pid, char *gene, char *prot)
{
	int32_t K = P.p.shape == 0? P.p.iso_per_gene : 1, g = pid / K, k = pid % K;
	snprintf(gene, 32, "G%06d", g);
	snprintf(prot, 48, "G%06d:P%06d.%d", g, g, k + 1);
}

static void ctg_name(const Pop &P, int32_t j, int32_t c, char *buf) { snprintf(buf, 48, "S%04d#0#c%d", j, c + 1); (void)P; }

// exon list straight from the op list (same arithmetic as pgio::cigar_to_exons on the text form)
static int64_t ops_to_exons(const std::vector<Op> &cg, bool rev, std::vector<pg_exon_t> &tmp, std::vector<pg_exon_t> &out)
{
	int64_t x = 0;
	tmp.clear(); tmp.push_back(pg_exon_t{0, 0});
	for (const Op &o : cg) {
		if (o.op == 'N') { tmp.back().oe = (int32_t)x; x += o.len; tmp.push_back(pg_exon_t{(int32_t)x, (int32_t)x}); }
		else x += 3 * (int64_t)o.len;
	}
	tmp.back().oe = (int32_t)x;
	out.clear();
	if (!rev) out = tmp;
	else for (int32_t i = (int32_t)tmp.size() - 1; i >= 0; --i) out.push_back(pg_exon_t{(int32_t)(x - tmp[i].oe), (int32_t)(x - tmp[i].os)});
	return x;
}

} // namespace

extern "C" void pgio_synth_default(pgio_synth_t *p, int32_t config)
{
	memset(p, 0, sizeof(*p));
	p->seed = (uint64_t)config;
	if (config == 5) { p->shape = 1; p->n_genome = p->n_genome_total = 1000; p->n_gene = 6000; p->iso_per_gene = 1; p->n_ctg = 30; p->hits_per_genome = 5000; }
	else if (config == 3 || config == 4) { p->shape = 0; p->n_genome = p->n_genome_total = 100; p->n_gene = 20000; p->iso_per_gene = 4; p->n_ctg = 24; p->hits_per_genome = 200000; p->seed = 3; }
	else { p->shape = 0; p->n_genome = p->n_genome_total = 10; p->n_gene = 20000; p->iso_per_gene = 1; p->n_ctg = 24; p->hits_per_genome = 50000; p->seed = 2; }
}

extern "C" pg_data_t *pgio_synth(const pg_opt_t *opt, const pgio_synth_t *p)
{
	Pop P;
	build_pop(P, p);
	pg_data_t *d = pgio_data_init();
	pgio::DataPriv *pv = (pgio::DataPriv*)d->d_ctg;
	int32_t K = p->shape == 0? p->iso_per_gene : 1;
	// global tables: ids are the generator's own indices (== first-seen order of genome 0's text, where every
	// protein appears in pid order)
	d->n_gene = d->m_gene = p->n_gene; d->n_prot = d->m_prot = P.n_prot;
	d->gene = (pg_gene_t*)calloc(d->n_gene, sizeof(pg_gene_t));
	d->prot = (pg_prot_t*)calloc(d->n_prot, sizeof(pg_prot_t));
	{
		char *gn = (char*)malloc((size_t)d->n_gene * 8), *pn = (char*)malloc((size_t)d->n_prot * 20);
		pv->arena.push_back(gn); pv->arena.push_back(pn);
		char a[32], b[48];
		for (int32_t pid = 0; pid < d->n_prot; ++pid) {
			names_of(P, pid, a, b);
			int32_t g = pid / K;
			if (pid % K == 0) { memcpy(gn + (size_t)g * 8, a, 8); d->gene[g].name = gn + (size_t)g * 8; d->gene[g].rep_pid = 0; }
			memcpy(pn + (size_t)pid * 20, b, strlen(b) + 1); d->prot[pid].name = pn + (size_t)pid * 20;
			d->prot[pid].gid = g; d->prot[pid].len = P.prot_len[pid];
			if ((int32_t)d->gene[g].len < P.prot_len[pid]) d->gene[g].len = P.prot_len[pid];
		}
	}
	d->n_genome = d->m_genome = p->n_genome;
	d->genome = (pg_genome_t*)calloc(p->n_genome > 0? p->n_genome : 1, sizeof(pg_genome_t));
	std::vector<std::vector<char*>> arenas(p->n_genome);
	std::atomic<int32_t> next(0);
	auto worker = [&]() {
		std::vector<Rec> recs;
		std::vector<int64_t> clen;
		std::vector<pg_exon_t> tmp, ex;
		std::vector<int32_t> cmap, cnt;
		for (;;) {
			int32_t jl = next.fetch_add(1);
			if (jl >= p->n_genome) break;
			int32_t j = p->genome_begin + jl;
			pg_genome_t *g = &d->genome[jl];
			gen_genome(P, j, recs, clen);
			char lab[32];
			snprintf(lab, 32, "S%04d", j);
			g->label = strdup(lab);
			cmap.assign(clen.size(), -1);
			cnt.assign(P.n_prot, -1);
			for (const Rec &rc : recs) {   // same order of effects as one reader line (read.c:139-236)
				pg_hit_t hit;
				memset(&hit, 0, sizeof(hit));
				hit.pid = rc.pid; hit.pid_dom = -1; hit.cid = hit.off_exon = hit.n_exon = -1;
				hit.rank = ++cnt[rc.pid];
				hit.qs = rc.qs; hit.qe = rc.qe;
				if (hit.qe - hit.qs < rc.qlen * opt->min_prot_ratio) continue;
				hit.rev = rc.rev;
				if (cmap[rc.ctg] < 0) {
					char *nm = (char*)malloc(48);
					ctg_name(P, j, rc.ctg, nm);
					arenas[jl].push_back(nm);
					if (g->n_ctg >= g->m_ctg) { g->m_ctg = g->m_ctg? g->m_ctg * 2 : 32; g->ctg = (pg_ctg_t*)realloc(g->ctg, sizeof(pg_ctg_t) * g->m_ctg); }
					g->ctg[g->n_ctg].name = nm; g->ctg[g->n_ctg].len = clen[rc.ctg];
					cmap[rc.ctg] = g->n_ctg++;
				}
				hit.cid = cmap[rc.ctg];
				hit.cs = rc.cs; hit.ce = rc.ce; hit.mlen = rc.mlen; hit.blen = rc.blen;
				if (hit.mlen < hit.blen * opt->min_prot_iden) continue;
				hit.score_ori = rc.ms;
				hit.score_adj = pgio::adjusted_score(opt, rc.ms, rc.mlen, rc.blen, rc.qs, rc.qe, rc.qlen);
				int64_t x = ops_to_exons(rc.cg, rc.rev, tmp, ex);
				if (x != hit.ce - hit.cs) { fprintf(stderr, "[pgio_synth] internal error: cigar/coordinate mismatch\n"); abort(); }
				pgio::push_hit(g, hit, ex, 0, rc.fs, rc.st);
			}
		}
	};
	int32_t nt = p->n_threads > 0? p->n_threads : (int32_t)std::thread::hardware_concurrency();
	if (nt < 1) nt = 1;
	if (nt > p->n_genome) nt = p->n_genome > 0? p->n_genome : 1;
	std::vector<std::thread> th;
	for (int32_t t = 0; t < nt; ++t) th.emplace_back(worker);
	for (auto &t : th) t.join();
	for (auto &a : arenas) for (char *s : a) pv->arena.push_back(s);
	return d;
}

extern "C" int32_t pgio_synth_write_paf(const pg_opt_t *opt, const pgio_synth_t *p, int32_t j, const char *path)
{
	(void)opt;
	Pop P;
	build_pop(P, p);
	std::vector<Rec> recs;
	std::vector<int64_t> clen;
	gen_genome(P, j, recs, clen);
	FILE *fp = fopen(path, "w");
	if (fp == 0) return -1;
	char a[32], b[48], cn[48];
	std::string cg;
	for (const Rec &rc : recs) {
		names_of(P, rc.pid, a, b);
		ctg_name(P, j, rc.ctg, cn);
		cg.clear();
		for (const Op &o : rc.cg) { cg += std::to_string(o.len); cg.push_back(o.op); }
		fprintf(fp, "%s\t%d\t%d\t%d\t%c\t%s\t%ld\t%ld\t%ld\t%d\t%d\t0\tAS:i:%d\tms:i:%d\tnp:i:%d\tfs:i:%d\tst:i:%d\tda:i:0\tdo:i:0\tcg:Z:%s\n",
			b, rc.qlen, rc.qs, rc.qe, "+-"[rc.rev], cn, (long)clen[rc.ctg], (long)rc.cs, (long)rc.ce, rc.mlen, rc.blen,
			rc.ms, rc.ms, rc.qe - rc.qs, rc.fs, rc.st, cg.c_str());
	}
	fclose(fp);
	return (int32_t)recs.size();
}
