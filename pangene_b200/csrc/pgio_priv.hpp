// pgio_priv.hpp -- internals shared by pgio.cpp and pgio_synth.cpp (host only)
#pragma once
#include <cstdint>
#include <string_view>
#include <vector>
#include "pangene_abi.h"
#include "pgx_packed.h"

namespace pgio {

// name -> dense id in first-seen order; owns the strings (stands in for the reference's dict.c).  Open addressing, linear probing,
// the 64-bit hash kept next to the id: a look-up touches one cache line of the table and, on a hash match, the name itself.
struct NameTable {
	struct Slot { uint64_t hash; int32_t id; };          // id < 0: empty
	std::vector<Slot> slots;
	std::vector<char*> names;
	std::vector<uint32_t> lens;
	const char *intern(std::string_view s, int32_t *id, bool *absent);
	int32_t find(std::string_view s) const;
	NameTable() = default;
	NameTable(const NameTable&) = delete;
	~NameTable();
private:
	static uint64_t hash_of(std::string_view s);
	void grow();
};

struct DataPriv {           // hangs off pg_data_t::d_ctg
	NameTable ctg, gene, prot;
	std::vector<char*> arena;   // bulk name storage of synthetic data
	~DataPriv() { for (char *p : arena) free(p); }
};

int64_t cds_mid(const pg_hit_t *a, const pg_exon_t *e);
int64_t cigar_to_exons(const char *cg, bool rev, std::vector<pg_exon_t> &tmp, std::vector<pg_exon_t> &out, int32_t *n_fs);
int32_t adjusted_score(const pg_opt_t *opt, int32_t ms, int32_t mlen, int32_t blen, int32_t qs, int32_t qe, int32_t plen);
void push_hit(pg_genome_t *g, pg_hit_t &hit, const std::vector<pg_exon_t> &ex, int32_t cg_fs, int32_t n_fs, int32_t n_stop);
// the same hit as a device-ready record appended to pk[0..*n) (capacity *m); its exons go to g->exon as usual
void push_hit_packed(pg_genome_t *g, pgx_packed_hit_t *&pk, int32_t &n, int32_t &m, pg_hit_t &hit, const std::vector<pg_exon_t> &ex, int32_t cg_fs, int32_t n_fs, int32_t n_stop);

}
