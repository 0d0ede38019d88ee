/*
 * pangene_abi.h -- the binary contract between a pangene host program and the hot-path library.
 *
 * The reference has no plugin/FFI mechanism: its hot path sits behind four plain C functions and a
 * handful of POD structs declared in /root/reference/pangene.h:19-141.  A drop-in replacement must
 * therefore agree with those structs BYTE FOR BYTE (field order, widths and bit-field packing as
 * laid out by gcc on x86-64).  This header declares them under the reference's own names so that a
 * host compiled against either header can call either library.  Every layout is pinned by the
 * _Static_assert block at the end (sizes/offsets measured on the reference build).
 *
 * Nothing here is CUDA- or torch-typed: plain pointers, int32/int64, doubles.
 */
#ifndef PANGENE_ABI_H
#define PANGENE_ABI_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* option bits of pg_opt_t::flag (reference pangene.h:8-17) */
#define PG_F_WRITE_BED_RAW      0x1
#define PG_F_WRITE_BED_WALK     0x2
#define PG_F_WRITE_BED_FLAG     0x4
#define PG_F_WRITE_NO_WALK      0x8
#define PG_F_WRITE_VTX_SEL      0x10
#define PG_F_FRAG_MODE          0x20   /* -F: contigs do not separate "local" genes (branch.c:39) */
#define PG_F_NO_JOINT_PSEUDO    0x40   /* -J (graph.c:14) */
#define PG_F_ORI_FOR_BRANCH     0x80   /* --ori-sc (graph.c:89) */
#define PG_F_CHECK_STRAND       0x100  /* -S (overlap.c:78,128) */
#define PG_F_DROP_SGL_EXON      0x200  /* -E (hit.c:180) */

/* 16-byte sort record used by the reference's radix sorts (pangene.h:19-21) */
typedef struct { uint64_t x, y; } pg128_t;

/* run-time options (pangene.h:23-42); 128 bytes */
typedef struct {
	uint32_t flag;
	int32_t  gene_delim;
	double   min_prot_ratio;    /* reader: drop hit if aligned fraction below this */
	double   min_prot_iden;     /* reader: drop hit if identity below this */
	double   score_adj_coef;    /* reader: score_adj = ms*exp(-coef*(div+uncov)) */
	double   min_ov_ratio;      /* -f: different-gene overlap needs >= this share of the shorter CDS */
	double   min_vertex_ratio;  /* -p */
	double   branch_diff;       /* -b */
	double   branch_diff_dist;  /* -y */
	double   branch_diff_cut;   /* -B */
	int32_t  max_avg_occ;       /* -c */
	int32_t  max_degree;        /* -g */
	int32_t  max_dist_loci;     /* -r */
	int32_t  n_branch_flt;      /* -T */
	int32_t  min_arc_cnt;       /* -a */
	int32_t  local_dist;        /* -D */
	int32_t  local_count;       /* -C */
	void    *excl, *incl, *preferred;   /* reader-side name sets; opaque to the hot path */
} pg_opt_t;

/* one coding exon, offsets relative to the owning hit's cs (pangene.h:44-46) */
typedef struct { int32_t os, oe; } pg_exon_t;

/* protein table entry (pangene.h:48-53); 32 bytes */
typedef struct {
	const char *name;
	int32_t len, gid;
	int32_t rep;
	int32_t n, avg_score_adj, max_score_ori;
} pg_prot_t;

/* gene table entry (pangene.h:55-59); 16 bytes */
typedef struct {
	const char *name;
	uint32_t len:30, preferred:1, included:1;
	int32_t rep_pid;
} pg_gene_t;

/* one protein-to-genome alignment (pangene.h:61-72); 88 bytes.
 * bit-field word at offset 60: rev=bit0 flt=1 flt_iso_sub_self=2 flt_iso_ov=3 flt_chain=4 pseudo=5
 * vtx=6 shadow=7 rep=8 weak_br=9..10 */
typedef struct {
	int32_t pid;
	int32_t qs, qe;
	int32_t cid;
	int32_t mlen, blen, lof;
	int32_t rank;
	int32_t score_ori, score_adj, score_dom;
	int32_t n_exon, off_exon;
	int32_t pid_dom, pid_dom0;
	uint32_t rev:1, flt:1, flt_iso_sub_self:1, flt_iso_ov:1, flt_chain:1, pseudo:1, vtx:1, shadow:1, rep:1, weak_br:2;
	int64_t cs, cm, ce;
} pg_hit_t;

#define PG_HB_REV       0x001u
#define PG_HB_FLT       0x002u
#define PG_HB_FLT_SUB   0x004u
#define PG_HB_FLT_OV    0x008u
#define PG_HB_FLT_CHAIN 0x010u
#define PG_HB_PSEUDO    0x020u
#define PG_HB_VTX       0x040u
#define PG_HB_SHADOW    0x080u
#define PG_HB_REP       0x100u
#define PG_HB_WBR_SHIFT 9
#define PG_HB_WBR_MASK  0x600u
#define PG_HIT_BITS_OFFSET 60   /* byte offset of the bit-field word inside pg_hit_t */

typedef struct { const char *name; int64_t len; } pg_ctg_t;          /* pangene.h:74-77 */

typedef struct {                                                       /* pangene.h:79-87 */
	int32_t n_ctg, m_ctg;   pg_ctg_t  *ctg;
	int32_t n_hit, m_hit;   pg_hit_t  *hit;
	int32_t n_exon, m_exon; pg_exon_t *exon;
	char *label;
} pg_genome_t;

typedef struct {                                                       /* pangene.h:89-97 */
	void *d_ctg, *d_gene, *d_prot;
	int32_t n_genome, m_genome; pg_genome_t *genome;
	int32_t n_gene, m_gene;     pg_gene_t   *gene;
	int32_t n_prot, m_prot;     pg_prot_t   *prot;
} pg_data_t;

typedef struct {                                                       /* pangene.h:99-105; 32 bytes */
	int32_t gid, n_dom, n_sub;
	int32_t n_genome;
	int32_t tot_cnt;
	uint32_t del:1, dummy:31;
	int32_t n_dist_loci[2];
} pg_seg_t;

typedef struct {                                                       /* pangene.h:107-114; 32 bytes */
	uint64_t x;             /* v<<32|w with v = sid<<1|strand */
	int32_t n_genome;
	int32_t tot_cnt;
	int32_t avg_dist;
	int32_t s1, s2;
	uint32_t del:1, weak_br:2, dummy:29;
} pg_arc_t;

typedef struct {                                                       /* pangene.h:116-124 */
	pg_data_t *d;
	int32_t *g2s;
	int32_t n_seg, m_seg; pg_seg_t *seg;
	int32_t n_arc, m_arc; pg_arc_t *arc;
	uint64_t *idx;          /* per oriented vertex: first_arc<<32 | n_arcs */
} pg_graph_t;

#ifdef __cplusplus
}
#define PG_ABI_ASSERT(c, m) static_assert(c, m)
#else
#define PG_ABI_ASSERT(c, m) _Static_assert(c, m)
#endif

PG_ABI_ASSERT(sizeof(pg_opt_t) == 128, "pg_opt_t layout");
PG_ABI_ASSERT(sizeof(pg_prot_t) == 32 && sizeof(pg_gene_t) == 16, "prot/gene layout");
PG_ABI_ASSERT(sizeof(pg_hit_t) == 88 && offsetof(pg_hit_t, cs) == 64 && offsetof(pg_hit_t, pid_dom0) == 56, "pg_hit_t layout");
PG_ABI_ASSERT(sizeof(pg_seg_t) == 32 && sizeof(pg_arc_t) == 32 && sizeof(pg_exon_t) == 8, "seg/arc/exon layout");
PG_ABI_ASSERT(sizeof(pg_genome_t) == 56 && sizeof(pg_data_t) == 72 && sizeof(pg_graph_t) == 56, "container layout");

#endif
