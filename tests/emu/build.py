"""TEST INFRASTRUCTURE: build tests/emu/_build/libsweep_emu.so -- the sweep kernels of pangene_b200/csrc/pgx_kernels.cuh compiled
for the host on top of simt_emu.hpp.  The kernel text is taken verbatim from the product header (between its emu markers)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.environ.get("PGX_EMU_SRC", os.path.join(ROOT, "pangene_b200", "csrc", "pgx_kernels.cuh"))   # (override: mutation testing)
OUT = os.environ.get("PGX_EMU_OUT", os.path.join(HERE, "_build"))

DRIVER = r'''
struct EmuCols { int32_t n; const int32_t *ctg, *cs, *ce, *pid, *gidp, *rank, *cds, *sori, *pmax; const uint64_t *key; const uint32_t *bits, *eoff; const int32_t *ex; };
extern "C" int sweep_emu(int variant, int mode, const EmuCols *e, double min_ov_ratio, int check_strand, uint32_t *bits_out, int32_t *pdom_out, int32_t *sdom_out, uint32_t *tie_cnt)
{
	using namespace pgx;
	SweepCols c; c.n = e->n; c.ctg = e->ctg; c.cs = e->cs; c.ce = e->ce; c.pid = e->pid; c.gidp = e->gidp; c.rank = e->rank; c.cds = e->cds; c.sori = e->sori; c.pmax = e->pmax;
	c.key = e->key; c.bits = e->bits; c.eoff = e->eoff; c.ex = (const int2*)e->ex;
	SweepOpt o; o.min_ov_ratio = min_ov_ratio; o.check_strand = check_strand;
	if (c.n <= 0) return 0;
#define RUN(K, T, ...) emu::launch((c.n + (T) - 1) / (T), (T), [&] { K(c, o, bits_out, pdom_out, sdom_out, tie_cnt __VA_ARGS__); })
#define COMMA ,
	const uint32_t *no_dirty = nullptr;
	if (variant == 0) {                                        // plain thread-per-hit kernel
		if (mode == 0) RUN(k_sweep<SW_LOAD>, 128); else if (mode == 1) RUN(k_sweep<SW_SHADOW>, 128); else RUN(k_sweep<SW_ISO>, 128);
	} else if (variant == 1) {                                 // production kernel, VAR 0 (no tile skipping)
		if (mode == 0) RUN((k_sweep_tiled<SW_LOAD COMMA 0>), SWQ_TILE, COMMA no_dirty); else if (mode == 1) RUN((k_sweep_tiled<SW_SHADOW COMMA 0>), SWQ_TILE, COMMA no_dirty); else RUN((k_sweep_tiled<SW_ISO COMMA 0>), SWQ_TILE, COMMA no_dirty);
	} else if (variant == 2) {                                 // VAR 1 (SW_SHADOW has one variant only)
		if (mode == 0) RUN((k_sweep_tiled<SW_LOAD COMMA 1>), SWQ_TILE, COMMA no_dirty); else if (mode == 1) RUN((k_sweep_tiled<SW_SHADOW COMMA 0>), SWQ_TILE, COMMA no_dirty); else RUN((k_sweep_tiled<SW_ISO COMMA 1>), SWQ_TILE, COMMA no_dirty);
	} else if (variant == 3) {                                 // VAR 2: as 1, isoform sweep on the cluster shortcut
		if (mode == 0) RUN((k_sweep_tiled<SW_LOAD COMMA 1>), SWQ_TILE, COMMA no_dirty); else if (mode == 1) RUN((k_sweep_tiled<SW_SHADOW COMMA 0>), SWQ_TILE, COMMA no_dirty); else RUN((k_sweep_tiled<SW_ISO COMMA 2>), SWQ_TILE, COMMA no_dirty);
	} else if (variant == 4) {                                 // VAR 3: as 1, no cluster pass in the first sweep
		if (mode == 0) RUN((k_sweep_tiled<SW_LOAD COMMA 3>), SWQ_TILE, COMMA no_dirty); else if (mode == 1) RUN((k_sweep_tiled<SW_SHADOW COMMA 0>), SWQ_TILE, COMMA no_dirty); else RUN((k_sweep_tiled<SW_ISO COMMA 1>), SWQ_TILE, COMMA no_dirty);
	}
	else return -1;
	return 0;
}
// related_threshold() as compiled from the product header, for a batch of (shorter CDS, ratio) pairs
extern "C" void sweep_emu_thresholds(const int32_t *shorter, const double *ratio, int64_t n, int32_t *out)
{
	for (int64_t i = 0; i < n; ++i) out[i] = pgx::related_threshold(false, shorter[i], shorter[i] + 7, ratio[i]);
}
'''


def build():
    text = open(SRC).read()
    a = text.index("// [emu-begin sweep]")
    b = text.index("// [emu-end sweep]")
    body = text[a:b]
    n_dyn = body.count("extern __shared__ __align__(16) unsigned char smem_raw[];")
    assert n_dyn >= 1
    body = body.replace("extern __shared__ __align__(16) unsigned char smem_raw[];", "unsigned char *smem_raw = emu_dyn_smem;")
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "sweep_emu.cpp")
    src = ('#include "simt_emu.hpp"\n#include "pangene_abi.h"\nnamespace pgx {\nconstexpr uint32_t HB_FIRST = 0x80000000u;\n'
           + body + "\n}\n" + DRIVER)
    so = os.path.join(OUT, "libsweep_emu.so")
    if os.path.exists(cpp) and os.path.exists(so) and open(cpp).read() == src and os.path.getmtime(so) >= os.path.getmtime(os.path.join(HERE, "simt_emu.hpp")):
        return so
    open(cpp, "w").write(src)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-I", HERE, "-I", os.path.join(ROOT, "include"),
                           cpp, "-o", so])
    return so


PRIM_DRIVER = r'''
// host drivers in the shape of radix_sort_pairs() / exclusive_sum_u32() / segmented_running_max() (pgx_sort.cuh, pgx_scan.cuh)
static void emu_scan(int variant, const uint32_t *in, uint32_t *out, int64_t n_in)
{
	using namespace pgx;
	const int64_t n_out = n_in + 1;
	const int tiles = (int)((n_out + SCAN_TILE - 1) / SCAN_TILE);
	if (variant == 2) { emu::launch(1, SCAN_SMALL_B, [&] { k_scan_sum_small(in, out, (int32_t)n_in); }); return; }
	if (variant == 1) {
		std::vector<unsigned long long> status(tiles + 1, 0ull);
		uint32_t ticket = 0;
		emu::launch(tiles, SCAN_B, [&] { k_scan_sum_lookback(in, out, n_in, status.data(), &ticket); });
		return;
	}
	std::vector<uint32_t> agg(tiles + 1, 0u);
	emu::launch(tiles, SCAN_B, [&] { k_scan_sum_tiles(in, out, agg.data(), n_in); });
	if (tiles > 1) {
		emu::launch(1, SCAN_B, [&] { k_scan_sum_agg(agg.data(), tiles); });
		emu::launch((int)((n_out + 255) / 256), 256, [&] { k_scan_sum_add(out, agg.data(), n_out); });
	}
}
extern "C" int prim_emu_scan(int variant, const uint32_t *in, uint32_t *out, int64_t n_in) { if (n_in < 0) return -1; emu_scan(variant, in, out, n_in); return 0; }

extern "C" int prim_emu_sort(uint64_t *keys, uint32_t *vals, int64_t n, int n_bits, int scan_variant)
{
	using namespace pgx;
	if (n <= 0) return 0;
	std::vector<uint64_t> kalt(n + 1);
	std::vector<uint32_t> valt(n + 1);
	uint64_t *kin = keys, *kout = kalt.data();
	uint32_t *vin = vals, *vout = valt.data();
	const int n_tiles = (int)((n + RS_TILE - 1) / RS_TILE);
	const int64_t hn = 256ll * n_tiles;
	std::vector<uint32_t> hist(hn + 2);
	bool first = true;
	if (n_bits < 1) n_bits = 1;
	for (int shift = 0; shift < n_bits; shift += 8) {
		emu::launch(n_tiles, RS_BLOCK, [&] { k_rs_hist(kin, n, shift, hist.data(), n_tiles); });
		emu_scan(scan_variant, hist.data(), hist.data(), hn);
		const uint32_t *vsrc = first? (const uint32_t*)nullptr : vin;
		emu::launch(n_tiles, RS_BLOCK, [&] { k_rs_scatter(kin, vsrc, n, shift, hist.data(), n_tiles, kout, vout); });
		std::swap(kin, kout); std::swap(vin, vout);
		first = false;
	}
	if (kin != keys) { memcpy(keys, kin, (size_t)n * 8); memcpy(vals, vin, (size_t)n * 4); }
	return 0;
}

extern "C" int prim_emu_segmax(const int32_t *seg, const int32_t *val, int32_t *out, int64_t n)
{
	using namespace pgx;
	if (n <= 0) return 0;
	const int tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
	std::vector<SegVal> agg(tiles + 2);
	emu::launch(tiles, SCAN_B, [&] { k_segmax_tiles(seg, val, out, agg.data(), n); });
	if (tiles > 1) {
		emu::launch(1, SCAN_B, [&] { k_segmax_agg(agg.data(), tiles); });
		emu::launch((int)((n + 255) / 256), 256, [&] { k_segmax_apply(seg, out, agg.data(), n); });
	}
	return 0;
}
'''


def _between(text, name):
    a = text.index("// [emu-begin %s]" % name)
    b = text.index("// [emu-end %s]" % name)
    return text[a:b]


def build_prim():
    """tests/emu/_build/libprim_emu.so: the radix-sort, scan and segmented-max kernels of pgx_sort.cuh / pgx_scan.cuh (text between
    their emu markers, unchanged) on the SIMT shim, with host drivers that mirror the launch sequences of the product."""
    csrc = os.path.join(ROOT, "pangene_b200", "csrc")
    scan = open(os.path.join(csrc, "pgx_scan.cuh")).read()
    sort = open(os.path.join(csrc, "pgx_sort.cuh")).read()
    body = _between(scan, "scan") + "\n" + _between(scan, "segmax") + "\n" + _between(sort, "sort")
    assert body.count("extern __shared__ __align__(16) unsigned char rs_smem[];") == 1
    body = body.replace("extern __shared__ __align__(16) unsigned char rs_smem[];", "unsigned char *rs_smem = emu_dyn_smem;")
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "prim_emu.cpp")
    src = '#include "simt_emu.hpp"\nnamespace pgx {\n' + body + "\n}\n" + PRIM_DRIVER
    so = os.path.join(OUT, "libprim_emu.so")
    if os.path.exists(cpp) and os.path.exists(so) and open(cpp).read() == src and os.path.getmtime(so) >= os.path.getmtime(os.path.join(HERE, "simt_emu.hpp")):
        return so
    open(cpp, "w").write(src)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-I", HERE, "-I", os.path.join(ROOT, "include"), cpp, "-o", so])
    return so


ROUND_DRIVER = r'''
// host driver in the shape of Session::gen_arc_fast (pgx_engine.cuh): one arc build of `ranks` virtual ranks that own consecutive genome
// ranges of the same survivor list; the allgather between k_arc_pack and k_arc_merge is a memcpy
struct EmuArcIn {
	int32_t n_hit, n_sv, n_seg0, n_gene, G, use_ori, grid, H, H2, arc_cap, ranks;
	const uint32_t *sv, *bits; const int32_t *ctg, *cm, *gidp, *sori, *sdom, *pdom0, *g2s, *prot_gid, *ctg2g;
	pgx::ArcDev *arc; unsigned long long *idx; int32_t *seg_ng, *seg_tot; uint32_t *nl_off; uint32_t *out_counts;   // out_counts: n_arc, n_nl, fallback, n_ovf
};
extern "C" int round_emu_arc_build(const EmuArcIn *in)
{
	using namespace pgx;
	const int NV = 2 * in->n_seg0, R = in->ranks, TBk = 256, grid = in->grid;
	std::vector<DevCounts> dcs(R);
	std::vector<std::vector<unsigned long long>> lkey(R), gkey(R);
	std::vector<std::vector<ArcAcc>> lacc(R), gacc(R);
	std::vector<std::vector<uint32_t>> llist(R), glist(R), bm(R), degcur(R), seen(R);
	std::vector<std::vector<int32_t>> segocc(R);
	const uint32_t part_cap = in->H / 2;
	std::vector<ArcWire> all((size_t)R * (part_cap + 1));
	std::vector<ArcTable> loc(R), glob(R);
	// which survivors belong to which rank: genomes [G*r/R, G*(r+1)/R); every rank sees its own slice of sv as ITS list, genome ids made local
	for (int r = 0; r < R; ++r) {
		const int g0 = in->G * r / R, g1 = in->G * (r + 1) / R, Gl = g1 - g0 > 0? g1 - g0 : 1;
		std::vector<uint32_t> sv;
		for (int t = 0; t < in->n_sv; ++t) { int g = in->ctg2g[in->ctg[in->sv[t]]]; if (g >= g0 && g < g1) sv.push_back(in->sv[t]); }
		std::vector<int32_t> ctg2g_loc;
		int n_ctg = 0; for (int i = 0; i < in->n_hit; ++i) n_ctg = std::max(n_ctg, in->ctg[i] + 1);
		ctg2g_loc.resize(n_ctg);
		for (int i = 0; i < in->n_hit; ++i) ctg2g_loc[in->ctg[i]] = in->ctg2g[in->ctg[i]] - g0;
		memset(&dcs[r], 0, sizeof(DevCounts));
		dcs[r].m_sv = (uint32_t)sv.size(); dcs[r].n_seg = in->n_seg0;
		lkey[r].assign(in->H, KEY_EMPTY); lacc[r].assign(in->H, ArcAcc{0, 0, 0}); llist[r].assign(in->H, 0);
		const size_t bmw = (((size_t)in->H * Gl + 31) / 32 + 3) & ~(size_t)3;
		bm[r].assign(2 * bmw, 0); degcur[r].assign(2 * (size_t)NV + 8, 0); seen[r].assign(((size_t)Gl * in->n_seg0 + 31) / 32 + 4, 0); segocc[r].assign(2 * (size_t)in->n_seg0 + 4, 0);
		loc[r].key = lkey[r].data(); loc[r].acc = lacc[r].data(); loc[r].list = llist[r].data(); loc[r].bm1 = bm[r].data(); loc[r].bm2 = bm[r].data() + bmw; loc[r].H = in->H;
		glob[r] = loc[r];
		if (R > 1) { gkey[r].assign(in->H, KEY_EMPTY); gacc[r].assign(in->H, ArcAcc{0, 0, 0}); glist[r].assign(in->H, 0); glob[r].key = gkey[r].data(); glob[r].acc = gacc[r].data(); glob[r].list = glist[r].data(); glob[r].bm1 = glob[r].bm2 = nullptr; }
		std::vector<unsigned long long> l2key(in->H2, KEY_EMPTY), l2d(in->H2, 0); std::vector<uint32_t> l2n(in->H2, 0), l2list(in->H2, 0); std::vector<int32_t> l2m(2 * (size_t)in->H2, 0);
		OvfTable L; L.key = l2key.data(); L.dsum = l2d.data(); L.n = l2n.data(); L.list = l2list.data(); L.m1 = l2m.data(); L.m2 = l2m.data() + in->H2; L.H = in->H2;
		std::vector<uint32_t> rec(sv.size() + 1); std::vector<SvInfo> info(sv.size() + 1);
		uint32_t tie = 0, err = 0;
		ArcSrc e; e.sv = sv.data(); e.ctg = in->ctg; e.cm = in->cm; e.gidp = in->gidp; e.sori = in->sori; e.sdom = in->sdom; e.pdom0 = in->pdom0; e.bits = in->bits;
		e.g2s = in->g2s; e.prot_gid = in->prot_gid; e.ctg2g = ctg2g_loc.data(); e.use_ori = in->use_ori; e.n_seg0 = in->n_seg0; e.Gl = Gl;
		DevCounts *dc = &dcs[r];
		uint32_t *deg = degcur[r].data();
		int32_t *ng = segocc[r].data(), *tot = segocc[r].data() + in->n_seg0;
		emu::launch(grid, TBk, [&] { k_survivor_info(e, dc, info.data(), seen[r].data(), ng, tot, &err); });
		emu::launch(grid, TBk, [&] { k_arc_insert(info.data(), ctg2g_loc.data(), Gl, loc[r], dc, rec.data(), R > 1? (uint32_t*)nullptr : deg, &tie); });
		emu::launch(grid, TBk, [&] { k_arc_fold(info.data(), rec.data(), ctg2g_loc.data(), loc[r], dc, Gl, L); });
		emu::launch(2, TBk, [&] { k_arc_ovf_fold(L, loc[r], dc); });
		for (uint32_t j = 0; j < L.H; ++j) if (l2key[j] != KEY_EMPTY || l2n[j] || l2d[j] || l2m[j] || l2m[j + in->H2]) return -2;   // the overflow table hands itself back clean
		if (err) return -3;
		if (R > 1) { ArcWire *mine = all.data() + (size_t)r * (part_cap + 1); emu::launch(grid, TBk, [&] { k_arc_pack(loc[r], dc, mine, part_cap); }); }
	}
	// "allreduce" of the segment occupancy, then every rank merges and finalises; rank 0's result is returned, the others must agree
	std::vector<int32_t> occ(2 * (size_t)in->n_seg0 + 4, 0);
	for (int r = 0; r < R; ++r) for (size_t i = 0; i < occ.size(); ++i) occ[i] += segocc[r][i];
	std::vector<ArcDev> first;
	for (int r = 0; r < R; ++r) {
		DevCounts *dc = &dcs[r];
		uint32_t *deg = degcur[r].data(), *cur = deg + NV;
		if (R > 1) emu::launch(grid, TBk, [&] { k_arc_merge(all.data(), R, part_cap, glob[r], dc, deg); });
		const int tiles = (NV + 1 + SCAN_TILE - 1) / SCAN_TILE;
		std::vector<unsigned long long> status(2 * (size_t)tiles + 2, 0ull); uint32_t ticket = 0; unsigned long long nl_total = 0;
		std::vector<uint32_t> off(NV + 8), nl_off(NV + 8), pos_of(2 * (size_t)in->H, 0xffffffffu);
		std::vector<ArcDev> tmp(in->arc_cap + 1), arc(in->arc_cap + 1);
		std::vector<unsigned long long> idx(NV + 2, 0ull);
		emu::launch(tiles, SCAN_B, [&] { k_vertex_scan(deg, NV, off.data(), nl_off.data(), dc, (uint32_t)in->arc_cap, &nl_total, status.data(), &ticket); });
		emu::launch(grid, TBk, [&] { k_arc_place(glob[r], R > 1? &dc->n_gpart : &dc->n_part, off.data(), cur, dc, tmp.data()); });
		std::vector<SegDev> segs(in->n_seg0 + 1);
		for (auto &sg : segs) { sg.n_genome = -7; sg.tot_cnt = -7; }
		emu::launch(grid, TBk, [&] { k_arc_rank(tmp.data(), off.data(), dc, arc.data(), idx.data(), pos_of.data(), segs.data(), occ.data(), occ.data() + in->n_seg0); });
		for (int sgi = 0; sgi < in->n_seg0; ++sgi) if (segs[sgi].n_genome != occ[sgi] || segs[sgi].tot_cnt != occ[in->n_seg0 + sgi]) return -7;   // the occupancy stored beside the ranking
		if (nl_total != dc->n_nl) return -4;
		// every arc is found again through the key table (what k_branch_hits_fast does)
		for (uint32_t p = 0; p < dc->n_arc && !dc->fallback; ++p) {
			const unsigned long long x = arc[p].x, xm = arc_mirror(x);
			const uint32_t s = arc_find(glob[r], x < xm? x : xm);
			if (s == ARC_NONE || pos_of[s * 2 + (xm < x)] != p || arc[p].bits != 0) return -6;
		}
		if (r == 0) {
			memcpy(in->arc, arc.data(), (size_t)dc->n_arc * sizeof(ArcDev)); memcpy(in->idx, idx.data(), (size_t)NV * 8); memcpy(in->nl_off, nl_off.data(), (size_t)(NV + 1) * 4);
			memcpy(in->seg_ng, occ.data(), (size_t)in->n_seg0 * 4); memcpy(in->seg_tot, occ.data() + in->n_seg0, (size_t)in->n_seg0 * 4);
			in->out_counts[0] = dc->n_arc; in->out_counts[1] = dc->n_nl; in->out_counts[3] = dc->n_ovf;
			first.assign(arc.begin(), arc.begin() + dc->n_arc);
		} else if (dc->n_arc != first.size() || memcmp(first.data(), arc.data(), first.size() * sizeof(ArcDev)) != 0) return -5;
		in->out_counts[2] |= dc->fallback;
	}
	return 0;
}

extern "C" int round_emu_compact(const uint32_t *bits, const uint32_t *perm, int32_t n, uint32_t *pos, uint32_t *list, uint32_t *count)
{
	using namespace pgx;
	const int tiles = (n + 1 + SCAN_TILE - 1) / SCAN_TILE;
	std::vector<unsigned long long> status(tiles + 1, 0ull); uint32_t ticket = 0;
	DevCounts dc; memset(&dc, 0xff, sizeof(dc));
	emu::launch(tiles, SCAN_B, [&] { k_compact_survivors(bits, perm, n, pos, list, count, status.data(), &ticket, &dc); });
	return (dc.n_part | dc.n_ovf | dc.n_l2 | dc.n_gpart) == 0? 0 : -1;
}

extern "C" int round_emu_seg_filter(const pgx::SegDev *seg, pgx::SegDev *out, int32_t *n_seg, int32_t cap, const unsigned long long *idx, int32_t max_occ_total, int32_t max_degree, int32_t max_dist_loci, int32_t *g2s, int32_t n_gene)
{
	using namespace pgx;
	DevCounts dc; memset(&dc, 0, sizeof(dc)); dc.n_seg = *n_seg;
	const int tiles = (cap + 1 + SCAN_TILE - 1) / SCAN_TILE;
	std::vector<unsigned long long> status(tiles + 1, 0ull); uint32_t ticket = 0, n_keep = 0xffffffffu;
	std::vector<uint32_t> pos(cap + 8, 0x55555555u);
	for (int g = 0; g < n_gene; ++g) g2s[g] = -1;
	emu::launch(tiles, SCAN_B, [&] { k_seg_keep_scan(seg, &dc, idx, max_occ_total, max_degree, max_dist_loci, pos.data(), &n_keep, status.data(), &ticket); });
	emu::launch(3, 256, [&] { k_seg_compact_fast(seg, out, &dc, pos.data(), &n_keep, g2s); });
	*n_seg = (int32_t)n_keep;
	return 0;
}

extern "C" int round_emu_dirty(const int32_t *ctg, const int32_t *cs, const int32_t *ce, const int32_t *pmax, int32_t n, int32_t tile, const int32_t *changed, int32_t n_changed, uint32_t *tiles)
{
	using namespace pgx;
	std::vector<int2> reach(n + 1);
	emu::launch(3, 256, [&] { k_reach(ctg, cs, ce, pmax, n, tile, reach.data()); });
	DirtyCols D; D.reach = reach.data(); D.tiles = tiles;
	for (int i = 0; i < n_changed; ++i) mark_dirty(D, changed[i]);
	return 0;
}
'''


def build_round():
    """tests/emu/_build/libround_emu.so: the kernels of pgx_round.cuh (the sync-free arc build and friends) on the SIMT shim"""
    csrc = os.path.join(ROOT, "pangene_b200", "csrc")
    scan = open(os.path.join(csrc, "pgx_scan.cuh")).read()
    types = open(os.path.join(csrc, "pgx_graphtypes.cuh")).read()
    rnd = open(os.path.join(csrc, "pgx_round.cuh")).read()
    body = _between(scan, "scan") + "\n" + _between(types, "graphtypes") + "\n" + _between(rnd, "round")
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "round_emu.cpp")
    src = '#include "simt_emu.hpp"\n#include <algorithm>\n#include "pangene_abi.h"\nnamespace pgx {\n' + body + "\n}\n" + ROUND_DRIVER
    so = os.path.join(OUT, "libround_emu.so")
    if os.path.exists(cpp) and os.path.exists(so) and open(cpp).read() == src and os.path.getmtime(so) >= os.path.getmtime(os.path.join(HERE, "simt_emu.hpp")):
        return so
    open(cpp, "w").write(src)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-I", HERE, "-I", os.path.join(ROOT, "include"), cpp, "-o", so])
    return so



def build_full(sanitize=False):
    """tests/emu/_build/libpgx_emu.so (sanitize=True: libpgx_emu_asan.so, the same text under AddressSanitizer + UBSan -- a CPU-side "memcheck" of
    every kernel's global and shared-memory indexing; load it with LD_PRELOAD=libasan): ALL of libpgx -- the C ABI of include/pgx.h, the engine (pgx_engine.cuh) and every kernel -- compiled for
    the host: the product sources are copied with three kinds of textual change (the <<< >>> launch inside PGX_LAUNCH becomes a shim launch, the
    two `extern __shared__` declarations point at the shim's dynamic shared memory, the fixed grids of the grid-stride launches shrink from
    multiples of 148 to 2-3 CTAs) and see tests/emu/cuda_shim/cuda_runtime.h instead of the CUDA runtime.  Small inputs run through pg_post_process() / pg_graph_gen() of this library exactly as they would on the device."""
    csrc = os.path.join(ROOT, "pangene_b200", "csrc")
    out = os.path.join(OUT, "full")
    os.makedirs(out, exist_ok=True)
    texts = {}
    for fn in sorted(os.listdir(csrc)):
        if not (fn.endswith(".cuh") or fn.endswith(".hpp") or fn == "pgx_api.cu"):
            continue
        t = open(os.path.join(csrc, fn)).read()
        if fn == "pgx_util.cuh":
            old = "kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \\"
            assert t.count(old) == 1
            t = t.replace(old, "emu::launch_any((long)(grid), (int)(block), [&] { kernel(__VA_ARGS__); }); \\")
        for decl, name in (("extern __shared__ __align__(16) unsigned char smem_raw[];", "smem_raw"), ("extern __shared__ __align__(16) unsigned char rs_smem[];", "rs_smem")):
            if decl in t:
                assert t.count(decl) == 1
                t = t.replace(decl, "unsigned char *%s = emu_dyn_smem;" % name)
        if fn == "pgx_engine.cuh":   # grid-stride launches sized for 148 SMs: a few CTAs give the same result and spare the shim a thousand CTA turns per launch
            for old, cnt, rep_ in (("static constexpr int GS_GRID = 148 * 8;", 1, "static constexpr int GS_GRID = 3;"), ("PGX_LAUNCH(k_round_clear, 148 * 4, 256,", 2, "PGX_LAUNCH(k_round_clear, 2, 256,"),
                                   ("PGX_LAUNCH(k_arc_ovf_fold, 148, TB,", 1, "PGX_LAUNCH(k_arc_ovf_fold, 2, TB,"), ("PGX_LAUNCH(k_seg_compact_fast, 148, TB,", 1, "PGX_LAUNCH(k_seg_compact_fast, 2, TB,")):
                assert t.count(old) == cnt, old
                t = t.replace(old, rep_)
        texts["pgx_api.cpp" if fn == "pgx_api.cu" else fn] = t
    assert "<<<" not in "".join(texts.values()).replace("<<<>>>", "")
    tag = {False: "", True: "_asan", "thread": "_tsan"}[sanitize]         # sanitize="thread": ThreadSanitizer instead (racecheck of every kernel)
    stamp = os.path.join(out, ".stamp" + tag)
    so = os.path.join(OUT, "libpgx_emu%s.so" % tag)
    deps = [os.path.join(HERE, "simt_emu.hpp"), os.path.join(HERE, "cuda_shim", "cuda_runtime.h")]
    import hashlib
    h = hashlib.md5(("".join(k + v for k, v in sorted(texts.items())) + "".join(open(d).read() for d in deps)).encode()).hexdigest()
    if os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == h:
        return so
    for k, v in texts.items():
        open(os.path.join(out, k), "w").write(v)
    san = ["-fsanitize=thread", "-fno-omit-frame-pointer"] if sanitize == "thread" else ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined"] if sanitize else []
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-DPGX_SIMT_EMU=1"] + (["-g"] + san if sanitize else []) + ["-I", os.path.join(HERE, "cuda_shim"), "-I", HERE,
                           "-I", out, "-I", os.path.join(ROOT, "include"), os.path.join(out, "pgx_api.cpp"), "-o", so, "-ldl", "-lrt"])
    open(stamp, "w").write(h)
    return so


if __name__ == "__main__":
    print(build())
    print(build_prim())
    print(build_round())
