// TEST INFRASTRUCTURE: "racecheck" of the tiled sweep kernels on the CPU.  tests/emu/_build/sweep_emu.cpp is the text of k_sweep_tiled /
// k_sweep from pangene_b200/csrc/pgx_kernels.cuh on the SIMT shim (one OS thread per CUDA thread, __syncthreads() and the warp collectives
// are real barriers, atomics are __atomic builtins) -- so ThreadSanitizer sees exactly the shared-memory accesses of a CTA and reports any
// pair of them that no barrier orders: a missing __syncthreads() in the kernel shows up here as a data race.  Every variant of the tiled
// kernel is also compared with the thread-per-hit kernel on the same columns.          (built and run by tools/tsan_sweep.sh)
#include "sweep_emu.cpp"
#include <random>

int main(int argc, char **argv)
{
	const int n = argc > 1? atoi(argv[1]) : 1500, PAD = 64;
	std::mt19937_64 rng(argc > 2? atoll(argv[2]) : 7);
	auto rnd = [&](int lo, int hi) { return (int)(lo + rng() % (uint64_t)(hi - lo + 1)); };
	std::vector<int32_t> ctg(n + PAD, 0x7fffffff), cs(n + PAD), ce(n + PAD), pid(n + PAD), gidp(n + PAD), rank(n + PAD), cds(n + PAD), sori(n + PAD), pmax(n + PAD);
	std::vector<uint64_t> key(n + PAD);
	std::vector<uint32_t> bits(n + PAD), eoff(n + PAD + 1);
	std::vector<int32_t> ex;
	// loci of 1..9 overlapping hits (isoforms of a gene, a paralog, a stray hit), now and then a long hit that covers hundreds of followers
	int32_t pos = 1000, c = 0, gene = 0;
	for (int i = 0; i < n;) {
		if (rnd(0, 400) == 0) { ++c; pos = 1000; }
		const int m = std::min(n - i, rnd(1, 9));
		const int32_t locus = pos, span = rnd(600, 6000);
		++gene;
		for (int k = 0; k < m; ++k, ++i) {
			ctg[i] = c; cs[i] = locus + rnd(0, 300); ce[i] = cs[i] + span + rnd(-200, 200);
			if (rnd(0, 250) == 0) ce[i] += 400000;                                  // a window that leaves the staged range
			const int g = rnd(0, 9) == 0? gene + 100000 : gene;                      // mostly isoforms of one gene
			pid[i] = g * 8 + rnd(0, 3); gidp[i] = g << 1 | (rnd(0, 20) == 0);
			rank[i] = rnd(0, 2); sori[i] = rnd(50, 3000);
			const int32_t sadj = rnd(0, 3) == 0? 1000 : rnd(50, 3000);                // equal adjusted scores at the top of many clusters
			key[i] = (uint64_t)(int64_t)sadj << 33 | (uint64_t)(gidp[i] & 1) << 32 | (uint32_t)(pid[i] * 2654435761u);
			bits[i] = (uint32_t)rnd(0, 1) | (rnd(0, 4) == 0? (PG_HB_FLT | PG_HB_PSEUDO) : 0u) | (rnd(0, 2) == 0? PG_HB_SHADOW : 0u) | (rnd(0, 30) == 0? (uint32_t)rnd(1, 2) << 9 : 0u);
			eoff[i] = (uint32_t)(ex.size() / 2);
			int32_t x = 0, len = ce[i] - cs[i], tot = 0;                             // 1..12 exons that tile [0, ce - cs)
			const int ne = rnd(1, 12);
			for (int e = 0; e < ne; ++e) {
				const int32_t rest = len - x, w = e == ne - 1? rest : std::max(1, rest / (ne - e) - rnd(0, 40));
				const int32_t el = std::max(1, w - (e == ne - 1? 0 : rnd(0, w / 2)));
				ex.push_back(x); ex.push_back(e == ne - 1? len : x + el); tot += ex.back() - x;
				x += w;
				if (x >= len) { ex.back() = std::min(ex.back(), len); break; }
			}
			ex.back() = len;                                                          // the last exon ends where the hit ends (read.c:75)
			cds[i] = 0; for (size_t q = eoff[i] * 2; q < ex.size(); q += 2) cds[i] += ex[q + 1] - ex[q];
		}
		pos = locus + (rnd(0, 3) == 0? 200 : span + rnd(100, 40000));               // some loci overlap the next one
	}
	// (contig, cs) order, running maximum of the ends
	std::vector<int> ord(n); for (int i = 0; i < n; ++i) ord[i] = i;
	std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return ctg[a] != ctg[b]? ctg[a] < ctg[b] : cs[a] < cs[b]; });
	auto perm32 = [&](std::vector<int32_t> &v) { std::vector<int32_t> t(v); for (int i = 0; i < n; ++i) v[i] = t[ord[i]]; };
	std::vector<int32_t> ex2; std::vector<uint32_t> eoff2(n + PAD + 1);
	eoff[n] = (uint32_t)(ex.size() / 2);
	for (int i = 0; i < n; ++i) { eoff2[i] = (uint32_t)(ex2.size() / 2); for (uint32_t q = eoff[ord[i]]; q < (ord[i] + 1 < n? eoff[ord[i] + 1] : eoff[n]); ++q) { ex2.push_back(ex[2 * q]); ex2.push_back(ex[2 * q + 1]); } }
	for (int i = n; i <= n + PAD; ++i) eoff2[i] = (uint32_t)(ex2.size() / 2);
	ex2.resize(ex2.size() + 2 * PAD, 0);
	perm32(ctg); perm32(cs); perm32(ce); perm32(pid); perm32(gidp); perm32(rank); perm32(cds); perm32(sori);
	{ std::vector<uint64_t> t(key); for (int i = 0; i < n; ++i) key[i] = t[ord[i]]; std::vector<uint32_t> b(bits); for (int i = 0; i < n; ++i) bits[i] = b[ord[i]]; }
	for (int i = 0, run = -1, prev = -1; i < n; ++i) { if (ctg[i] != prev) { run = -1; prev = ctg[i]; } run = std::max(run, ce[i]); pmax[i] = run; }
	bits[0] |= 0x80000000u;
	EmuCols e; e.n = n; e.ctg = ctg.data(); e.cs = cs.data(); e.ce = ce.data(); e.pid = pid.data(); e.gidp = gidp.data(); e.rank = rank.data(); e.cds = cds.data(); e.sori = sori.data(); e.pmax = pmax.data();
	e.key = key.data(); e.bits = bits.data(); e.eoff = eoff2.data(); e.ex = ex2.data();
	int bad = 0;
	for (int mode = 0; mode < 3; ++mode) {
		std::vector<uint32_t> rb(n, 0xdeadbeefu); std::vector<int32_t> rp(n, -7), rs(n, -7); uint32_t tie = 0;
		if (sweep_emu(0, mode, &e, 0.5, 0, rb.data(), rp.data(), rs.data(), &tie)) return 2;
		long n_dom = 0; for (int i = 0; i < n; ++i) n_dom += mode == 2? ((rb[i] & PG_HB_FLT_OV) != 0) : rp[i] >= 0;
		for (int variant = 1; variant <= 4; ++variant) {
			std::vector<uint32_t> gb(n, 0xdeadbeefu); std::vector<int32_t> gp(n, -7), gs(n, -7); uint32_t t2 = 0;
			if (sweep_emu(variant, mode, &e, 0.5, 0, gb.data(), gp.data(), gs.data(), &t2)) return 2;
			long diff = 0;
			for (int i = 0; i < n; ++i) diff += ((mode == 0 && variant > 1)? 0 : gb[i] != rb[i]) + (mode != 2? gp[i] != rp[i] : 0) + (mode == 0? gs[i] != rs[i] : 0);
			printf("mode %d variant %d: %ld of %d hits dominated / filtered, %ld differences from the thread-per-hit kernel\n", mode, variant, n_dom, n, diff);
			bad += diff != 0;
		}
	}
	printf(bad? "SWEEP_RACE FAIL\n" : "SWEEP_RACE PASS (results)\n");
	return bad? 1 : 0;
}
