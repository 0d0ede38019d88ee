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
#define RUN(K, T) emu::launch((c.n + (T) - 1) / (T), (T), [&] { K(c, o, bits_out, pdom_out, sdom_out, tie_cnt); })
#define BY_MODE(K, T) do { if (mode == 0) RUN(K<SW_LOAD>, T); else if (mode == 1) RUN(K<SW_SHADOW>, T); else RUN(K<SW_ISO>, T); } while (0)
	if (variant == 0) BY_MODE(k_sweep, 128);                   // plain thread-per-hit kernel
	else if (variant == 1) BY_MODE(k_sweep_tiled, SWQ_TILE);   // production kernel
	else return -1;
	return 0;
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


if __name__ == "__main__":
    print(build())
    print(build_prim())
