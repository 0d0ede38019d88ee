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


if __name__ == "__main__":
    print(build())
