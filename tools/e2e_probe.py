import ctypes as C, os, sys, time
sys.path.insert(0,'/root/repo')
from pangene_b200 import engine, hostio
opt=hostio.default_opt()
d0=hostio.synth(opt, hostio.synth_params(3))
L=engine.lib(); engine.set_verbose(3 if os.environ.get("PGX_TIMING") else 1); L.pgx_set_lazy(1)
for it in range(int(os.environ.get("E2E_ITERS", "4"))):
    d=hostio.lib().pgio_data_clone(d0)
    t0=time.perf_counter()
    L.pg_post_process(C.byref(opt), d); g=L.pg_graph_init(d); L.pg_graph_gen(C.byref(opt), g)
    print('e2e %.1f ms'%((time.perf_counter()-t0)*1e3), flush=True)
    L.pg_graph_destroy(g); hostio.lib().pgio_data_clone_destroy(d)
