"""A/B harness: time the fine-level streamed SpMM alone (pgb200_ert_bench_spmm) for a workload.
   python ab/bench_spmm.py [workload] [reps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygimli_b200 import workloads, _capi
from pygimli_b200.dist import ShardedERT
w = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cols = [int(x) for x in sys.argv[3:5]] if len(sys.argv) > 4 else None
r = workloads.WORKLOADS[w](1.0)
mesh, scheme, kw = r[0], r[1], (r[3] if len(r) > 3 else None)
fop = ShardedERT(mesh, scheme, kw=kw)
if cols:
    fop.core.setShard(cols[0], cols[1], 0, fop.D * (cols[1] - cols[0]) // fop.nS)
model = workloads.model_for(fop.M)
if os.environ.get("PGB200_MMA_DBG", "0") == "0":
    fop.response(model)
else:
    # assemble only: the solve would not converge with parts of the kernel disabled
    os.environ["PGB200_MMA_DBG_SAVE"] = os.environ["PGB200_MMA_DBG"]
    raise SystemExit("set PGB200_MMA_DBG_LATE instead")
h = fop.core._ensure_handle()
late = os.environ.get("PGB200_MMA_DBG_LATE")
out = []
for role in (0, 1, 2):
    ms = C.c_double(0.0)
    _capi.check(_capi.lib().pgb200_ert_bench_spmm(h, role, reps, C.byref(ms)))
    out.append(ms.value * 1e3)
print(f"{w} cols={cols} spmm {out[0]:.1f} us  post {out[1]:.1f} us  residual {out[2]:.1f} us   (dbg_late={late})")
