"""Warm per-kernel timeline of one step: python profiles/summarize_trace.py [--workload c3] > profiles/rNN_trace_c3.txt
Runs on the GPU box.  Uses the library's trace mode (pgb200_ert_set_profile(h, 2): one CUDA event per launch, CUDA graph
off), groups the launches by launch site (source line of csrc/pgb200_ert.cu -> kernel name) and multilevel level."""
import argparse
import collections
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cols", type=int, nargs=2, default=None, help="solve only the source columns [c0, c1) (what one rank of an N-GPU run does)")
    args = ap.parse_args()
    from pygimli_b200 import workloads
    from pygimli_b200.dist import ShardedERT
    src = open(os.path.join(ROOT, "pygimli_b200", "csrc", "pgb200_ert.cu")).read().split("\n")
    r = workloads.WORKLOADS[args.workload](args.scale)
    mesh, scheme, kw = r[0], r[1], (r[3] if len(r) > 3 else None)
    fop = ShardedERT(mesh, scheme, kw=kw)
    if args.cols:
        fop.core.setShard(args.cols[0], args.cols[1], 0, fop.D * (args.cols[1] - args.cols[0]) // fop.nS)
    model = workloads.model_for(fop.M)
    fop.response(model)                                   # warm-up (graph mode)
    fop.create_jacobian(model)
    fop.core.setProfile(2)
    fop.response(model)
    fop.create_jacobian(model)
    lines, ms = fop.core.trace()
    fop.core.setProfile(0)
    st = fop.core.stats()

    def name(line):
        for back in range(0, 6):                          # the kernel name is on the launch line or just above it
            m = re.findall(r"\b(k_\w+|PANEL_GO|SPMM_GO|RSGO|RGO)\b", src[line - 1 - back]) if line - 1 - back >= 0 else []
            if m:
                return m[0]
        return f"line {line}"
    agg = collections.OrderedDict()
    for code, t in zip(lines, ms):
        role = {0: "", 1: " spmm", 2: " post", 3: " residual"}[(int(code) % 256) // 16]
        key = (name(int(code) // 256) + role, int(code) % 16, int(code) // 256)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(t)
    total = float(ms.sum())
    iters = max(1.0, st["pcg_iterations"])
    print(f"# warm per-launch timeline, workload {args.workload}: {len(ms)} launches, {total:.1f} ms, {int(iters)} PCG iterations"
          f" ({total / iters * 1e3:.0f} us per iteration incl. set-up/epilogue; events add ~2 us per launch, CUDA graph off)")
    print(f"{'kernel (launch site)':38s} {'level':>5s} {'launches':>9s} {'total ms':>10s} {'us/launch':>10s} {'us/iter':>9s} {'share':>7s}")
    for (nm, lvl, line), (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{nm + ' @' + str(line):38s} {lvl:5d} {cnt:9d} {t:10.2f} {t / cnt * 1e3:10.1f} {t / iters * 1e3:9.1f} {100 * t / total:6.1f}%")


if __name__ == "__main__":
    main()
