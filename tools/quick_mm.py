#!/usr/bin/env python3
"""tools/quick_mm.py -- quick per-launch timing of the many-to-many packed kernels on a reduced panel (for A/B runs of
tuning builds: AMBIVERT_B200_LIB=ambivert_b200/align/tuning/lib_x.so python tools/quick_mm.py ...).  Prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=20000)
    ap.add_argument("--targets", type=int, default=2000)
    ap.add_argument("--mode", default="global")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    import torch
    from ambivert_b200 import batch as B, synth
    B.set_device(0)
    for kv in args.opt:
        B.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    targets = synth.panel_targets(args.targets)
    queries, _ = synth.panel_queries(targets, args.queries)
    targets = [t.upper() for t in targets]
    job = B.BestHitsJob(queries, targets, B.MODE_BY_NAME[args.mode], B.SEED_GLOBAL)
    st = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = {}
    total = []
    with torch.cuda.stream(st):
        for it in range(args.steps + 1):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            job.launch(st.cuda_stream)
            e1.record()
            st.synchronize()
            if it == 0:
                continue
            total.append(e0.elapsed_time(e1))
            for r in job.profile():
                rows.setdefault("%d%s" % (r["K"], "b" if r["lanes"] == 31 else ""), []).append((r["ms"], r["cells"]))
    import hashlib
    sc, ix = job.fetch()
    digest = hashlib.sha256(sc.tobytes() + ix.tobytes()).hexdigest()[:12]
    out = {"digest": digest, "lib": os.environ.get("AMBIVERT_B200_LIB", "default"), "mode": args.mode, "opt": args.opt, "step_ms": min(total),
           "tcups_step": job.cells / min(total) / 1e9,
           "per_K": {K: {"ms": min(m for m, _ in v), "tcups": v[0][1] / min(m for m, _ in v) / 1e9} for K, v in sorted(rows.items())}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
