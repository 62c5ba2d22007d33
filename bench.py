#!/usr/bin/env python3
"""bench.py -- hashtable-build throughput of the AmBiVErT alignment hot path on B200.

A "step" is one pass of the hot path over one batch: every thresholded unique amplicon of the batch
scored against every manifest target (affine-gap global alignment, DNA 1/-4/N=0, gaps -7/-1) and
reduced to the best target per amplicon -- what ambivert.py:515-529 does for one run.  The default
workload is BASELINE.json configs[3]: 100,000 unique amplicons x 2,000 targets,
synthetic, sharded over --gpus N GPUs of one box (strong scaling: the job is fixed), with one NCCL
gather of the best hits per step.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path
    python bench.py --impl reference ...                      # the reference's CPU C path (oracle/_ref)

Prints ONE JSON line (rank 0).  `value` = GCUPS with the inputs resident in HBM; `e2e` = the same
through the host-buffer API with the H2D/D2H copies inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# lane-ops per packed cell pair of the two kernel generations and the abv_int_peak mix that measures each
CELL_OPS_STREAMING = 4       # mm16g_kernel: 1 add + VIMNMX3.S16x2 + 2 VIADDMNMX.S16x2          -> abv_int_peak(8)
CELL_OPS_FIRST_GEN = 5       # mm16_kernel:  2 VIADD.16x2 + VIMNMX3.S16x2 + 2 VIADDMNMX.S16x2   -> abv_int_peak(4)
TB16_LOCAL_K5_LOOP_INSTR = 646   # SASS instructions of tb16_kernel<local,K=5>'s fill loop (4 steps): profiles/r02_sass_tb16_local_k5_step.txt
TB16_LOOP_STEPS = 4
MODE_NAMES = {"global": 2, "glocal": 3, "local": 1}
WORKLOADS = {"config4": (100000, 2000), "config5": (1000000, 10000)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--mode", default="global", choices=list(MODE_NAMES))
    ap.add_argument("--workload", default="config4", choices=list(WORKLOADS),
                    help="config4 = BASELINE.json configs[3] (100k amplicons x 2k manifest targets, the headline); "
                         "config5 = configs[4] (1M x 250 bp amplicons x 10k FASTA targets; meant for --gpus 8)")
    ap.add_argument("--queries", type=int, default=0, help="override the workload's amplicon count (the label then says so)")
    ap.add_argument("--targets", type=int, default=0, help="override the workload's target count")
    ap.add_argument("--glocal-steps", type=int, default=5, help="timed steps of the glocal block (>= 5 unless --no-glocal)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work of one cpu_baseline sample / reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-glocal", action="store_true", help="skip the secondary glocal-mode measurement of the same job")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="library tuning option (abv_set_option), e.g. no_levels=1")
    ap.add_argument("--merge-pairs", type=int, default=1000000, help="read pairs of the secondary overlap-merge measurement (0 = skip)")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the pipeline block (merge -> match -> align through host.AmpliconData)")
    ap.add_argument("--pipeline-gpus", type=int, default=0, help="GPUs the pipeline block's match_to_reference(gpus=) uses from this ONE process (0 = all visible at N=1)")
    return ap.parse_args()


def workload(args):
    """synthetic inputs of SURVEY.md 8(d): config 4 = manifest-style targets of 200-260 bp with soft-masked primers x
    mutated-copy amplicons; config 5 = 250 bp FASTA targets (seed 0xD5) x amplicons generated the same way"""
    from ambivert_b200 import synth
    n_q0, n_t0 = WORKLOADS[args.workload]
    n_q, n_t = args.queries or n_q0, args.targets or n_t0
    if args.workload == "config5":
        targets = synth.fasta_targets(n_t)
        queries, source = synth.panel_queries(targets, n_q, seed=0xD5 + 1)
        what, cfg = "FASTA targets of 250 bp", "BASELINE.json configs[4]"
    else:
        targets = synth.panel_targets(n_t)
        queries, source = synth.panel_queries(targets, n_q)
        what, cfg = "manifest targets", "BASELINE.json configs[3]"
    if (n_q, n_t) != (n_q0, n_t0):
        cfg = "generator of %s at a NON-STANDARD size (%d x %d is the config)" % (cfg, n_q0, n_t0)
    targets = [t.upper() for t in targets]          # the call site upper-cases the target (ambivert.py:169)
    name = "hashtable build: %d unique amplicons x %d %s, %s alignment (%s)" % (len(queries), len(targets), what, args.mode, cfg)
    return queries, targets, name


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_arm(queries, targets, mode, seconds, threads=None):
    """the reference's CPU implementation of the path (oracle/_ref compiled from lib/align/*.c, else the
    oracle port) on a bounded sample of the same workload, all host threads, full per-call work
    (direction matrix + traceback + fragment list, as the C extension does)"""
    import oracle as O
    O.build(ref=True)
    port = O.port()
    kind, fns, build = "port", None, "oracle/align_oracle.c, gcc -O2"
    if O.Ref.available():
        kind, fns, build = "reference", O.ref().fns(mode), "lib/align/*.c, gcc -O2 (setuptools-like)"
    threads = threads or os.cpu_count() or 1
    t_codes = [O.encode(t) for t in targets]
    seed = -1000000

    def run(sample):
        q_codes = [O.encode(q) for q in sample]
        sec, bs, bi = port.timed_best_hits(mode, q_codes, t_codes, 5, O.DNA_SCORE, -7, -1, seed, threads, fns=fns)
        cells = float(sum(map(len, sample))) * float(sum(map(len, targets)))
        return sec, cells, bs, bi
    n0 = min(len(queries), max(threads, 8))
    sec, cells, _, _ = run(queries[:n0])                       # calibration
    if kind == "reference" and os.path.exists(O.REF_O3_SO):    # the same sources at -O3 -funroll-loops: keep the faster build
        o3 = O.Ref(O.REF_O3_SO)
        fns_o2, fns = fns, o3.fns(mode)
        sec3, _, _, _ = run(queries[:n0])
        if sec3 < sec:
            sec, build = sec3, "lib/align/*.c, gcc -O3 -funroll-loops (faster than -O2 on this host)"
        else:
            fns = fns_o2
    rate = cells / max(sec, 1e-6)
    per_q = cells / n0
    n = int(min(len(queries), max(n0, seconds * rate / per_q)))
    return dict(kind=kind, threads=threads, run=run, n=n, build=build)


def reference_arm(args, rank):
    if rank != 0:
        return
    queries, targets, name = workload(args)
    mode = MODE_NAMES[args.mode]
    arm = cpu_arm(queries, targets, mode, args.cpu_seconds)
    n = arm["n"]
    for _ in range(args.warmup):
        arm["run"](queries[:max(1, n // 8)])
    t_total, c_total = 0.0, 0.0
    for k in range(args.steps):
        lo = (k * n) % max(1, len(queries) - n + 1)
        sec, cells, _, _ = arm["run"](queries[lo:lo + n])
        t_total += sec; c_total += cells
    gcups = c_total / t_total / 1e9
    sample = "%d of %d queries x all %d targets per step (%.2e cells), %d pthreads, %s" % (
        n, len(queries), len(targets), c_total / args.steps, arm["threads"],
        "unmodified " + arm["build"] if arm["kind"] == "reference" else arm["build"])
    line = {"impl": "reference", "metric": "hashtable-build alignment throughput", "value": gcups, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": name, "sampled": sample},
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": arm["threads"], "kind": arm["kind"], "sample": sample},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "projected_full_job_s": float(sum(map(len, queries))) * float(sum(map(len, targets))) / (gcups * 1e9)}
    emit(line)


def as_shipped_measurement(queries, targets, gpu_score, gpu_idx, n_sample=64):
    """the reference AS SHIPPED -- its unmodified Python package with its own C extension
    (oracle/_ref/pkg_cpu, vendored by `make -C oracle ref`), AmpliconData.match_to_reference(threads=ncores): one ctypes call
    per (amplicon, target) from a multiprocessing.dummy.Pool per amplicon (ambivert.py:461-485).  Run in a subprocess on the
    package's bundled data and on n_sample amplicons x all targets of this workload; the chosen targets of the sample are
    compared with the GPU's (min_score 0.1 applied as the reference does, ambivert.py:528)."""
    import oracle as O
    import tempfile
    if not O.as_shipped_available():
        return {"unavailable": "oracle/_ref/pkg_cpu was not vendored (no /root/reference when build() ran)"}
    script = os.path.join(ROOT, "oracle", "as_shipped.py")
    threads = os.cpu_count() or 1

    def run(extra):
        out = subprocess.run([sys.executable, "-W", "ignore", script, "--threads", str(threads)] + extra, stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True, timeout=600)
        return json.loads(out.stdout.strip().splitlines()[-1])
    try:
        bundled = run(["--bundled"])
        pick = np.unique(np.linspace(0, len(queries) - 1, min(n_sample, len(queries))).astype(np.int64))
        with tempfile.TemporaryDirectory() as d:
            with open(os.path.join(d, "q.txt"), "w") as f:
                f.write("\n".join(queries[i].decode("ascii") for i in pick))
            with open(os.path.join(d, "t.txt"), "w") as f:
                f.write("\n".join(t.decode("ascii") for t in targets))
            sample = run(["--queries", os.path.join(d, "q.txt"), "--targets", os.path.join(d, "t.txt")])
    except (subprocess.SubprocessError, ValueError, IndexError, OSError) as e:
        return {"unavailable": "as-shipped run failed: %r" % (e,)}
    want = [int(gpu_idx[i]) if gpu_score[i] > 0.1 else -1 for i in pick]
    return {"value": sample["gcups"], "unit": "GCUPS", "cores": threads, "kind": "reference",
            "api": sample["api"], "extension": sample["extension"],
            "sample": "%d amplicons evenly drawn from the workload x all %d targets (%.2e cells, %.1f s)" % (len(pick), len(targets), sample["cells"], sample["seconds"]),
            "chosen_targets_match_gpu": sample["chosen"] == want,
            "bundled_test_data": {"value": bundled["gcups"], "unit": "GCUPS", "seconds": bundled["seconds"], "cells": bundled["cells"],
                                  "what": "%s: %d merged amplicons x %d manifest targets (BASELINE.json configs[0] data)" % (bundled["what"], bundled["amplicons"], bundled["targets"])}}


def ncu_traffic(mode, K, lanes=32):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this workload, from the
    committed ncu capture (profiles/traffic.json); None if that kernel/workload was not captured"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("mm16g_kernel<%s,K=%d,lanes=%d>" % (mode, K, lanes))
    except (OSError, ValueError):
        return None


def tb16_hbm(n_pairs, kernel_ms):
    """the traceback stream of the merge stage against the HBM peak: the direction planes (4 bits per cell) are written once, the walk
    reads its path from L2; traffic per pair from the committed ncu capture (profiles/traffic.json)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            per_pair = json.load(f).get("tb16_kernel<local,K=5> bytes per read pair")
    except (OSError, ValueError):
        per_pair = None
    peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f).get("hbm_gbs")
    except (OSError, ValueError):
        pass
    alg = 150 * 150 * 4 / 8.0 + 300 + 16                       # direction planes + the two reads + score / fragment record, per pair
    out = {"algorithmic_bytes_per_pair": alg, "achieved_gbs": alg * n_pairs / (kernel_ms * 1e-3) / 1e9, "peak_gbs": peak if peak else 6650.0,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peak else "fallback of B200_PROFILING.md",
           "traffic_bytes_per_pair_ncu": per_pair}
    out["frac_of_hbm"] = out["achieved_gbs"] / out["peak_gbs"]
    return out


def hbm_info(cells, n_q, n_t, total_cells, io_bytes_step, avg_ms):
    """the sequence and result streams of the dominant launch against the measured HBM peak: for information,
    the path is compute-bound (hundreds of thousands of cells per byte)"""
    alg = io_bytes_step * cells / total_cells
    peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f).get("hbm_gbs")
    except (OSError, ValueError):
        pass
    gbs = alg / (avg_ms * 1e-3) / 1e9
    return {"algorithmic_bytes_per_launch": alg, "achieved_gbs": gbs, "peak_gbs": peak if peak else 6650.0,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peak else "fallback of B200_PROFILING.md",
            "frac": gbs / (peak if peak else 6650.0), "cells_per_byte": cells / alg,
            "note": "sequence + result streams; the kernel is bound by the integer/DPX pipe, not by HBM"}


def merge_measurement(n_pairs):
    """secondary: BASELINE.json configs[1], the overlap merge of unique F/R read pairs -- one batched local
    alignment WITH traceback per pair (ambivert.py:320-328), host buffers in, scores + fragment lists out"""
    from ambivert_b200 import batch as B, synth
    import torch
    a, b = synth.read_pairs_packed(n_pairs)
    pin = lambda x: torch.from_numpy(x).pin_memory().numpy()            # host buffers are pinned, as for the headline e2e
    a, b = (pin(a[0]), a[1]), (pin(b[0]), b[1])
    B.align_pairs((a[0][:150 * 1000], a[1][:1001]), (b[0][:150 * 1000], b[1][:1001]), B.LOCAL)      # warm-up
    best = None
    bufs = B.align_pairs_buffers(n_pairs, 4 * n_pairs)                  # results land in pinned memory too (scores, fragment lists)
    for _ in range(3):
        t0 = time.perf_counter()
        sc, st, cn, fr = B.align_pairs(a, b, B.LOCAL, out=bufs)
        wall = time.perf_counter() - t0
        tm = B.timing()
        if best is None or wall < best[0]:
            best = (wall, tm, int(fr.shape[0]), int((cn > 0).sum()))
    wall, tm, nfr, naligned = best
    # the whole stage: alignment + merged reads assembled on the device (abv_merge_pairs, ambivert.py:320-327)
    stage = None
    out = B.pinned_empty(a[0].size + b[0].size + 16)                    # merged reads land in pinned memory too
    for _ in range(3):
        t0 = time.perf_counter()
        msc, mst, moff, mbuf = B.merge_pairs(a, b, 20, out=out)
        w = time.perf_counter() - t0
        if stage is None or w < stage[0]:
            stage = (w, int((mst == 1).sum()), int(mbuf.size))
    # Issue-rate roofline of tb16_kernel<local,K=5>: one warp instruction per scheduler and clock, every warp instruction of the
    # fill loop advances 32 lanes x 2 halves = 64 cells by 1 / (instructions per cell pair).  The count is the steady loop of the
    # shipped binary (four steps of K = 5 cell pairs per lane: tools/sass_report.py 'tb16_kernel<1, 5>', committed as
    # profiles/r02_sass_tb16_local_k5_step.txt); the clock is the maximum SM clock nvidia-smi reports.
    roof = None
    try:
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        mhz = float(subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits"],
                                   stdout=subprocess.PIPE, text=True, timeout=20).stdout.split()[0])
        issue = sms * 4 * mhz * 1e6                                  # warp instructions per second, all schedulers
        per_cell_pair = TB16_LOCAL_K5_LOOP_INSTR / (TB16_LOOP_STEPS * 5.0)
        peak = issue * 64 / per_cell_pair / 1e9
        ach = tm["cells"] / tm["kernel_ms"] / 1e6
        roof = {"bound": "issue", "kernel": "tb16_kernel<local,K=5>", "achieved": ach, "peak": peak, "unit": "GCUPS", "frac": ach / peak,
                "peak_source": "%d SMs x 4 schedulers x %.0f MHz = %.3e warp-inst/s x 64 cells / %.1f instructions per cell pair (SASS of the fill loop: "
                               "%d instructions per %d steps of 5 cell pairs; profiles/r02_sass_tb16_local_k5_step.txt)" % (
                                   sms, mhz, issue, per_cell_pair, TB16_LOCAL_K5_LOOP_INSTR, TB16_LOOP_STEPS),
                "avg_launch_ms": tm["kernel_ms"], "hbm": tb16_hbm(n_pairs, tm["kernel_ms"]), "note": "fill + walk of every chunk (CUDA events in the library); 29-step ramp per item and "
                "2 idle lanes at K = 5 are inside the achieved figure (profiles/r02_issue_sensitivity.md section 6)"}
    except Exception as e:                                           # noqa: BLE001  (the block is informative)
        roof = {"unavailable": repr(e)}
    return {"roofline": roof,
            "workload": "overlap merge: %d unique 150 bp F/R read pairs, local alignment + traceback (BASELINE.json configs[1])" % n_pairs,
            "cells": tm["cells"], "kernel_ms": tm["kernel_ms"], "gcups_kernel": tm["cells"] / tm["kernel_ms"] / 1e6,
            "e2e_s": wall, "gcups_e2e": tm["cells"] / wall / 1e9, "pairs_per_s_e2e": n_pairs / wall,
            "h2d_bytes": tm["h2d_bytes"], "d2h_bytes": tm["d2h_bytes"], "fragments": nfr, "pairs_with_alignment": naligned,
            "stage_with_merged_reads": {"e2e_s": stage[0], "pairs_per_s_e2e": n_pairs / stage[0], "merged": stage[1], "merged_bytes": stage[2],
                                        "api": "abv_merge_pairs: alignment + traceback + merged read assembly on the device"},
            "kernel": "tb16_kernel<local,K=5> (two pairs per warp, 16-bit halves, traceback)", "timing": "best of 3 calls of abv_align_pairs (CUDA events inside the library for the kernel)"}


def pipeline_measurement(n_unique, n_targets, gpus):
    """what the user calls: the on-path pipeline stages through the host mirror of the reference's AmpliconData
    (ambivert_b200/host.py; ambivert.py:244-262, :301-331, :365-541, :543-574) at config-4 size -- n_unique unique F/R read
    pairs x n_targets manifest targets -- from Python objects in to Python objects out.  Per stage: wall time, the device
    part of it (H2D + kernels + D2H, CUDA events inside the library) and the rest (Python + host-side planning)."""
    from ambivert_b200 import batch as B, synth
    from ambivert_b200.host import AmpliconData
    refs, pairs = synth.panel_pipeline_inputs(n_targets=n_targets, n_unique=n_unique, seed=0xC4A)
    recs = [(("f%d" % i, f, "I" * len(f)), ("r%d" % i, r, "I" * len(r))) for i, (f, r, _c) in enumerate(pairs)]

    def run():
        amp = AmpliconData()
        for k, sq in refs:
            amp.reference_sequences[k] = sq
        stages = {}

        def stage(name, fn):
            t0 = time.perf_counter()
            fn()
            wall = time.perf_counter() - t0
            tm = B.timing()
            dev = (tm["h2d_ms"] + tm["kernel_ms"] + tm["d2h_ms"]) * 1e-3
            stages[name] = {"wall_s": wall, "device_s": dev, "kernel_ms": tm["kernel_ms"], "python_and_host_s": max(0.0, wall - dev),
                            "python_share": max(0.0, wall - dev) / wall if wall > 0 else 0.0, "cells": tm["cells"]}
        stage("cluster (add_read_pairs)", lambda: amp.add_read_pairs(recs))
        stage("merge_overlaps", lambda: amp.merge_overlaps(threshold=0, minimum_overlap=20))
        stage("match_to_reference", lambda: amp.match_to_reference(gpus=gpus))
        stage("align_to_reference", lambda: amp.align_to_reference())
        return amp, stages
    run()                                                   # warm-up: device contexts, allocation caches, kernel attributes
    amp, stages = run()
    m, a = stages["match_to_reference"], stages["align_to_reference"]
    out = {"workload": "merge -> match -> align through host.AmpliconData: %d unique 150 bp F/R read pairs x %d manifest targets (config-4 size), %s" % (
               len(pairs), len(refs), "gpus=%r" % (gpus,)),
           "api": "ambivert_b200.host.AmpliconData.add_read_pairs / merge_overlaps / match_to_reference(gpus=) / align_to_reference",
           "stages": stages, "total_wall_s": sum(x["wall_s"] for x in stages.values()),
           "merged": len(amp.merged), "unmergable": len(amp.unmergable), "matched": len(amp.reference), "aligned": len(amp.aligned),
           "potential_variants": len(amp.potential_variants),
           "match_gcups_wall": m["cells"] / m["wall_s"] / 1e9 if m["wall_s"] > 0 else None,
           # job 3 (ambivert.py:543-574): one global alignment WITH traceback per matched amplicon, packed kernel tb16_kernel<global,K>
           "job3_final_alignment": {"pairs": len(amp.aligned), "cells": a["cells"], "kernel_ms": a["kernel_ms"],
                                    "gcups_kernel": a["cells"] / a["kernel_ms"] / 1e6 if a["kernel_ms"] > 0 else None,
                                    "kernel": "tb16_kernel<global,K=8> (+ assemble kernels for the gapped rows)"}}
    return out


def cluster_measurement(n_pairs):
    """secondary: the step before the path -- md5 keys + grouping of read pairs on the device
    (AmpliconData.add_reads, ambivert.py:244-262), beside hashlib in a Python loop on a sample"""
    import hashlib
    from ambivert_b200 import batch as B, synth
    (f, fo), (r, ro) = synth.read_pairs_packed(n_pairs, seed=0xF2)
    dup = np.random.default_rng(3).integers(0, n_pairs // 4, n_pairs)          # ~4 copies per unique pair, shuffled
    f2, r2 = f.reshape(n_pairs, -1)[dup].reshape(-1), r.reshape(n_pairs, -1)[dup].reshape(-1)
    import torch
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()   # host buffers are pinned, as for the headline e2e
    f2, r2 = pin(f2), pin(r2)
    B.cluster_pairs((f2[:150000], fo[:1001]), (r2[:150000], ro[:1001]))
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        cluster_of, keys, count, first = B.cluster_pairs((f2, fo), (r2, ro))
        w = time.perf_counter() - t0
        tm = B.timing()
        if best is None or w < best[0]:
            best = (w, tm["kernel_ms"], len(keys))
    m = min(n_pairs, 100000)
    fb, rb = f2.tobytes(), r2.tobytes()
    t0 = time.perf_counter()
    ref = {}
    for i in range(m):
        ref.setdefault(hashlib.md5(fb[150 * i:150 * i + 150] + rb[150 * i:150 * i + 150]).hexdigest(), []).append(i)
    py_s = time.perf_counter() - t0
    same = [keys[k] for k in cluster_of[:m]] == [k for k, v in sorted(((k, i) for k, v in ref.items() for i in v), key=lambda x: x[1])]
    return {"workload": "md5 keying + grouping of %d read pairs (2 x 150 bp), ~4 copies per unique pair" % n_pairs,
            "e2e_s": best[0], "device_ms": best[1], "pairs_per_s_e2e": n_pairs / best[0], "clusters": best[2],
            "python_hashlib_pairs_per_s": m / py_s, "python_sample": m, "keys_match_hashlib_on_sample": bool(same)}


def variants_measurement(n):
    """secondary: the step after the path -- the variant scan of n aligned amplicons on the device
    (caller(), ambivert/call_variants.py:32-128) beside the oracle restatement of that Python loop on a sample"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import variants_oracle as VO
    from ambivert_b200 import batch as B, synth
    base = synth.aligned_rows(4000, 0xF4)
    rows = [base[i % len(base)] for i in range(n)]
    s = B.pack([r[0].encode() for r in rows])
    r = B.pack([x[1].encode() for x in rows])
    B.call_variants(s, r)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        status, start, count, recs = B.call_variants(s, r)
        w = time.perf_counter() - t0
        tm = B.timing()
        if best is None or w < best[0]:
            best = (w, tm["kernel_ms"], int(len(recs)))
    m = 2000
    t0 = time.perf_counter()
    same = True
    for k in range(m):
        want = VO.scan(rows[k][0], rows[k][1])
        got = recs[start[k]:start[k] + count[k]].tolist()
        same = same and [("XDI"[a], b, c, (rows[k][1] if a == 1 else rows[k][0])[e:e + d]) for a, b, c, d, e in got] == want
    py_s = time.perf_counter() - t0
    return {"workload": "variant scan of %d aligned amplicons (~235 sites each)" % n, "e2e_s": best[0], "device_ms": best[1],
            "amplicons_per_s_e2e": n / best[0], "records": best[2], "python_port_amplicons_per_s": m / py_s, "python_sample": m,
            "records_match_port_on_sample": bool(same)}


def parity_sample(bounds, n_want):
    """query indices drawn from ALL shards: per shard its first and last query and evenly spaced ones between"""
    world = len(bounds) - 1
    per = max(2, -(-int(n_want) // world))
    idx = set()
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        if hi > lo:
            idx.update(np.unique(np.linspace(lo, hi - 1, min(per, hi - lo)).astype(np.int64)).tolist())
    return np.asarray(sorted(idx), np.int64)


def result_digest(score, idx):
    """sha256 over the gathered (best_score, best_idx) int32 arrays in query order: equal digests at N = 1/2/4/8 mean the
    hash table does not depend on the sharding"""
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(score, np.int32).tobytes() + np.ascontiguousarray(idx, np.int32).tobytes()).hexdigest()


def cuda_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ambivert_b200 import batch as B
    from ambivert_b200.dist import ShardedBestHits, gather_best_hits, shard_bounds_by_work

    if not torch.cuda.is_available() or B.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    B.set_device(local_rank)
    for kv in args.opt:
        B.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    mode = MODE_NAMES[args.mode]
    queries, targets, name = workload(args)
    total_cells = float(sum(map(len, queries))) * float(sum(map(len, targets)))
    bounds = shard_bounds_by_work([len(q) for q in queries], world)

    # measured denominator of the roofline: the instruction mix of one packed DP cell pair, register-only.
    # global/glocal: the streaming kernel's mix; local: the first-generation kernel's mix
    def measured_peak(streaming):
        which, ops = (8, CELL_OPS_STREAMING) if streaming else (4, CELL_OPS_FIRST_GEN)
        lane_ops = max(B.int_peak(which, 20000)[0] for _ in range(6))      # best of 6: the first runs see ramping clocks
        return {"which": which, "cell_ops": ops, "lane_ops": lane_ops, "gcups": lane_ops / ops * 2 / 1e9}
    peak = measured_peak(args.mode in ("global", "glocal"))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def resident(mode_name, steps, warmup, sample_clocks):
        """device-resident job of this rank's shard: `warmup` untimed + `steps` timed steps (L2 flushed before each, one
        gather per step at N > 1), CUDA events on the launch stream, max over ranks"""
        sh = ShardedBestHits(queries, targets, rank, world, mode=MODE_NAMES[mode_name])
        result = None
        for _ in range(warmup):
            flush.zero_()
            sh.step(dist if world > 1 else None)
        torch.cuda.synchronize(); barrier()
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prof = {}
        torch.cuda.synchronize(); barrier()
        ev0.record()
        for _ in range(steps):                                          # back to back: no host synchronisation inside the timed region
            flush.zero_()
            result = sh.step(dist if world > 1 else None)
        ev1.record()
        torch.cuda.synchronize(); barrier()
        for row in sh.job.profile_history(steps):                       # per-launch CUDA-event times of the timed steps (the last 32 at most)
            p = prof.setdefault((row["K"], row["lanes"]), {"ms": 0.0, "cells": row["cells"], "queries": row["queries"], "n": 0})
            p["ms"] += row["ms"]; p["n"] += 1
        out = {"clocks": sampler.stop() if sampler else None, "elapsed_ms": max_over_ranks(ev0.elapsed_time(ev1)),
               "launches": sh.job.kernel_launches() * steps, "prof": prof, "steps": steps, "warmup": warmup}
        if rank == 0:
            out["score"], out["idx"] = result[0].cpu().numpy(), result[1].cpu().numpy()
        sh.close()
        return out

    def roofline_of(run, mode_name, pk, io_bytes):
        """the dominant packed-kernel launch of this rank against the measured rate of its own instruction mix"""
        if not run["prof"]:
            return None
        # the launch with the most cells: it goes first, on the caller's stream, so its event pair times it alone (the launches of
        # the other strip widths run on side streams and fill its tail; their event pairs include the time they wait for SMs)
        (K, lanes), p = max(run["prof"].items(), key=lambda kv: kv[1]["cells"])
        avg_ms = p["ms"] / p["n"]
        achieved = p["cells"] / (avg_ms * 1e-3) / 1e9
        streaming = mode_name in ("global", "glocal")
        return {"bound": "int_alu", "kernel": ("mm16g_kernel<%s,K=%d>" if streaming else "mm16_kernel<%s,K=%d>") % (mode_name, K) + (" (border lane: 31 x %d columns)" % K if lanes == 31 else ""),
                "achieved": achieved, "peak": pk["gcups"], "unit": "GCUPS", "frac": achieved / pk["gcups"],
                "peak_source": "measured in this run by abv_int_peak(%d): register-only chains of the kernel's cell mix (%d lane-ops per cell pair: %s + VIMNMX3.S16x2 + 2 VIADDMNMX.S16x2), %.3e lane-ops/s" % (
                    pk["which"], pk["cell_ops"], "1 add" if pk["cell_ops"] == 4 else "2 adds", pk["lane_ops"]),
                "avg_launch_ms": avg_ms, "cells_per_launch": p["cells"],
                "share_of_step": avg_ms / max(1e-9, run["elapsed_ms"] / run["steps"]),
                "launch_overlap": "packed launches of one step run concurrently (largest first); share_of_step = this launch's duration / step time",
                "launches_timed": p["n"], "rank": 0,
                "traffic": ncu_traffic(mode_name, K, lanes) if (world == 1 and args.workload == "config4" and not (args.queries or args.targets)) else None,
                "hbm": hbm_info(p["cells"], p["queries"], len(targets), total_cells / world, io_bytes, avg_ms)}

    with torch.cuda.stream(stream):
        # ---------------- device-resident arm: `value` -------------------------------------------
        main = resident(args.mode, args.steps, args.warmup, True)

        # ---------------- end-to-end arm: host buffers in, host results out ------------------------
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        qbuf, qoff = B.pack(queries[lo:hi])
        tbuf, toff = B.pack(targets)
        pin = lambda a: torch.from_numpy(a).pin_memory()
        p_q, p_qo, p_t, p_to = pin(qbuf), pin(qoff), pin(tbuf), pin(toff)
        h2d = int(qbuf.nbytes + qoff.nbytes + tbuf.nbytes + toff.nbytes)
        d2h = 8 * len(queries)

        def e2e_step():
            job = B.BestHitsJob((p_q.numpy(), p_qo.numpy()), (p_t.numpy(), p_to.numpy()), mode, B.SEED_GLOBAL)
            ds = torch.empty(max(hi - lo, 1), dtype=torch.int32, device=dev)[:hi - lo]
            di = torch.empty(max(hi - lo, 1), dtype=torch.int32, device=dev)[:hi - lo]
            job.launch(stream.cuda_stream, ds.data_ptr(), di.data_ptr())
            out = gather_best_hits(ds, di, bounds, dist, rank, world) if world > 1 else (ds, di)
            host = (out[0].cpu(), out[1].cpu()) if rank == 0 else None    # device->host read of the result
            stream.synchronize()
            job.close()
            return host
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            host = e2e_step()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()

        # ---------------- the same job in the mode BASELINE.json's config names (glocal), at every N ----------------
        glocal = None
        if args.mode == "global" and not args.no_glocal:
            glocal = resident("glocal", max(1, args.glocal_steps), 2, False)

    if rank == 0:
        gpu_score, gpu_idx = main["score"], main["idx"]
        assert (host[0].numpy() == gpu_score).all() and (host[1].numpy() == gpu_idx).all(), "end-to-end results differ from the resident job's"
        elapsed_ms = main["elapsed_ms"]
        value = total_cells * args.steps / (elapsed_ms * 1e-3) / 1e9
        line = {"metric": "hashtable-build alignment throughput", "value": value, "unit": "GCUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "s16x2 (int32-exact)", "data": "synthetic",
                "config": {"workload": name, "cells_per_step": total_cells, "l2": "flushed between steps (256 MiB memset)",
                           "sharding": "queries by contiguous range over %d GPU(s), targets replicated, one NCCL gather" % world},
                "hashtable_build_s": elapsed_ms / args.steps / 1e3,
                "gcups_per_gpu": value / world,
                "e2e": {"value": total_cells * args.steps / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "s_per_step": e2e_s / args.steps},
                "gpu_launches": main["launches"], "clocks": main["clocks"],
                "roofline": roofline_of(main, args.mode, peak, h2d + d2h),
                "result_sha256": result_digest(gpu_score, gpu_idx)}
        if glocal:
            gms = glocal["elapsed_ms"] / glocal["steps"]
            line["glocal"] = {"mode": "glocal", "value": total_cells / (gms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": gms,
                              "steps": glocal["steps"], "warmup": glocal["warmup"], "n_gpus": world, "gpu_launches": glocal["launches"],
                              "roofline": roofline_of(glocal, "glocal", peak, h2d + d2h),
                              "result_sha256": result_digest(glocal["score"], glocal["idx"]),
                              "note": "same workload and sharding, glocal_align semantics (glocal_align.c:73-202), one gather per step; the pipeline itself scores with global_align (ambivert.py:157-199)"}
        if world == 1 and args.merge_pairs > 0:
            line["merge"] = merge_measurement(args.merge_pairs)
            line["cluster"] = cluster_measurement(args.merge_pairs)
            line["variants"] = variants_measurement(100000)
        if world == 1 and not args.no_pipeline:
            n_q0, n_t0 = WORKLOADS["config4"]
            line["pipeline"] = pipeline_measurement(args.queries or n_q0, args.targets or n_t0, args.pipeline_gpus or B.device_count())
        if not args.no_cpu_baseline:
            # the reference's CPU path on a sample drawn from ALL shards, beside the GPU's results for the same queries:
            # N = 1: the bounded (~--cpu-seconds) sample that also gives the CPU rate; N > 1: >= 64 queries, parity only
            arm = cpu_arm(queries, targets, mode, args.cpu_seconds)
            pick = parity_sample(bounds, arm["n"] if world == 1 else max(64, 8 * world))
            sec, cells, bs, bi = arm["run"]([queries[i] for i in pick])
            line["cpu_baseline"] = {"value": cells / sec / 1e9, "unit": "GCUPS", "cores": arm["threads"], "kind": arm["kind"],
                                    "sample": "%d of %d queries, evenly drawn from all %d shard(s) incl. each shard's first and last query, x all %d targets (%.2e cells, %.1f s); %s" % (
                                        len(pick), len(queries), world, len(targets), cells, sec, arm["build"]),
                                    "matches_gpu": bool((bs == gpu_score[pick]).all() and (bi == gpu_idx[pick]).all())}
            if world == 1:
                line["cpu_baseline"]["as_shipped"] = as_shipped_measurement(queries, targets, gpu_score, gpu_idx)
            if glocal:
                garm = cpu_arm(queries, targets, MODE_NAMES["glocal"], 4.0)
                gpick = parity_sample(bounds, max(64, 8 * world))
                gsec, gcells, gbs, gbi = garm["run"]([queries[i] for i in gpick])
                line["glocal"]["cpu_baseline"] = {"value": gcells / gsec / 1e9, "unit": "GCUPS", "cores": garm["threads"], "kind": garm["kind"],
                                                  "sample": "%d queries drawn from all %d shard(s) x all %d targets, glocal_align_raw" % (len(gpick), world, len(targets)),
                                                  "matches_gpu": bool((gbs == glocal["score"][gpick]).all() and (gbi == glocal["idx"][gpick]).all())}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else a library prints (e.g. NCCL's version
    banner) was redirected to stderr in main()"""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # stray prints of native libraries go to stderr
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    cuda_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
