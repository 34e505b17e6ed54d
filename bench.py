#!/usr/bin/env python
"""bench.py -- headline benchmark of the SINA hot path on B200.

    python bench.py [--gpus N --steps K --warmup W]            our CUDA path
    python bench.py --impl reference [--steps K --warmup W]    the reference's CPU code on the host cores

Workload (BASELINE.json configs[1]): 10 000 full-length 16S queries (~1 500 nt) against a synthetic
SILVA-like reference MSA of 50 000 sequences x 50 000 columns, reference defaults (k=10 fast, --fs-max 40).
One "step" = the whole hot path (k-mer family finding -> family graph -> mesh DP -> backtrack -> gap
placement) over the step's queries. `value` = sequences/s with the queries already resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (sg_run_batch: H2D of the queries, D2H of the aligned
columns inside the timed region). Under torchrun each rank owns one GPU, holds a replica of the index and
aligns its own shard of queries (weak scaling, no collective on the data path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_REFS, W_COLS, L_REF, KMER = 50000, 50000, 1500, 10
CHUNK = int(os.environ.get("SG_BATCH", "2368"))   # queries per graph/DP/backtrack launch (the library's default)
CHUNK_ISO = 2368                                 # launch size of the kernel-only pass behind `roofline`: the production chunk (SG_BATCH)
SEED = 20260117


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=10000, help="queries per step per GPU")
    ap.add_argument("--refs", type=int, default=N_REFS)
    ap.add_argument("--kind", default="full", choices=["full", "v4"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --queries per step on every GPU; strong: --total-queries per step split over the GPUs")
    ap.add_argument("--total-queries", type=int, default=0, help="strong scaling: queries per step over all GPUs")
    ap.add_argument("--no-shares", dest="shares", action="store_false",
                    help="skip the sub-objects measured by default on one GPU: one GPU's share of BASELINE configs[2] (V4) and "
                         "configs[3] (full-length) against a 500k-row reference, each with its own parity / roofline / "
                         "cpu_baseline (they add 2-3 minutes)")
    a = ap.parse_args()
    if a.scaling == "strong":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        total = a.total_queries or a.queries
        a.total_queries = total
        a.queries = max(1, total // world)
    return a


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_data(args, rank):
    from sina_b200 import synth
    tree, m, c, o = synth.synth_msa(args.refs, W=W_COLS, L=L_REF, seed=SEED)
    qm, qo = synth.synth_queries(tree, args.queries, args.kind, seed=1000 + rank)
    return tree, m, c, o, qm, qo


class _NS:
    """attribute bag (a variant of the command-line arguments for a share)"""
    def __init__(self, **kw):
        self.__dict__.update(kw)


class CpuPath:
    """The reference's own CPU code (oracle/_ref, kind 'reference') or, where that binary is absent, the C port
    (kind 'port'), with its database and k-mer index built once. run() times the whole path (family finding +
    graph + DP + backtrack + gap placement) over a bounded sample of queries on all host cores."""

    def __init__(self, args, m, c, o):
        from oracle import oracle as O
        self.O, self.args = O, args
        self.msa = O.MSA(m, c, o, W_COLS)
        self.kind = "reference" if os.path.exists(O.REF_SO) else "port"
        if self.kind == "reference":
            self.ref = O.Ref()
            self.db = self.ref.db(self.msa)
            self.ix = self.ref.kidx_build(self.db, KMER, 0)
        else:
            self.orc = O.Oracle()
            self.ix = self.orc.index_build(self.msa, KMER, 0)

    def run(self, qm, qo, nsample, nthreads=0):
        O = self.O
        nsample = min(nsample, len(qo) - 1)
        sub_off = (qo[:nsample + 1] - qo[0]).astype(np.uint64)
        sub_m = qm[int(qo[0]):int(qo[nsample])]
        if self.kind == "reference":
            queries = [O.decode(sub_m[int(sub_off[i]):int(sub_off[i + 1])]) for i in range(nsample)]
            t0 = time.perf_counter()
            res, oc, qoff, cells, posts, nt = self.ref.run_batch(self.ix, queries, O.FamParams(), O.AlignParams(), nthreads=nthreads)
            dt = time.perf_counter() - t0
        else:
            t0 = time.perf_counter()
            res, oc, om, cells, posts, nt = self.orc.run_batch(self.ix, self.msa, sub_m, sub_off, O.FamParams(), O.AlignParams(), nthreads=nthreads)
            dt = time.perf_counter() - t0
        # k-mer search alone on one core (the reference's find() is serial per query), same algorithmic-bytes formula
        # as roofline_kmer: 4*P + 2*N + 8*max per query
        kfind = None
        try:
            nf = min(64, nsample)
            t1 = time.perf_counter()
            P = 0
            for i in range(nf):
                q = sub_m[int(sub_off[i]):int(sub_off[i + 1])]
                if self.kind == "reference":
                    P += self.ref.find(self.ix, O.decode(q), 41)[2]
                else:
                    P += self.orc.find(self.ix, q, 41)[2]
            dtf = time.perf_counter() - t1
            kfind = {"gbs": (4.0 * P + (2.0 * self.args.refs + 8 * 41) * nf) / dtf / 1e9, "queries_per_s": nf / dtf,
                     "cores": 1, "sample": "%d queries, find(max=41) only" % nf}
        except Exception:
            pass
        self.last = {"res": res, "cols": oc, "qoff": qoff if self.kind == "reference" else sub_off, "n": nsample}
        return {"kmer_search": kfind, "value": nsample / dt, "unit": "sequences/s", "cores": int(nt), "kind": self.kind,
                "sample": "%d of the step's %s queries vs the same %d-row index, whole path, %.1f s on %d threads"
                          % (nsample, self.args.kind, self.args.refs, dt, int(nt)),
                "mcells_per_s": cells / dt / 1e6, "seconds": dt}

    def parity(self, qo, oc_gpu, res_gpu):
        """the sample just run on the CPU against the GPU's output for the same queries: status, number of bases, every
        alignment column, head / tail / quality, and the DP score bit for bit"""
        L = self.last
        mism, first = 0, None
        for i in range(L["n"]):
            r, g = L["res"][i], res_gpu[i]
            ok = int(r.status) == int(g["status"])
            if ok and int(r.status) in (0, 1):
                a, b = int(L["qoff"][i]), int(qo[i] - qo[0])
                n = int(g["n_out"])
                ok = bool((L["cols"][a:a + n] == oc_gpu[b:b + n]).all())
            if ok and int(r.status) == 0:
                ok = (np.float32(r.score).view(np.uint32) == np.float32(g["score"]).view(np.uint32)
                      and (int(r.head), int(r.tail), int(r.qual)) == (int(g["head"]), int(g["tail"]), int(g["qual"])))
            if not ok:
                mism += 1
                first = i if first is None else first
        out = {"checked": L["n"], "mismatches": mism, "against": self.kind,
               "what": "status, aligned columns of every base, head/tail/quality, score bits"}
        if first is not None:
            out["first_mismatch_query"] = first
        return out

    def close(self):
        if self.kind == "reference":
            self.ref.kidx_free(self.ix)
            self.ref.db_free(self.db)
        else:
            self.orc.index_free(self.ix)


def cpu_sample_size(args):
    """bounded sample: about 10-30 s of CPU work for the bench's own arm, a few seconds per step for --impl reference"""
    if args.cpu_sample:
        return args.cpu_sample
    per_core = 6.0 if args.kind == "full" else 60.0   # measured order of magnitude, sequences/s per host core
    target_s = 12.0 if args.impl == "ours" else 4.0
    return int(max(16, min(args.queries, per_core * (os.cpu_count() or 1) * target_s)))


def run_reference(args):
    """--impl reference: the reference's CPU implementation timed on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tree, m, c, o, qm, qo = make_data(args, 0)
    nsample = cpu_sample_size(args)
    cpu = CpuPath(args, m, c, o)
    times, last = [], None
    for it in range(args.warmup + args.steps):
        # each step = a bounded sample of the workload (different queries every step)
        a = (it * nsample) % max(1, args.queries - nsample)
        sub = cpu.run(qm, qo[a:], nsample)
        if it >= args.warmup:
            times.append(sub["seconds"])
            last = sub
    cpu.close()
    t = float(np.mean(times))
    val = nsample / t
    line = {"impl": "reference", "metric": "sequences aligned/sec", "value": val, "unit": "sequences/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, nsample, sample_of=args.queries),
            "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": last["cores"], "kind": last["kind"],
                             "sample": last["sample"]},
            "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, queries_per_step, sample_of=None):
    if args.refs == N_REFS and args.kind == "full":
        which = "BASELINE configs[1]"
    elif args.refs == 500000 and args.kind == "v4":
        which = "BASELINE configs[2] (one GPU's share)"
    elif args.refs == 500000 and args.kind == "full":
        which = "BASELINE configs[3] (one GPU's share)"
    else:
        which = "variant of BASELINE configs[1]"
    return {"workload": "%s: %d %s 16S queries (~%d nt) per step per GPU vs synthetic %d-seq reference MSA (%d columns), "
                        "k=%d fast, fs-max 40, reference defaults" % (which, queries_per_step, "full-length" if args.kind == "full" else "V4",
                                                                     1500 if args.kind == "full" else 280, args.refs, W_COLS, KMER),
            "queries_per_step_per_gpu": queries_per_step, "refs": args.refs, "columns": W_COLS, "k": KMER,
            **({"bounded_sample": "each step of this arm is a bounded sample of %d of the workload's %d queries per step (same index, same "
                                  "options): a rate on the same configuration, not a smaller configuration" % (queries_per_step, sample_of)} if sample_of else {}),
            **({"strong_scaling_total_queries_per_step": args.total_queries} if getattr(args, "scaling", "weak") == "strong" else {}),
            "l2": "inputs larger than L2 (index 0.4 GB + >30 GB traceback written per step)",
            "timing": "value: CUDA events on the library's own stream around the K steps (sg_session_timer), max over "
                      "ranks; e2e: host clock between device-wide synchronisations (host copies are part of it); per-stage "
                      "times are CUDA events on the launching streams and overlap across the chunk pipeline",
            "parallelism": "queries sharded over GPUs, index replicated, no collective"}


def measure(mods, ix, m, c, o, qm, qo, wargs, steps, warmup, rank, world, local, cpu=True):
    """One workload on this rank's GPU: device-resident number (CUDA events on the library's stream), end-to-end number
    through the host-buffer C-ABI call, kernel-only pass for the roofline, and on rank 0 of a 1-GPU run the CPU baseline
    with the parity check of its sample against the GPU's output. Returns the JSON line (rank 0) or None."""
    torch, dist, sina_b200 = mods
    nq = wargs.queries
    fp, ap = sina_b200.FamParams(), sina_b200.AlignParams()
    sess = sina_b200.Session(ix, nq, int(qo[-1]))
    sess.upload(qm, qo)  # queries resident in HBM before the timed region

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        sess.run(fp, ap)   # famfinder + aligner in one call (sg_session_run), as sg_run_batch runs them
        sess.sync()

    # the caller's output buffers, allocated once as a C++ host keeps them (plain pageable memory)
    e2e_out = (np.zeros(len(qm), np.uint32), np.zeros(len(qm), np.uint8), np.zeros(nq, sina_b200.RESULT_DTYPE))

    def step_e2e():
        return ix.run(qm, qo, fp, ap, out=e2e_out)

    # ---- device-resident number
    for _ in range(warmup):
        step_device()
    sess.stats(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    sess.timer_start()          # CUDA event on the library's stream (torch.cuda.Event only sees torch's streams)
    for _ in range(steps):
        step_device()
    dt = sess.timer_stop() / 1e3  # device clock between the two events: every kernel of the K steps lies inside
    barrier()
    dt_host = time.perf_counter() - t0
    st = sess.stats()
    # ---- end-to-end through the host-buffer C-ABI call
    for _ in range(min(warmup, 1) or 1):
        oc, om, res = step_e2e()
    barrier()
    t1 = time.perf_counter()
    for _ in range(steps):
        oc, om, res = step_e2e()
    barrier()
    dt_e2e = time.perf_counter() - t1
    sampler.stop_flag = True
    sampler.join(timeout=2)
    n_ok = int((res["status"] == 0).sum())

    # ---- kernel-only pass: one workspace / one stream, so that the CUDA events bracket every kernel alone
    # (in the timed region above the chunks of several workspaces overlap and the per-stage event times include
    # whatever ran beside them)
    sess.close()
    os.environ["SG_STREAMS"] = "1"
    user_batch = os.environ.get("SG_BATCH")
    os.environ["SG_BATCH"] = str(CHUNK_ISO)   # whole waves: 2368 = 4 x (148 SMs x 4 resident CTAs), the production chunk
    os.environ["SG_FIRST_DIV"] = "1"          # ... and no short first launch
    nq_iso = min(nq, 3 * CHUNK_ISO)
    iso = sina_b200.Session(ix, nq_iso, int(qo[nq_iso]))
    iso.upload(qm[:int(qo[nq_iso])], qo[:nq_iso + 1])
    iso.family(fp)
    iso.align(ap)
    iso.sync()
    iso.stats(reset=True)
    iso.family(fp)
    iso.align(ap)
    iso.sync()
    st_iso = iso.stats()
    iso.close()
    os.environ.pop("SG_STREAMS", None)
    os.environ.pop("SG_BATCH", None)
    os.environ.pop("SG_FIRST_DIV", None)
    if user_batch is not None:
        os.environ["SG_BATCH"] = user_batch

    if world > 1:
        tt = torch.tensor([dt, dt_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        return None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    total_q = nq * steps * world
    value = total_q / dt
    # dominant kernel: mesh DP (mesh_kernel). Algorithmic bytes = 1 B packed traceback per cell;
    # duration = CUDA events around the kernel on its stream in the kernel-only pass, this rank.
    dp_s = st_iso["ms_dp"] / 1e3
    cells_iso = float(st_iso["cells"])
    gcups = cells_iso / dp_s / 1e9 if dp_s > 0 else 0.0
    cells = float(st["cells"])
    sm_mhz = sampler.summary().get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    ops_per_cell = 3 + 7 * 1.65
    issue_ceiling = 148 * 128 * sm_mhz * 1e6 / ops_per_cell / 1e9  # GCUPS at the measured clock
    find_s = st_iso["ms_find"] / 1e3
    posts = float(st_iso["postings"])
    kmer_bytes = 4.0 * posts + (2.0 * wargs.refs + 8.0 * 41) * nq_iso
    traffic, kmer_traffic = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj["mesh_kernel"]["dram_bytes_per_query"] * min(CHUNK_ISO, nq_iso) / 1e9   # GB per launch of one full chunk
        kmer_traffic = tj.get("find_tile_kernel", {}).get("dram_bytes_per_query_by_refs", {}).get(str(wargs.refs))
    except Exception:
        pass
    launches_iso = max(1, -(-nq_iso // CHUNK_ISO))
    iso_sum = sum(st_iso[k] for k in ("ms_find", "ms_family", "ms_graph", "ms_dp", "ms_backtrack"))
    line = {
        "metric": "sequences aligned/sec", "value": value, "unit": "sequences/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": wargs.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wargs, nq),
        "e2e": {"value": total_q / dt_e2e, "unit": "sequences/s", "h2d_bytes_per_step": int(len(qm) + qo.nbytes) * world,
                "d2h_bytes_per_step": int(oc.nbytes + om.nbytes + res.nbytes) * world},
        "gpu_launches": int(st["kernel_launches"]),
        "roofline": {"kernel": "mesh_kernel", "bound": "hbm", "achieved": gcups * 1.0,
                     "peak": hbm_peak, "unit": "GB/s", "frac": gcups / hbm_peak,
                     "traffic": traffic, "traffic_unit": "GB per launch (%d-query chunk), ncu dram read+write" % CHUNK_ISO,
                     "kernel_only_pass": "%d queries in launches of %d on one stream (the timed region runs %d-query launches on %s streams, where the kernel's events overlap other kernels)" % (nq_iso, CHUNK_ISO, CHUNK, os.environ.get("SG_STREAMS", "3")),
                     "algorithmic_gb_per_launch": cells_iso / launches_iso / 1e9,
                     "peak_source": peak_src,
                     "limiter": "not HBM: 1 B of traceback per cell is the only mandatory HBM traffic, so `frac` is low by construction "
                                "(the contract's `bound` only knows hbm / tensor). The kernel is bound by instruction issue on the "
                                "ALU pipe (fp32 min / compare / select) and the per-step barrier: judge it by frac_issue",
                     "gcups": gcups, "issue_ceiling_gcups": issue_ceiling,
                     "frac_issue": gcups / issue_ceiling if issue_ceiling else None,
                     "duration_ms_per_launch": st_iso["ms_dp"] / launches_iso,
                     "share_of_step_isolated": st_iso["ms_dp"] / iso_sum if iso_sum else None},
        "roofline_kmer": {"kernel": "find_tile_kernel", "bound": "hbm", "achieved": kmer_bytes / find_s / 1e9 if find_s > 0 else 0.0,
                          "peak": hbm_peak, "unit": "GB/s", "frac": (kmer_bytes / find_s / 1e9 / hbm_peak) if find_s > 0 else 0.0,
                          "bytes_per_query": "4*P + 2*N + 8*max (SURVEY §8d: u32 postings, one pass over the int16 score vector)",
                          "postings_per_query": posts / nq_iso,
                          "traffic": kmer_traffic, "traffic_unit": "bytes per query, ncu dram read+write (the index stores u16 postings and the counters live in shared memory)"},
        "stages_ms_per_step_isolated": {k: st_iso[k] * (nq / nq_iso) for k in ("ms_find", "ms_family", "ms_graph", "ms_dp", "ms_backtrack")},
        "cells_per_query": cells / (nq * steps), "aligned_ok": n_ok, "ms_per_step_host_clock": dt_host / steps * 1e3,
        "clocks": sampler.summary(),
    }
    if world == 1 and cpu:
        cpu_path = CpuPath(wargs, m, c, o)
        cb = cpu_path.run(qm, qo, cpu_sample_size(wargs))
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "mcells_per_s", "kmer_search")}
        # parity of the timed configuration: the CPU sample's alignments against the GPU's for the same queries
        line["parity"] = cpu_path.parity(qo, oc, res)
        cpu_path.close()
    return line


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import sina_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or sina_b200.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: sina_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL_DEBUG=VERSION (set on some boxes) makes NCCL print a banner to stdout, in front of the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mods = (torch, dist, sina_b200)

    tree, m, c, o, qm, qo = make_data(args, rank)
    ix = sina_b200.Index(m, c, o, W_COLS, k=KMER, device=local)
    line = measure(mods, ix, m, c, o, qm, qo, args, args.steps, args.warmup, rank, world, local, cpu=not args.no_cpu_baseline)
    ix.close()

    if args.shares and world == 1:
        # one GPU's share of BASELINE configs[3] (100k full-length over 8 GPUs) and configs[2] (1M V4 over 8 GPUs), both
        # against a 500k-row reference: own index, own parity / roofline / cpu_baseline
        from sina_b200 import synth
        del m, c, o, qm, qo
        refs = 500000
        tree, m, c, o = synth.synth_msa(refs, W=W_COLS, L=L_REF, seed=SEED)
        ix = sina_b200.Index(m, c, o, W_COLS, k=KMER, device=local)
        shares = {}
        for name, kind, nq in (("configs[3]", "full", 12500), ("configs[2]", "v4", 125000)):
            w = _NS(**vars(args))
            w.refs, w.kind, w.queries, w.scaling = refs, kind, nq, "weak"
            w.cpu_sample = args.cpu_sample or (256 if kind == "full" else 1024)
            qm, qo = synth.synth_queries(tree, nq, kind, seed=1000)
            shares[name] = measure(mods, ix, m, c, o, qm, qo, w, max(1, min(args.steps, 2)), 1, rank, world, local,
                                   cpu=not args.no_cpu_baseline)
        ix.close()
        line["shares"] = shares
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
