#!/usr/bin/env python
"""bench.py - SLIMM profiling hot path on B200: mapped SAM records/s through
coverage -> filter -> reassign -> LCA -> profile, and the achieved fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5|cfg2|cfg3|cfg4] [--impl reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE).  A step is one pass of the whole hot
path over the workload's records:
  * `value`  : records already resident in HBM when the timed region starts (slimm_gpu_push_device),
  * `e2e`    : the same pass through the C ABI with HOST (pinned) buffers, H2D copies of the three
               record arrays and D2H of the results inside the timed region,
  * `roofline`: the kernel with the largest share of the step; its algorithmic bytes /
               its CUDA-event duration / measured HBM peak (MEASURED_PEAKS.json),
  * `cpu_baseline`: the unmodified reference binary (oracle/_ref/slimm) on a bounded sample, 1 core.
`--impl reference` times that reference binary as its own arm.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs
WORKLOADS = {
    "cfg2": dict(desc="synthetic 1k-genome DB, 10M records, 20% multi-mapped, bin width 1000", G=1000, N=10_000_000,
                 w=1000, cc=0.95, multi_frac=0.2, k_lo=2, k_hi=8, neigh=8),
    "cfg3": dict(desc="synthetic 10k-genome DB, 100M records, 40% multi-mapped, bin width 1000", G=10_000,
                 N=100_000_000, w=1000, cc=0.95, multi_frac=0.4, k_lo=2, k_hi=8, neigh=8),
    "cfg4": dict(desc="LCA stress: 50k genomes, reads on 2..64 references, -cc 1.0", G=50_000, N=100_000_000, w=1000,
                 cc=1.0, multi_frac=1.0, k_lo=2, k_hi=64, neigh=64),
    "cfg5": dict(desc="1B records over 50k references at bin width 100, 20% multi-mapped", G=50_000,
                 N=1_000_000_000, w=100, cc=0.95, multi_frac=0.2, k_lo=2, k_hi=8, neigh=8),
}
AVG_READ_LEN = 100
SEED = 12345


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def make_community(wl):
    """Contigs + taxonomy + lineage table on the host (identical on every rank)."""
    from slimm_b200 import synth
    rng = np.random.default_rng(SEED)
    tax, accs = synth.make_taxonomy(wl["G"])
    contigs = synth.make_contigs(wl["G"], rng, accs)
    db = synth.database_for(tax)
    lineage = db.lineage_table(contigs.accessions)
    return tax, contigs, db, lineage


def profile_tail(api, gpu, contigs, lineage, taxa_arrays, cc):
    """D2H of the stage results + host rank aggregation (the 'profile' end of the path)."""
    rows, n_rows = gpu.profile_raw(1, 0.01)      # slimm_gpu_profile: result readback + rank aggregation
    return gpu.summary(), n_rows


def run_reference_sample(wl, n_sample, steps, warmup, tmp_root=None, with_cli=False):
    """Times the UNMODIFIED reference binary (oracle/_ref/slimm) on a bounded sample of the workload.
    Returns (records/s median over steps, seconds per step list, sample description, kind)."""
    from slimm_b200 import sldb, synth
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "slimm")
    rng = np.random.default_rng(SEED)
    tax, accs = synth.make_taxonomy(wl["G"])
    contigs = synth.make_contigs(wl["G"], rng, accs)
    rec = synth.make_records(contigs, n_sample, np.random.default_rng(SEED + 1), multi_frac=wl["multi_frac"],
                             k_lo=wl["k_lo"], k_hi=wl["k_hi"], neigh=wl["neigh"])
    td = tempfile.mkdtemp(prefix="slimm_ref_", dir=tmp_root)
    try:
        if os.path.exists(ref_bin):
            kind = "reference"
            sam = os.path.join(td, "in.sam")
            synth.write_sam_for_records(sam, contigs, rec)
            dbp = os.path.join(td, "db.sldb")
            sldb.write_sldb(synth.database_for(tax), dbp)
            out = os.path.join(td, "out") + "/"
            os.makedirs(out)
            cmd = [ref_bin, "-w", str(wl["w"]), "-cc", str(wl["cc"]), "-o", out, dbp, sam]
            times = []
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                r = subprocess.run(cmd, capture_output=True, text=True)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    raise RuntimeError("reference slimm failed: " + r.stderr[-400:])
                if i >= warmup:
                    times.append(dt)
            sample = (f"{rec.read_id.size} records of the same generator (G={wl['G']}, w={wl['w']}), whole slimm process "
                      f"wall time incl. SAM decode, DB load and bin init, single-threaded binary")
            if with_cli:
                CLI_RESULT.clear()
                CLI_RESULT.update(run_cli_sample(cmd, ref_bin, out, rec.read_id.size, statistics.median(times)))
        else:
            import oracle
            kind = "port"
            lineage = synth.database_for(tax).lineage_table(contigs.accessions)
            times = []
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                oracle.run(contigs.lengths, lineage, wl["w"], AVG_READ_LEN, wl["cc"], rec.read_id, rec.ref_id, rec.begin_pos)
                dt = time.perf_counter() - t0
                if i >= warmup:
                    times.append(dt)
            sample = f"{rec.read_id.size} records, oracle C port on pre-decoded SoA (oracle/_ref/slimm not built)"
    finally:
        shutil.rmtree(td, ignore_errors=True)
    med = statistics.median(times)
    return rec.read_id.size / med, times, sample, kind


CLI_RESULT = {}


def run_cli_sample(ref_cmd, ref_bin, out_dir, n_records, ref_seconds):
    """The drop-in command line (slimm_b200/bin/slimm: threaded SAM decoder -> pinned batches -> the C ABI -> TSV) on the SAME
    SAM file, database and options the reference binary was just timed on; the decode rate is reported separately."""
    import re
    cli = os.path.join(ROOT, "slimm_b200", "bin", "slimm")
    if not os.path.exists(cli):
        return {"unavailable": "slimm_b200/bin/slimm not built"}
    ref_profile = os.path.join(out_dir, "in_profile.tsv")
    def plain_rows(path):         # rows of taxa proper ("<parent>*" / "0*" rows carry order-dependent f32 sums)
        return sorted(l for l in open(path).read().splitlines() if not l.split("\t")[1:2] or not l.split("\t")[1].endswith("*"))
    ref_rows = plain_rows(ref_profile) if os.path.exists(ref_profile) else None
    cmd = [cli, "-v"] + ref_cmd[1:]
    best, err = None, ""
    for _ in range(2):            # the first run pays CUDA context creation and page-in of the library
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            return {"unavailable": "slimm_b200/bin/slimm failed: " + r.stderr[-200:]}
        best, err = dt if best is None else min(best, dt), r.stderr
    m = re.search(r"decode: (\d+) records, (\d+) reads in ([0-9.e+-]+) s \(([0-9.e+-]+) M records/s, (\d+) host threads\); GPU stages: ([0-9.e+-]+) ms", err)
    res = {"seconds": best, "records_per_s": n_records / best, "reference_seconds": ref_seconds,
           "what": "whole process wall time of slimm_b200/bin/slimm on the SAM file of cpu_baseline.sample (process start, CUDA "
                   "context, .sldb load, threaded SAM decode, pinned uploads, GPU stages, _profile.tsv)"}
    if m:
        res.update({"decode_M_records_per_s": float(m.group(4)), "decode_host_threads": int(m.group(5)), "gpu_stages_ms": float(m.group(6))})
    if ref_rows is not None and os.path.exists(ref_profile):
        res["profile_rows_equal_to_reference"] = plain_rows(ref_profile) == ref_rows
    return res


def reference_sample_size(wl, budget_s):
    # reference speed is about 0.25 M records/s plus a fixed bin-initialisation cost that grows with the
    # number of bins (3 histograms of len/w+1 u32 per contig)
    bins = 3.5e6 * wl["G"] / wl["w"]
    fixed = 3 * bins * 4 / 2.0e9
    n = int(max(100_000, min(5_000_000, (budget_s - fixed) * 250_000)))
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SLIMM_BENCH_WORKLOAD", "cfg5"), choices=sorted(WORKLOADS))
    ap.add_argument("--records", type=int, default=0, help="override the workload's record count (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--bins", default=os.environ.get("SLIMM_BENCH_BINS", "keep"), choices=["keep", "skip"],
                    help="keep: the cov/uniq_cov bins are written back to HBM (fetchable, as -co/-ro need); skip: profile-only run")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    wl = dict(WORKLOADS[args.workload])
    if args.records:
        wl["N"] = args.records
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    base = {"metric": "mapped SAM records/s through coverage->filter->reassign->LCA->profile", "unit": "records/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic (seeded generator, SURVEY.md 8(d))",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "n_refs": wl["G"], "records": wl["N"],
                       "bin_width": wl["w"], "cov_cut_off": wl["cc"], "sharding": f"by read over {world} GPU(s)",
                       "l2": "inputs larger than L2 (no flush needed)" if wl["N"] * 12 > 400e6 else
                             "512 MB scratch write between steps (outside the timed events)"}}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        budget = 150.0 / max(1, args.steps + args.warmup)
        n_sample = reference_sample_size(wl, budget)
        rps, times, sample, kind = run_reference_sample(wl, n_sample, args.steps, args.warmup)
        line = dict(base)
        line.update({"impl": "reference", "value": rps, "ms_per_step": 1e3 * statistics.median(times), "n_gpus": world,
                     "warmup": args.warmup, "cpu_baseline": {"value": rps, "unit": "records/s", "cores": 1, "kind": kind,
                                                             "sample": sample},
                     "e2e": {"value": rps, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    from slimm_b200 import api, synth_torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    tax, contigs, db, lineage = make_community(wl)
    taxa_arrays = api.taxa_arrays({t: v for t, v in db.taxid__name.items()})
    n_local = wl["N"] // world + (1 if rank < wl["N"] % world else 0)
    recs = synth_torch.make_records_device(contigs.lengths, contigs.weights, n_local, dev, seed=SEED + 17 * rank,
                                           multi_frac=wl["multi_frac"], k_lo=wl["k_lo"], k_hi=wl["k_hi"], neigh=wl["neigh"])
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    gpu = api.SlimmGpu(contigs.lengths, lineage, wl["w"], AVG_READ_LEN, device=local_rank,
                       flags=api.SKIP_BINS if args.bins == "skip" else 0)
    base["config"]["bins"] = ("written back to HBM (fetchable)" if args.bins == "keep" else
                              "consumed in shared memory, not written back (profile-only run, SLIMM_GPU_SKIP_BINS)")
    gpu.set_stream(stream.cuda_stream)
    gpu.enable_timing(True)
    gpu.set_taxa(taxa_arrays)
    flush = None if wl["N"] * 12 > 400e6 else torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    if world > 1:
        from slimm_b200 import dist as sdist
        gpu.set_shard(rank, world)
        # items travel as peer-to-peer stores inside the split kernel (CUDA IPC over NVLink); NCCL all-to-all otherwise
        use_p2p = os.environ.get("SLIMM_BENCH_P2P", "1") != "0" and sdist.connect_peers(gpu, dev, wl["N"])
        base["config"]["exchange"] = "peer-to-peer stores fused into the split kernel" if use_p2p else "NCCL all-to-all of the items"
    phases = {}   # SLIMM_BENCH_PHASES=1: CUDA-event durations of the exchange steps, summed over all steps (adds a sync per step)

    def hot_path():
        """coverage -> filter -> assign -> profile; with several GPUs the items are routed to the rank that owns their
        histogram slice (one NCCL all-to-all, 4 B/record) and only per-reference statistics and the assign block are
        summed over ranks (slimm_b200/dist.py)."""
        if world > 1:
            gpu.set_shard(rank, world)
            sdist.run_sharded(gpu, dev, wl["cc"], 0, wl["N"], phase_ms=phases if os.environ.get("SLIMM_BENCH_PHASES") else None)
        else:
            gpu.coverage()
            gpu.filter(wl["cc"], 0)
            gpu.assign()
        return profile_tail(api, gpu, contigs, lineage, taxa_arrays, wl["cc"])

    def step_resident():
        gpu.reset()
        gpu.push_device(recs.read_id.data_ptr(), recs.ref_id.data_ptr(), recs.begin_pos.data_ptr(), recs.n)
        return hot_path()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, n_warm, n_steps, collect_kernel=False):
        for _ in range(n_warm):
            step_fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        kt = []
        barrier()
        for a, b in ev:
            if flush is not None:
                flush.fill_(1)
            a.record(stream)
            out = step_fn()
            b.record(stream)
            if collect_kernel:
                kt.append(gpu.timings())
        barrier()
        ms = [a.elapsed_time(b) for a, b in ev]
        total = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item()), out, kt

    launches0 = gpu.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms, (summ, rows), ktimes = timed(step_resident, warmup, args.steps, collect_kernel=True)
    clocks = sampler.stop()
    launches = (gpu.launch_count() - launches0) // (warmup + args.steps) * args.steps
    ms_per_step = total_ms / args.steps
    value = wl["N"] / (ms_per_step * 1e-3)

    # roofline: algorithmic bytes per kernel as in DESIGN.md section 3 (N*(16+16) + P*8 + U*8 + 16*B),
    # for this rank's share; the kernel with the largest share of the step is the one reported
    peak, peak_src = measured_peak_gbs()
    P, U, B = summ.n_pairs / world, summ.uniq_matches_count / world, summ.n_bins / world
    kernel_ms = {k: statistics.mean(t[k] for t in ktimes) for k in ktimes[0]}
    bucketed = kernel_ms["accumulate"] > 0
    alg = {"coverage": 16.0 * n_local + (0.0 if bucketed else 8.0 * P + 8.0 * U + 8.0 * B),
           "accumulate": 8.0 * P + 8.0 * U + 8.0 * B, "stats": 8.0 * B, "assign": 16.0 * n_local}
    if bucketed and kernel_ms.get("stats", 0) == 0:   # the per-reference scan ran inside the accumulate stage (fine slices in shared memory)
        alg["accumulate"] += 8.0 * B
    names = {"coverage": "k_coverage", "accumulate": "k_fine_count + k_fine_split + k_fine_accumulate (bins and per-reference scan in shared memory)"
             if kernel_ms.get("stats", 0) == 0 else "k_accumulate (+ histogram memset on the side stream)",
             "stats": "k_ref_stats", "assign": "k_assign_reads"}
    dom = max((k for k in alg if kernel_ms.get(k, 0) > 0), key=lambda k: kernel_ms[k])
    traffic = None
    try:   # DRAM bytes per record of each kernel from the committed ncu capture (profiles/), scaled to this launch
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.workload in tr and dom in tr[args.workload]:
            traffic = tr[args.workload][dom]["dram_bytes_per_record"] * n_local
    except Exception:
        pass
    pipe_bytes = 32.0 * n_local + 8.0 * P + 8.0 * U + 16.0 * B
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": alg[dom] / (kernel_ms[dom] * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": alg[dom] / (kernel_ms[dom] * 1e-3) / 1e9 / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom], "kernel_ms": kernel_ms[dom],
                "per_kernel": {k: {"ms": kernel_ms[k], "algorithmic_bytes": alg[k],
                                   "frac": (alg[k] / (kernel_ms[k] * 1e-3) / 1e9 / peak) if kernel_ms[k] > 0 else None}
                               for k in alg},
                "pipeline": {"algorithmic_bytes_per_step": pipe_bytes * world, "ms_per_step": ms_per_step,
                             "achieved": pipe_bytes / (ms_per_step * 1e-3) / 1e9,
                             "frac": pipe_bytes / (ms_per_step * 1e-3) / 1e9 / peak, "kernel_ms": kernel_ms}}

    # end to end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        h = [torch.empty(recs.n, dtype=torch.int32, pin_memory=True) for _ in range(3)]
        for dst, src in zip(h, (recs.read_id, recs.ref_id, recs.begin_pos)):
            dst.copy_(src)
        torch.cuda.synchronize()
        del recs
        torch.cuda.empty_cache()

        def step_e2e():
            gpu.reset()
            gpu.push_ptrs(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[0].numel())
            return hot_path()

        e_total_ms, _, _ = timed(step_e2e, 1, max(1, min(args.steps, 3)))
        e_steps = max(1, min(args.steps, 3))
        d2h = 10 * wl["G"] * 4 + 96   # per-taxon aggregates of two ranks + scalars (slimm_gpu_profile)
        e2e = {"value": wl["N"] / (e_total_ms / e_steps * 1e-3), "unit": "records/s",
               "h2d_bytes_per_step": 12 * h[0].numel(), "d2h_bytes_per_step": d2h, "steps": e_steps,
               "note": "slimm_gpu_push from pinned host SoA (3 x u32 per record) + all stages + result readback"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            gpu.close()                   # the command line below creates its own context on this GPU
            torch.cuda.empty_cache()
            rps, times, sample, kind = run_reference_sample(wl, reference_sample_size(wl, 25.0), 1, 0, with_cli=True)
            cpu = {"value": rps, "unit": "records/s", "cores": 1, "kind": kind, "sample": sample,
                   "host_cores_available": os.cpu_count()}
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "records/s", "cores": 1, "kind": "unavailable", "sample": str(e)[:200]}

    if rank == 0:
        if phases:
            sys.stderr.write("phases (ms, summed over all resident + e2e steps incl. warm-up): " + json.dumps({k: round(v, 2) for k, v in phases.items()}) + "\n")
        line = dict(base)
        line.update({"value": value, "ms_per_step": ms_per_step, "e2e": e2e, "gpu_launches": int(launches),
                     "roofline": roofline, "cpu_baseline": cpu, "cli": dict(CLI_RESULT) or None, "clocks": clocks,
                     "result": {"hits": summ.hits_count, "reads": summ.matches_count, "uniq": summ.uniq_matches_count,
                                "uniq2": summ.uniq_matches_count2, "valid_refs": summ.n_valid, "rows": rows,
                                "pairs": summ.n_pairs, "bins": summ.n_bins, "sorted_input": summ.input_was_sorted}})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
