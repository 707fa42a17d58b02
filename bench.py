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
    "cfg4": dict(desc="LCA stress: 50k genomes, reads on 2..64 references drawn inside the primary's species..phylum, -cc 1.0",
                 G=50_000, N=100_000_000, w=1000, cc=1.0, multi_frac=1.0, k_lo=2, k_hi=64, neigh=64, neigh_mode="taxonomy"),
    "cfg5": dict(desc="1B records over 50k references at bin width 100, 20% multi-mapped", G=50_000,
                 N=1_000_000_000, w=100, cc=0.95, multi_frac=0.2, k_lo=2, k_hi=8, neigh=8),
}
AVG_READ_LEN = 100
SEED = 12345
N_BLOCKS = 64            # the records are generated block by block (seed per block): rank r of n takes blocks [64 r/n, 64 (r+1)/n),
READ_ID_STRIDE = 1 << 24  # so the dataset - and therefore the result - is the same whatever the number of GPUs
PREFIX_BLOCKS = 8        # sharded_equals_single: the single-GPU path is re-run on this many blocks (cfg5: 125 M records)


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


def run_reference_sample(wl, n_sample, steps, warmup, tmp_root=None, with_cli=False, n_small=0):
    """Times the UNMODIFIED reference binary (oracle/_ref/slimm) on a bounded sample of the workload: `steps` runs on
    `n_sample` records (timed) after `warmup` runs on `n_small` records (the same generator's prefix; they double as the
    second point of the fixed + per-record cost fit).
    Returns (records/s median over steps, seconds per step list, sample description, kind, fit dict or None)."""
    from slimm_b200 import sldb, synth
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "slimm")
    rng = np.random.default_rng(SEED)
    tax, accs = synth.make_taxonomy(wl["G"])
    contigs = synth.make_contigs(wl["G"], rng, accs)
    rec = synth.make_records(contigs, n_sample, np.random.default_rng(SEED + 1), multi_frac=wl["multi_frac"],
                             k_lo=wl["k_lo"], k_hi=wl["k_hi"], neigh=wl["neigh"], neigh_mode=wl.get("neigh_mode", "index"))
    td = tempfile.mkdtemp(prefix="slimm_ref_", dir=tmp_root)
    fit = None
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

            def run_once(c):
                t0 = time.perf_counter()
                r = subprocess.run(c, capture_output=True, text=True)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    raise RuntimeError("reference slimm failed: " + r.stderr[-400:])
                return dt

            small_times = []
            if warmup and n_small and n_small < rec.read_id.size:
                cut = int(np.searchsorted(rec.read_id, rec.read_id[n_small], side="left"))   # whole reads only
                small = synth.Records(rec.read_id[:cut], rec.ref_id[:cut], rec.begin_pos[:cut], int(rec.read_id[cut - 1]) + 1)
                sam_s = os.path.join(td, "small.sam")
                synth.write_sam_for_records(sam_s, contigs, small)
                small_times = [run_once(cmd[:-1] + [sam_s]) for _ in range(warmup)]
                n_small = cut
            else:
                for _ in range(warmup):
                    run_once(cmd)
            times = [run_once(cmd) for _ in range(steps)]
            sample = (f"{rec.read_id.size} records of the same generator (G={wl['G']}, w={wl['w']}), whole slimm process "
                      f"wall time incl. SAM decode, DB load and bin init, single-threaded binary")
            if small_times:
                t_big, t_small = statistics.median(times), statistics.median(small_times)
                per_rec = max((t_big - t_small) / (rec.read_id.size - n_small), 1e-12)
                fit = {"records": [int(n_small), int(rec.read_id.size)], "seconds": [t_small, t_big],
                       "marginal_records_per_s": 1.0 / per_rec, "fixed_cost_s": max(t_small - n_small * per_rec, 0.0),
                       "fixed_share_of_timed_run": max(t_small - n_small * per_rec, 0.0) / t_big,
                       "what": "two sample sizes of the same generator: time = fixed (database load, 3 zeroed histograms of "
                               "len/w+1 bins per contig, reference src/slimm.hpp:430-445) + records / marginal rate"}
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
    return rec.read_id.size / med, times, sample, kind, fit


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


def reference_fixed_cost_s(wl):
    # the reference zero-fills 3 histograms of len/w+1 u32 bins per contig (src/slimm.hpp:430-445) and loads the database:
    # about 0.6 s per GB of bins on this class of host (measured: 10.5 s at cfg5's 21 GB) plus ~0.2 ms per contig
    bins = 3.5e6 * wl["G"] / wl["w"]
    return 3 * bins * 4 / 2.0e9 + 2e-5 * wl["G"]


def reference_sample_size(wl, budget_s, fixed_share=0.25):
    """Records for one run of the reference: fixed cost at most `fixed_share` of the run when the budget allows
    (marginal rate about 0.25 M records/s), never more than `budget_s` seconds."""
    fixed = reference_fixed_cost_s(wl)
    want = fixed * (1.0 - fixed_share) / fixed_share * 250_000          # fixed = share * (fixed + n / rate)
    cap = max(budget_s - fixed, 1.0) * 250_000
    return int(max(100_000, min(want, cap, 12_000_000)))


def block_range(rank, world):
    return N_BLOCKS * rank // world, N_BLOCKS * (rank + 1) // world


def records_of_block(wl, b):
    per = wl["N"] // N_BLOCKS
    return per + (wl["N"] - per * N_BLOCKS if b == N_BLOCKS - 1 else 0)


def make_blocks(wl, contigs, dev, b_lo, b_hi):
    """Blocks [b_lo, b_hi) of the workload's records on the device, concatenated (read ids ascend across blocks)."""
    import torch
    from slimm_b200 import synth_torch
    parts = []
    for b in range(b_lo, b_hi):
        parts.append(synth_torch.make_records_device(contigs.lengths, contigs.weights, records_of_block(wl, b), dev, seed=SEED + 1 + b,
                                                     multi_frac=wl["multi_frac"], k_lo=wl["k_lo"], k_hi=wl["k_hi"], neigh=wl["neigh"],
                                                     neigh_mode=wl.get("neigh_mode", "index"), read_id_base=b * READ_ID_STRIDE))
    if len(parts) == 1:
        return parts[0]
    out = synth_torch.DeviceRecords(torch.cat([p.read_id for p in parts]), torch.cat([p.ref_id for p in parts]),
                                    torch.cat([p.begin_pos for p in parts]), sum(p.n_reads for p in parts))
    del parts
    torch.cuda.empty_cache()
    return out


def result_fingerprint(gpu, rank_level=1, ac=0.01):
    """Everything the parity contract names, as comparable python values (called on one rank; D2H copies)."""
    s = gpu.summary()
    st = gpu.ref_stats()
    rows = gpu.profile(rank_level, ac)
    return {"summary": (s.hits_count, s.matches_count, s.uniq_matches_count, s.uniq_matches_count2, s.reference_count, s.n_valid,
                        s.failed_by_cov, s.failed_by_uniq_cov, s.failed_by_min_read, s.n_pairs,
                        np.float32(s.coverage_cut_off).tobytes(), np.float32(s.uniq_coverage_cut_off).tobytes()),
            "reads_count": st.reads_count.tobytes(), "uniq_reads_count": st.uniq_reads_count.tobytes(),
            "uniq_reads_count2": st.uniq_reads_count2.tobytes(), "nz_bins": st.nz_bins.tobytes(),
            "uniq_nz_bins": st.uniq_nz_bins.tobytes(), "valid": st.valid.tobytes(),
            "lca_counts": sorted(gpu.lca_counts().items()), "lca_children": gpu.lca_children().tobytes(),
            "profile_rows": [(r.taxon, r.kind, r.read_count, r.first_child, np.float64(r.abundance).tobytes()) for r in rows]}


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
    ap.add_argument("--no-equality-check", action="store_true", help="several GPUs: skip the single-GPU re-run of the prefix")
    ap.add_argument("--bins", default=os.environ.get("SLIMM_BENCH_BINS", "keep"), choices=["keep", "skip"],
                    help="keep: the cov/uniq_cov bins are written back to HBM (fetchable, as -co/-ro need); skip: profile-only run")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    wl = dict(WORKLOADS[args.workload])
    if args.records:
        wl["N"] = args.records
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    # one config dict for both arms (the driver compares them key by key)
    base = {"metric": "mapped SAM records/s through coverage->filter->reassign->LCA->profile", "unit": "records/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic (seeded generator, SURVEY.md 8(d); generated in 64 seeded blocks, "
                                                         "so the records are the same for every number of GPUs)",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "n_refs": wl["G"], "records": wl["N"],
                       "bin_width": wl["w"], "cov_cut_off": wl["cc"], "sharding": f"by read over {world} GPU(s)",
                       "l2": "inputs larger than L2 (no flush needed)" if wl["N"] * 12 > 400e6 * world else
                             "512 MB scratch write between steps (outside the timed events)",
                       "bins": ("written back to HBM (fetchable)" if args.bins == "keep" else
                                "consumed in shared memory, not written back (profile-only run, SLIMM_GPU_SKIP_BINS)"),
                       "exchange": "none (one GPU)" if world == 1 else
                                   "items grouped by histogram slice locally, every owner's contiguous block copied into its buffer over NVLink peer memory "
                                   "(k_peer_copy; NCCL all-to-all of the items when peer mapping is unavailable)"}}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        # timed steps: a sample whose fixed cost (bin initialisation + database load) is at most a quarter of a run, as long as
        # the whole arm stays below ~25 minutes; warm-up steps: a tenth of it (a CPU process needs no warm-up; they give the
        # second point of the fixed + per-record fit)
        n_runs = max(1, args.steps)
        budget = (1500.0 - args.warmup * (reference_fixed_cost_s(wl) + 6.0)) / n_runs
        n_sample = reference_sample_size(wl, budget)
        rps, times, sample, kind, fit = run_reference_sample(wl, n_sample, args.steps, args.warmup, n_small=max(20_000, n_sample // 10))
        line = dict(base)
        line.update({"impl": "reference", "value": rps, "ms_per_step": 1e3 * statistics.median(times), "n_gpus": world,
                     "warmup": args.warmup, "cpu_baseline": {"value": rps, "unit": "records/s", "cores": 1, "kind": kind,
                                                             "sample": sample, "fit": fit,
                                                             "host_cores_available": os.cpu_count()},
                     "e2e": {"value": rps, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    from slimm_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    tax, contigs, db, lineage = make_community(wl)
    taxa_arrays = api.taxa_arrays({t: v for t, v in db.taxid__name.items()})
    b_lo, b_hi = block_range(rank, world)
    recs = make_blocks(wl, contigs, dev, b_lo, b_hi)
    n_local = recs.n
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    gpu = api.SlimmGpu(contigs.lengths, lineage, wl["w"], AVG_READ_LEN, device=local_rank,
                       flags=api.SKIP_BINS if args.bins == "skip" else 0)
    gpu.set_stream(stream.cuda_stream)
    gpu.enable_timing(True)
    gpu.set_taxa(taxa_arrays)
    flush = None if wl["N"] * 12 > 400e6 * world else torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    use_p2p = False
    if world > 1:
        from slimm_b200 import dist as sdist
        gpu.set_shard(rank, world)
        # items travel as peer-to-peer stores inside the split kernel (CUDA IPC over NVLink); NCCL all-to-all otherwise
        use_p2p = os.environ.get("SLIMM_BENCH_P2P", "1") != "0" and sdist.connect_peers(gpu, dev, wl["N"])
    phases = {}   # SLIMM_BENCH_PHASES=1: CUDA-event durations of the exchange steps, summed over all steps (adds a sync per step)

    def hot_path(global_hits):
        """coverage -> filter -> assign -> profile; with several GPUs the items are routed to the rank that owns their
        histogram slice and only per-reference statistics and the assign block are summed over ranks (slimm_b200/dist.py)."""
        if world > 1:
            gpu.set_shard(rank, world)
            sdist.run_sharded(gpu, dev, wl["cc"], 0, global_hits, phase_ms=phases if os.environ.get("SLIMM_BENCH_PHASES") else None)
        else:
            gpu.coverage()
            gpu.filter(wl["cc"], 0)
            gpu.assign()
        return profile_tail(api, gpu, contigs, lineage, taxa_arrays, wl["cc"])

    def step_resident():
        gpu.reset()
        gpu.push_device(recs.read_id.data_ptr(), recs.ref_id.data_ptr(), recs.begin_pos.data_ptr(), recs.n)
        return hot_path(wl["N"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, n_warm, n_steps, collect_kernel=False):
        for _ in range(n_warm):
            step_fn()
        phases.clear()                                        # (SLIMM_BENCH_PHASES) the timed steps only
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
    phase_ms = {k: v / args.steps for k, v in phases.items()} if phases else None
    result = {"hits": summ.hits_count, "reads": summ.matches_count, "uniq": summ.uniq_matches_count,
              "uniq2": summ.uniq_matches_count2, "valid_refs": summ.n_valid, "rows": rows,
              "pairs": summ.n_pairs, "bins": summ.n_bins, "sorted_input": summ.input_was_sorted}

    per_rank = None
    if world > 1:                                             # every rank's kernel times (the line's own kernel_ms are rank 0's)
        keys = sorted(ktimes[0])
        mine = torch.tensor([statistics.mean(t[k] for t in ktimes) for k in keys], dtype=torch.float64, device=dev)
        allk = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allk, mine)
        per_rank = {k: [round(float(a[i]), 3) for a in allk] for i, k in enumerate(keys) if any(float(a[i]) > 0 for a in allk)}

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
    names = {"coverage": "k_coverage_tile", "accumulate": "k_fine_count + k_fine_split + k_fine_accumulate (bins and per-reference scan in shared memory)"
             if kernel_ms.get("stats", 0) == 0 else "k_accumulate (+ histogram memset on the side stream)",
             "stats": "k_ref_stats", "assign": "k_assign_reads"}
    # the dominant KERNEL: the fine-slice accumulate stage is six kernels (count, scan, split, packed / wide / cluster accumulate; the
    # largest of them about 2.5 ms at cfg5), so it is listed in per_kernel as a stage but never reported as "the dominant kernel"
    composite = {"accumulate"} if (bucketed and kernel_ms.get("stats", 0) == 0) else set()
    dom = max((k for k in alg if kernel_ms.get(k, 0) > 0 and k not in composite), key=lambda k: kernel_ms[k])
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
                "note": "frac = SURVEY 8(d) algorithmic bytes / CUDA-event time / measured copy peak; the assign stage is credited 16 B per "
                        "record by that formula while it only re-reads the compact stream of the multi-target reads (see profiles/)",
                "per_kernel": {k: {"ms": kernel_ms[k], "algorithmic_bytes": alg[k],
                                   "frac": (alg[k] / (kernel_ms[k] * 1e-3) / 1e9 / peak) if kernel_ms[k] > 0 else None,
                                   **({"what": "a stage of six kernels; with compact bins it writes 4 of the 8 algorithmic bytes per bin"}
                                      if k in composite else {})}
                               for k in alg},
                "pipeline": {"algorithmic_bytes_per_step": pipe_bytes * world, "ms_per_step": ms_per_step,
                             "achieved": pipe_bytes / (ms_per_step * 1e-3) / 1e9,
                             "frac": pipe_bytes / (ms_per_step * 1e-3) / 1e9 / peak, "kernel_ms": kernel_ms}}

    # end to end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region.  The records travel in the
    # wire format of grouped input (slimm_gpu_push_packed: one new-read bit + u16 reference id + i32 position = 6.125 bytes per
    # record, what the decoder emits for mapper output), pushed in 64 M-record batches so the unpack kernels hide under the copies
    e2e = None
    if not args.no_e2e:
        n = recs.n
        packed = wl["G"] <= 65536
        if packed:
            h_pos = torch.empty(n, dtype=torch.int32, pin_memory=True)
            h_ref = torch.empty(n, dtype=torch.int16, pin_memory=True)
            h_bits = torch.empty((n + 31) // 32, dtype=torch.int32, pin_memory=True)
            h_pos.copy_(recs.begin_pos)
            wts = (torch.ones(32, dtype=torch.int64, device=dev) << torch.arange(32, dtype=torch.int64, device=dev))
            step_rec = 1 << 26
            for a in range(0, n, step_rec):
                b = min(n, a + step_rec)
                h_ref[a:b].copy_((recs.ref_id[a:b] & 0xFFFF).to(torch.int16))
                new = torch.ones(b - a, dtype=torch.bool, device=dev)
                new[1:] = recs.read_id[a + 1:b] != recs.read_id[a:b - 1]
                if a:
                    new[0] = recs.read_id[a] != recs.read_id[a - 1]
                pad = (-(b - a)) % 32
                if pad:
                    new = torch.cat([new, torch.zeros(pad, dtype=torch.bool, device=dev)])
                words = (new.view(-1, 32).to(torch.int64) * wts).sum(1)
                h_bits[a // 32:a // 32 + words.numel()].copy_(torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32))
                del new, words
            h2d = 4 * n + 2 * n + 4 * h_bits.numel()
            batch = 1 << 26

            def step_e2e():
                gpu.reset()
                for a in range(0, n, batch):
                    m = min(batch, n - a)
                    gpu.push_packed_ptrs(h_bits.data_ptr() + a // 8, h_ref.data_ptr() + 2 * a, h_pos.data_ptr() + 4 * a, m)
                return hot_path(wl["N"])
            note = ("slimm_gpu_push_packed from pinned host buffers (new-read bit + u16 reference id + i32 position per record, 64 M-record "
                    "batches) + all stages + result readback")
        else:
            h = [torch.empty(n, dtype=torch.int32, pin_memory=True) for _ in range(3)]
            for dst, src in zip(h, (recs.read_id, recs.ref_id, recs.begin_pos)):
                dst.copy_(src)
            h2d = 12 * n

            def step_e2e():
                gpu.reset()
                gpu.push_ptrs(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), n)
                return hot_path(wl["N"])
            note = "slimm_gpu_push from pinned host SoA (3 x u32 per record) + all stages + result readback"
        torch.cuda.synchronize()
        e_steps = max(1, min(args.steps, 3))
        e_total_ms, (e_summ, _), _ = timed(step_e2e, 1, e_steps)
        d2h = 10 * wl["G"] * 4 + 96   # per-taxon aggregates of two ranks + scalars (slimm_gpu_profile)
        e2e = {"value": wl["N"] / (e_total_ms / e_steps * 1e-3), "unit": "records/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps, "note": note,
               "same_result_as_resident": (e_summ.matches_count, e_summ.uniq_matches_count2, e_summ.n_valid) ==
                                          (summ.matches_count, summ.uniq_matches_count2, summ.n_valid)}
        # what the link alone gives this rank while every rank copies (no kernels): the ceiling of e2e
        torch.cuda.synchronize()
        barrier()
        probe = torch.empty(min(n, 1 << 28), dtype=torch.int32, device=dev)
        src = (h_pos if packed else h[0])[:probe.numel()]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(3):
            probe.copy_(src, non_blocking=True)
        ev1.record(stream)
        torch.cuda.synchronize()
        e2e["h2d_probe_GBps_this_rank"] = 3 * 4 * probe.numel() / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
        if world > 1:                                             # every rank's link while all of them copy: where the host side saturates
            mine = torch.tensor([e2e["h2d_probe_GBps_this_rank"]], dtype=torch.float64, device=dev)
            allp = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allp, mine)
            e2e["h2d_probe_GBps_per_rank"] = [round(float(a[0]), 1) for a in allp]
            e2e["h2d_probe_GBps_sum"] = round(sum(float(a[0]) for a in allp), 1)
            try:
                e2e["cpu_affinity_this_rank"] = len(os.sched_getaffinity(0))
            except Exception:
                pass
        del probe
        if packed:
            del h_pos, h_ref, h_bits
        else:
            del h
    del recs
    torch.cuda.empty_cache()

    # several GPUs: the sharded run against this library's own single-GPU run on the same records (a prefix of the blocks)
    equality = None
    if world > 1 and not args.no_equality_check:
        n_pre = min(N_BLOCKS, max(PREFIX_BLOCKS, world)) if wl["N"] > 130_000_000 else N_BLOCKS
        pre_records = sum(records_of_block(wl, b) for b in range(n_pre))
        lo, hi = n_pre * rank // world, n_pre * (rank + 1) // world
        part = make_blocks(wl, contigs, dev, lo, hi)
        gpu.reset()
        gpu.set_shard(rank, world)
        gpu.push_device(part.read_id.data_ptr(), part.ref_id.data_ptr(), part.begin_pos.data_ptr(), part.n)
        sdist.run_sharded(gpu, dev, wl["cc"], 0, pre_records)
        sharded = result_fingerprint(gpu) if rank == 0 else None
        torch.cuda.synchronize()
        del part
        if rank == 0:
            allrec = make_blocks(wl, contigs, dev, 0, n_pre)
            single = api.SlimmGpu(contigs.lengths, lineage, wl["w"], AVG_READ_LEN, device=local_rank,
                                  flags=api.SKIP_BINS if args.bins == "skip" else 0)
            single.set_stream(stream.cuda_stream)
            single.set_taxa(taxa_arrays)
            single.push_device(allrec.read_id.data_ptr(), allrec.ref_id.data_ptr(), allrec.begin_pos.data_ptr(), allrec.n)
            single.run(wl["cc"], 0)
            ref = result_fingerprint(single)
            single.close()
            del allrec
            differ = [k for k in ref if ref[k] != sharded[k]]
            equality = {"sharded_equals_single": not differ, "records": pre_records, "blocks": n_pre, "fields": sorted(ref),
                        "differing_fields": differ, "lca_taxa": len(ref["lca_counts"]), "profile_rows": len(ref["profile_rows"])}
        barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            gpu.close()                   # the command line below creates its own context on this GPU
            torch.cuda.empty_cache()
            rps, times, sample, kind, _ = run_reference_sample(wl, reference_sample_size(wl, 25.0), 1, 0, with_cli=True)
            cpu = {"value": rps, "unit": "records/s", "cores": 1, "kind": kind, "sample": sample,
                   "host_cores_available": os.cpu_count(),
                   "note": "whole-process rate of a 25 s run: mostly the reference's fixed cost at this shape (bin initialisation); "
                           "`bench.py --impl reference` fits fixed and per-record cost from two sample sizes"}
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "records/s", "cores": 1, "kind": "unavailable", "sample": str(e)[:200]}

    if rank == 0:
        if phase_ms:
            sys.stderr.write("phases (ms per timed resident step, rank 0): " + json.dumps({k: round(v, 3) for k, v in phase_ms.items()}) + "\n")
        line = dict(base)
        line.update({"value": value, "ms_per_step": ms_per_step, "e2e": e2e, "gpu_launches": int(launches),
                     "roofline": roofline, "cpu_baseline": cpu, "cli": dict(CLI_RESULT) or None, "clocks": clocks,
                     "result": result, "exchange_used": ("p2p" if use_p2p else "nccl") if world > 1 else None,
                     "per_rank_kernel_ms": per_rank, "exchange_phases_ms": phase_ms})
        if equality is not None:
            line.update(equality)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
