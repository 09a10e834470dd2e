#!/usr/bin/env python3
"""bench.py — k-mers counted per second on B200, against the HBM roofline, with the CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path (extract → count → sorted (key,count) table in HBM) over one batch
of synthetic input.  N=1 runs BASELINE.json configs[1] ("cfg2": 1e9 bases, 400-base records, k=21
canonical); N>1 is weak scaling: every rank holds its own shard of that size, routes its keys to their
owner GPU (hash prefix) with one NCCL all-to-all, and counts the keys it owns.  One JSON line on rank 0.

`value`      device-timed, inputs resident in HBM when the clock starts.
`e2e`        the same job through the public host call with the input in pinned HOST memory: H2D of the
             bases + offsets and D2H of the result summary (n_distinct, n_total) inside the timed region;
             the table digest is compared with the device-resident run's after the clock stops.
`roofline`   dominant kernel: algorithmic bytes / its CUDA-event time, vs MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` the CPU oracle (a C restatement — the Rust reference cannot be built here) on a bounded
             prefix of the same input, all host cores.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (bases per GPU, record length, k, canonical, description)
    "cfg2": dict(bases=1_000_000_000, rec_len=400, k=21, canonical=True,
                 desc="synthetic FASTA, 1e9 bases (2.5M records x 400), i.i.d. ACGT, k=21 canonical, u64 keys"),
    "cfg3": dict(bases=1_250_000_000, rec_len=400, k=31, canonical=True,
                 desc="synthetic FASTA, 1.25e9 bases per GPU (10e9 at 8 GPUs), i.i.d. ACGT, k=31 canonical, u64 keys"),
    "cfg4": dict(bases=500_000_000, rec_len=0, k=63, canonical=True,
                 desc="synthetic FASTA, 5e8 bases per GPU, read length U[100,10000], N-runs, k=63 canonical, u128 keys"),
    # BASELINE.json configs[4]: 1e11 bases over 8 GPUs; 150-base reads from both strands of a 1 Mbase genome with 5 %
    # homopolymer / tandem repeats (hot keys).  Not a default bench line (and its e2e leg is skipped: 12.5 GB of pinned
    # host memory per rank); set KMC_DIST_COMBINE=1 for the count-locally-then-exchange-rows route.
    "cfg5": dict(bases=12_500_000_000, rec_len=150, k=31, canonical=True, genome=1_000_000,
                 desc="synthetic FASTA, 1.25e10 bases per GPU (1e11 at 8 GPUs), 150-base reads from a 1 Mbase genome "
                      "with 5 % repeats, k=31 canonical, u64 keys, low cardinality"),
}


def hbm_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth(torch, n_bases, rec_len, seed, device, n_runs=False):
    """Seeded synthetic reads, generated on the device in chunks (no 1e9-element int64 temporaries)."""
    g = torch.Generator(device=device).manual_seed(seed)
    bases = torch.empty(n_bases, dtype=torch.uint8, device=device)
    CH = 1 << 27
    for s in range(0, n_bases, CH):
        e = min(n_bases, s + CH)
        c = torch.randint(0, 4, (e - s,), device=device, generator=g, dtype=torch.uint8)
        # 0,1,2,3 → 'A','C','G','T' = 65,67,71,84
        bases[s:e] = 65 + 2 * c + 2 * (c == 2).to(torch.uint8) + 13 * (c == 3).to(torch.uint8)
    if rec_len > 0:
        off = torch.arange(0, n_bases + 1, rec_len, dtype=torch.int64, device=device)
        if int(off[-1]) != n_bases:
            off = torch.cat([off, torch.tensor([n_bases], dtype=torch.int64, device=device)])
    else:  # cfg4: read length U[100,10000], N-runs of geometric length (mean 50) starting w.p. 1e-4 per base
        lens = torch.randint(100, 10001, (n_bases // 100 + 1,), device=device, generator=g, dtype=torch.int64)
        off = torch.cumsum(lens, 0)
        off = off[off < n_bases]
        off = torch.cat([torch.zeros(1, dtype=torch.int64, device=device), off,
                         torch.tensor([n_bases], dtype=torch.int64, device=device)])
    if n_runs:
        n_starts = max(1, int(n_bases * 1e-4))
        starts = torch.randint(0, n_bases, (n_starts,), device=device, generator=g, dtype=torch.int64)
        u = torch.rand(n_starts, device=device, generator=g).clamp_min(1e-9)
        run = (torch.log(u) / -0.02).long().clamp(1, 2000)  # geometric-like, mean 50
        for j in range(int(run.max())):
            idx = starts[run > j] + j
            bases[idx[idx < n_bases]] = 78
    return bases, off


def synth_genome(torch, n_bases, read_len, genome_len, seed, device):
    """cfg5: reads sampled uniformly from both strands of a fixed random genome whose first 5 % is poly-A and an (AC)n
    tandem repeat.  The genome is the same on every rank (seed 5); the read positions depend on `seed`."""
    gg = torch.Generator(device=device).manual_seed(5)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    codes = torch.randint(0, 4, (genome_len,), device=device, generator=gg)
    rep = genome_len // 20
    codes[:rep // 2] = 0
    codes[rep // 2:rep] = torch.arange(rep - rep // 2, device=device) % 2
    both = torch.cat([lut[codes], lut[(3 - codes).flip(0)]])
    g = torch.Generator(device=device).manual_seed(seed)
    n_reads = n_bases // read_len
    out = torch.empty(n_reads * read_len, dtype=torch.uint8, device=device)
    ar = torch.arange(read_len, device=device)
    CH = 1 << 20
    for s in range(0, n_reads, CH):
        e = min(n_reads, s + CH)
        st = torch.randint(0, genome_len - read_len, (e - s,), device=device, generator=g)
        strand = torch.randint(0, 2, (e - s,), device=device, generator=g) * genome_len
        out[s * read_len:e * read_len] = both[((st + strand)[:, None] + ar[None, :])].reshape(-1)
    off = torch.arange(0, n_reads * read_len + 1, read_len, dtype=torch.int64, device=device)
    return out, off


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline(orc, bases_np, off_np, k, canonical, threads):
    t0 = time.perf_counter()
    tab = orc.contiguous_mt(bases_np, off_np, k, canonical, threads=threads)
    dt = time.perf_counter() - t0
    return tab.n_total / dt / 1e9, dt, tab


def run_reference(args, wl):
    """--impl reference: the reference's CPU path.  The Rust binary cannot be built in this image (no
    cargo/rustc, crates.io deps), so this times the CPU oracle — its C restatement — with all host cores,
    each step on a bounded prefix of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import orc
    orc.build()
    cores = os.cpu_count() or 1
    sample = min(wl["bases"], args.cpu_sample)
    rng = np.random.default_rng(2)
    bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, sample)]
    rec_len = wl["rec_len"] or 5000
    off = np.arange(0, sample + 1, rec_len, dtype=np.uint64)
    if int(off[-1]) != sample:
        off = np.append(off, np.uint64(sample))
    for _ in range(args.warmup):
        cpu_baseline(orc, bases[:sample // 4], off[off <= sample // 4], wl["k"], wl["canonical"], cores)
    times, n_total = [], 0
    for _ in range(args.steps):
        v, dt, tab = cpu_baseline(orc, bases, off, wl["k"], wl["canonical"], cores)
        times.append(dt)
        n_total = tab.n_total
    ms = 1e3 * sum(times) / len(times)
    value = n_total / (ms / 1e3) / 1e9
    sample_txt = f"first {sample:.3g} bases of the workload shape per step (numpy-seeded i.i.d. ACGT)"
    print(json.dumps({
        "impl": "reference", "metric": "k-mers counted/sec", "value": value, "unit": "Gk/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64" if wl["k"] <= 32 else "u128", "data": "synthetic",
        "config": {"workload": args.workload + ": " + wl["desc"], "k": wl["k"], "canonical": wl["canonical"]},
        "cpu_baseline": {"value": value, "unit": "Gk/s", "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": "Gk/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--bases", type=float, default=0, help="override bases per GPU (development)")
    ap.add_argument("--strategy", type=int, default=0)
    ap.add_argument("--cpu-sample", type=float, default=2e8, help="bases in the CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.cpu_sample = int(args.cpu_sample)
    wl = dict(WORKLOADS[args.workload])
    if args.bases:
        wl["bases"] = int(args.bases)
    if args.impl == "reference":
        return run_reference(args, wl)

    import numpy as np
    import torch
    import kmer_count_b200 as K

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libkmc has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = os.environ.get("KMC_NCCL_DEBUG", "WARN")  # keep NCCL's banner off stdout: one JSON line
        dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("KMC_KERNEL_TIMING", "1")
    K.build()

    n = wl["bases"]
    seed = 2 + 1000 * rank
    if "genome" in wl:
        bases, off = synth_genome(torch, n, wl["rec_len"], wl["genome"], seed, dev)
        n = bases.numel()
    else:
        bases, off = synth(torch, n, wl["rec_len"], seed, dev, n_runs=(args.workload == "cfg4"))
    n_recs = off.numel() - 1
    torch.cuda.synchronize()
    first_mib = hashlib.sha256(bases[: 1 << 20].cpu().numpy().tobytes()).hexdigest()

    from kmer_count_b200.dist import DistCounter
    stream = torch.cuda.current_stream()
    dc = DistCounter(k=wl["k"], canonical=wl["canonical"], strategy=args.strategy, device=local,
                     world=world, rank=rank, dist=dist, torch=torch)
    dc.set_stream(stream.cuda_stream)

    def step_device():
        dc.reset()
        dc.submit_device(bases.data_ptr(), off.data_ptr(), n, n_recs)
        return dc.finish()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # nvidia-smi needs ~100 ms to produce its first row: start it before the warm-up steps
    for _ in range(args.warmup):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    kstats, totals, step_wall = [], None, []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        totals = step_device()
        step_wall.append(round(1e3 * (time.perf_counter() - t_s), 2))
        kstats.append(dc.stats())
    ev1.record(stream)
    barrier()
    if rank == 0 and not sampler.rows:
        time.sleep(0.15)  # very short runs: let the sampler deliver at least one row (GPU still warm)
    clocks = sampler.stop() if rank == 0 else None
    ms_local = ev0.elapsed_time(ev1) / args.steps
    digest = dc.digest()

    skip_e2e = n > 4_000_000_000   # cfg5: 12.5 GB of pinned host memory per rank is not a reasonable thing to ask for
    if not skip_e2e:
        # ---- e2e: pinned host input → result summary on the host, every step
        hb = torch.empty(n, dtype=torch.uint8).pin_memory()
        ho = torch.empty(n_recs + 1, dtype=torch.int64).pin_memory()
        hb.copy_(bases)
        ho.copy_(off)
        torch.cuda.synchronize()
        hb_np, ho_np = hb.numpy(), ho.numpy().view(np.uint64)

        def step_e2e():
            dc.reset()
            dc.submit_host(hb_np, ho_np)
            return dc.finish()          # (n_distinct, n_total) read back from the device: the step's result on the host

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_res = step_e2e()
        barrier()
        e2e_ms_local = 1e3 * (time.perf_counter() - t0) / args.steps
        # the table the last e2e step left in HBM must be the one the device-resident steps produced (checked outside the
        # timed region: the digest is one more pass over the 11 GB table, ~2 ms, and not part of producing the result)
        e2e_res = (*e2e_res, dc.digest())
    else:
        e2e_ms_local = float("inf")
        e2e_res = (totals[0], totals[1], digest)
    # whole-job digests (sum over ranks mod 2^64: which rank owns which keys may differ between the two runs)
    dg = torch.tensor(np.array([digest, e2e_res[2]], np.uint64).view(np.int64), device=dev)
    if dist is not None:
        dist.all_reduce(dg, op=dist.ReduceOp.SUM)
    dg = dg.cpu().numpy().view(np.uint64)
    assert int(dg[0]) == int(dg[1]), "e2e and device-resident runs disagree"
    digest_all = int(dg[0])

    t = torch.tensor([ms_local, e2e_ms_local], dtype=torch.float64, device=dev)
    cnt = torch.tensor([totals[1], totals[0]], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, e2e_ms = float(t[0]), float(t[1])
    all_walls = [step_wall]
    if dist is not None and os.environ.get("KMC_BENCH_STEPS") == "1":
        all_walls = [None] * world
        dist.all_gather_object(all_walls, step_wall)
    n_total, n_distinct = int(cnt[0]), int(cnt[1])

    if rank == 0:
        peak, peak_src = hbm_peak()
        wide = wl["k"] > 32
        W = 16 if wide else 8
        dn = n_distinct / max(1, n_total)
        b_alg = 1 + (W + 4) + 4 + (W + 4) * dn  # bytes per k-mer occurrence
        value = n_total / (ms / 1e3) / 1e9
        pipe_gbs = n_total / world * b_alg / (ms / 1e3) / 1e9
        # dominant kernel of this rank's last steps (CUDA events inside libkmc, on the launching stream)
        agg = {}
        for st in kstats:
            for name, v in st.get("kernels", {}).items():
                a = agg.setdefault(name, [0, 0.0])
                a[0] += v["launches"]
                a[1] += v["ms"]
        per_step_total = sum(v[1] for v in agg.values()) / args.steps if agg else 0.0
        dom = max(agg.items(), key=lambda kv: kv[1][1]) if agg else (None, [0, 0.0])
        kernel_bytes = dc.algorithmic_bytes(dom[0], n_total // world, n_distinct // world, n, W) if dom[0] else None
        dom_ms = dom[1][1] / max(1, dom[1][0])
        launches_per_step = dom[1][0] / args.steps
        roof = {"bound": "hbm", "kernel": dom[0], "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                "achieved": (kernel_bytes / launches_per_step / (dom_ms / 1e3) / 1e9) if kernel_bytes else None,
                "avg_launch_ms": dom_ms, "launches_per_step": launches_per_step,
                "share_of_step": (dom[1][1] / args.steps) / per_step_total if per_step_total else None,
                "traffic": None,
                "pipeline": {"b_alg_per_kmer": b_alg, "achieved": pipe_gbs, "frac": pipe_gbs / peak}}
        roof["frac"] = roof["achieved"] / peak if roof["achieved"] else None
        prof = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(prof):
            roof["traffic"] = json.load(open(prof)).get(dom[0] or "", None)
        out = {
            "metric": "k-mers counted/sec", "value": value, "unit": "Gk/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u128" if wide else "u64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "k": wl["k"], "canonical": wl["canonical"],
                       "bases_per_gpu": n, "records_per_gpu": n_recs, "seed": seed, "first_mib_sha256": first_mib,
                       "l2": "inputs (>= 1 GB) larger than the 126 MB L2; no explicit flush",
                       "strategy": kstats[-1].get("strategy_used"), "parallelism": f"{getattr(dc, 'path', None) or 'single'}-partition x{world}"},
            "n_total": n_total, "n_distinct": n_distinct, "digest": digest_all,
            "roofline": roof,
            "e2e": None if skip_e2e else {"value": n_total / (e2e_ms / 1e3) / 1e9, "unit": "Gk/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(n + 8 * (n_recs + 1)), "d2h_bytes_per_step": 16,
                    "note": "kmc_submit_host (pinned) + kmc_finish (n_distinct, n_total read back; libkmc also reads ~40 KB "
                            "of histogram and cursors for its plan); "
                            "wall clock, max over ranks; table digest compared with the device-resident run afterwards"},
            "gpu_launches": int(sum(st["kernel_launches"] for st in kstats)),
            "kernels_ms_per_step": {kname: v[1] / args.steps for kname, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
            "phases_ms": kstats[-1].get("phases_ms"),
            "clocks": clocks,
            "step_wall_ms": all_walls if os.environ.get("KMC_BENCH_STEPS") == "1" else step_wall,
        }
        if not args.no_cpu:
            from oracle import orc  # cpu_baseline leg only
            orc.build()
            cores = os.cpu_count() or 1
            sample = min(n, args.cpu_sample)
            sb = bases[:sample].cpu().numpy()
            so = off[off <= sample].cpu().numpy().astype(np.uint64)
            if int(so[-1]) != sample:
                so = np.append(so, np.uint64(sample))
            v, dt, tab = cpu_baseline(orc, sb, so, wl["k"], wl["canonical"], cores)
            out["cpu_baseline"] = {"value": v, "unit": "Gk/s", "cores": cores, "kind": "port", "seconds": dt,
                                   "sample": f"first {sample:.3g} bases of rank 0's input, one pass, oracle/liborc.so "
                                             f"orc_contiguous_mt (C restatement; the Rust reference cannot be built here)"}
        print(json.dumps(out))
    dc.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
