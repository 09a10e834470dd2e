#!/usr/bin/env python3
"""bench.py — k-mers counted per second on B200, against the HBM roofline, with the CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path (extract → count → sorted (key,count) table in HBM) over one batch
of synthetic input.  N=1 runs BASELINE.json configs[1] ("cfg2": 1e9 bases, 400-base records, k=21
canonical); N>1 is weak scaling: every rank holds its own shard of that size, routes its keys to their
owner GPU (hash prefix) over NVLink, and counts the keys it owns.  One JSON line on rank 0.

`value`      device-timed, inputs resident in HBM when the clock starts.
`e2e`        the same job through the public host call with the input in pinned HOST memory: H2D of the
             bases + offsets and D2H of the result summary (n_distinct, n_total) inside the timed region;
             the table digest is compared with the device-resident run's after the clock stops.
`roofline`   frac = SURVEY.md §8d: B_alg = 17 + 12 D/N (25 + 20 D/N for 128-bit keys) bytes per k-mer x keys per GPU /
             step time, vs MEASURED_PEAKS.json hbm_gbs; kernel_frac = the dominant kernel's own algorithmic bytes / its
             CUDA-event time (measured inside libkmc on the launching stream).
`k31`        the same measurement on BASELINE.json configs[2]'s shape (k=31, 1.25e9 bases per GPU: 1e10 at 8 GPUs).
`cfg1`       (N=1) the reference's own job: sample.fasta, lr-gapped 27+27 — GPU against the CPU oracle, digests compared.
`parity`     after the timed region every rank counts a small shard through the same multi-GPU path and rank 0 compares
             (n_total, n_distinct, sum of the ranks' table digests) with the CPU oracle on the concatenated input.
`cpu_baseline` the CPU oracle (a C restatement — the Rust reference cannot be built here) on a bounded
             prefix of the same input bytes, all host cores.
Inputs come from the counter-based generator (kmc_gen_* on the device, k-mer-count_b200/gen.py on the host: the same
bytes), so both arms and every torch version see the same data.
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
    # host memory per rank).
    "cfg5": dict(bases=12_500_000_000, rec_len=150, k=31, canonical=True, genome=1_000_000,
                 desc="synthetic FASTA, 1.25e10 bases per GPU (1e11 at 8 GPUs), 150-base reads from a 1 Mbase genome "
                      "with 5 % repeats, k=31 canonical, u64 keys, low cardinality"),
}
PARITY_BASES = 10_000_000   # per rank, for the after-the-clock oracle comparison


def hbm_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def seed_of(name, rank):
    return {"cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}[name] + 1000 * rank


def host_input(gen, np, name, wl, rank, n_bases):
    """The first n_bases of `rank`'s input of workload `name`, on the host (gen.py: the same bytes the device generator
    writes) → (bases uint8, rec_off uint64)."""
    seed = seed_of(name, rank)
    if "genome" in wl:
        genome = gen.repeat_genome(5, wl["genome"])
        n_reads = n_bases // wl["rec_len"]
        return gen.reads(seed, genome, wl["rec_len"], 0, n_reads), np.arange(n_reads + 1, dtype=np.uint64) * wl["rec_len"]
    bases = gen.bases(seed, 0, n_bases)
    if wl["rec_len"] > 0:
        off = np.arange(0, n_bases + wl["rec_len"], wl["rec_len"], dtype=np.uint64).clip(max=n_bases)
        off = np.unique(off)
    else:
        gen.add_n_runs(seed, 0, bases)
        off = gen.read_offsets(seed, n_bases)
    return bases, off


def device_input(torch, np, gen, kc, name, wl, rank, n_bases, dev):
    """`rank`'s input of workload `name` generated on the device (kmc_gen_*) → (bases, rec_off) tensors."""
    seed = seed_of(name, rank)
    if "genome" in wl:
        genome = torch.from_numpy(gen.repeat_genome(5, wl["genome"])).to(dev)
        n_reads = n_bases // wl["rec_len"]
        bases = torch.empty(n_reads * wl["rec_len"], dtype=torch.uint8, device=dev)
        kc.gen_reads(seed, genome.data_ptr(), genome.numel(), wl["rec_len"], 0, n_reads, bases.data_ptr())
        off = torch.arange(0, n_reads + 1, dtype=torch.int64, device=dev) * wl["rec_len"]
        torch.cuda.synchronize()
        return bases, off
    bases = torch.empty(n_bases, dtype=torch.uint8, device=dev)
    kc.gen_bases(seed, 0, n_bases, bases.data_ptr())
    if wl["rec_len"] > 0:
        off = torch.arange(0, n_bases + wl["rec_len"], wl["rec_len"], dtype=torch.int64, device=dev).clamp(max=n_bases)
        off = torch.unique(off)
    else:
        kc.gen_nruns(seed, 0, n_bases, bases.data_ptr())
        off = torch.from_numpy(gen.read_offsets(seed, n_bases).astype(np.int64)).to(dev)
    torch.cuda.synchronize()
    return bases, off


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


def config_of(name, wl, n_bases):
    """The workload, in the same words for both arms."""
    return {"workload": name + ": " + wl["desc"], "k": wl["k"], "canonical": wl["canonical"], "bases_per_gpu": n_bases,
            "record_len": wl["rec_len"], "seed": seed_of(name, 0), "generator": "philox4x32-10 (kmc_gen / gen.py)",
            "l2": "inputs (>= 1 GB) larger than the 126 MB L2; no explicit flush"}


def run_reference(args, wl):
    """--impl reference: the reference's CPU path.  The Rust binary cannot be built in this image (no
    cargo/rustc, crates.io deps), so this times the CPU oracle — its C restatement — with all host cores,
    each step on a bounded prefix of the very bytes the GPU arm counts (rank 0's input)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import orc
    from kmer_count_b200 import gen
    orc.build()
    cores = os.cpu_count() or 1
    sample = min(wl["bases"], args.cpu_sample)
    bases, off = host_input(gen, np, args.workload, wl, 0, sample)
    for _ in range(args.warmup):
        q = len(off) // 4
        cpu_baseline(orc, bases[:int(off[q])], off[:q + 1], wl["k"], wl["canonical"], cores)
    times, n_total = [], 0
    for _ in range(args.steps):
        v, dt, tab = cpu_baseline(orc, bases, off, wl["k"], wl["canonical"], cores)
        times.append(dt)
        n_total = tab.n_total
    ms = 1e3 * sum(times) / len(times)
    value = n_total / (ms / 1e3) / 1e9
    sample_txt = f"first {sample:.3g} bases of rank 0's input per step (the same generator bytes as the GPU arm)"
    print(json.dumps({
        "impl": "reference", "metric": "k-mers counted/sec", "value": value, "unit": "Gk/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64" if wl["k"] <= 32 else "u128", "data": "synthetic",
        "config": config_of(args.workload, wl, wl["bases"]),
        "cpu_baseline": {"value": value, "unit": "Gk/s", "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": "Gk/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


class Runner:
    """One workload on this rank's GPU through DistCounter: device-resident steps, e2e steps, stats."""

    def __init__(self, torch, np, K, gen, dist, world, rank, local, args):
        self.torch, self.np, self.K, self.gen, self.dist = torch, np, K, gen, dist
        self.world, self.rank, self.local, self.args = world, rank, local, args
        self.dev = torch.device("cuda", local)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, name, wl, steps, warmup, sampler=None, e2e=True):
        torch, np, dist = self.torch, self.np, self.dist
        from kmer_count_b200.dist import DistCounter
        stream = torch.cuda.current_stream()
        dc = DistCounter(k=wl["k"], canonical=wl["canonical"], strategy=self.args.strategy, device=self.local,
                         world=self.world, rank=self.rank, dist=dist, torch=torch)
        dc.set_stream(stream.cuda_stream)
        bases, off = device_input(torch, np, self.gen, dc.kc, name, wl, self.rank, wl["bases"], self.dev)
        n, n_recs = bases.numel(), off.numel() - 1
        first_mib = hashlib.sha256(bases[: 1 << 20].cpu().numpy().tobytes()).hexdigest()

        def step_device():
            dc.reset()
            dc.submit_device(bases.data_ptr(), off.data_ptr(), n, n_recs)
            return dc.finish()

        if sampler is not None:
            sampler.start()   # nvidia-smi needs ~100 ms to produce its first row: start it before the warm-up steps
        for _ in range(warmup):
            step_device()
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        kstats, totals, step_wall = [], None, []
        for _ in range(steps):
            t_s = time.perf_counter()
            totals = step_device()
            step_wall.append(round(1e3 * (time.perf_counter() - t_s), 2))
            kstats.append(dc.stats())
        ev1.record(stream)
        self.barrier()
        if sampler is not None and not sampler.rows:
            time.sleep(0.15)  # very short runs: let the sampler deliver at least one row (GPU still warm)
        clocks = sampler.stop() if sampler is not None else None
        ms_local = ev0.elapsed_time(ev1) / steps
        digest = dc.digest()

        skip_e2e = (not e2e) or n > 4_000_000_000   # cfg5: 12.5 GB of pinned host memory per rank is not a reasonable ask
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
                return dc.finish()      # (n_distinct, n_total) read back from the device: the step's result on the host

            step_e2e()
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_res = step_e2e()
            self.barrier()
            e2e_ms_local = 1e3 * (time.perf_counter() - t0) / steps
            # the table the last e2e step left in HBM must be the one the device-resident steps produced (checked outside
            # the timed region: the digest is one more pass over the table and not part of producing the result)
            e2e_res = (*e2e_res, dc.digest())
            del hb, ho
        else:
            e2e_ms_local = float("inf")
            e2e_res = (totals[0], totals[1], digest)
        # whole-job digests (sum over ranks mod 2^64: which rank owns which keys may differ between the two runs)
        dg = torch.tensor(np.array([digest, e2e_res[2]], np.uint64).view(np.int64), device=self.dev)
        t = torch.tensor([ms_local, e2e_ms_local], dtype=torch.float64, device=self.dev)
        cnt = torch.tensor([totals[1], totals[0]], dtype=torch.int64, device=self.dev)
        if dist is not None:
            dist.all_reduce(dg, op=dist.ReduceOp.SUM)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dg = dg.cpu().numpy().view(np.uint64)
        assert int(dg[0]) == int(dg[1]), "e2e and device-resident runs disagree"
        ms, e2e_ms = float(t[0]), float(t[1])
        n_total, n_distinct = int(cnt[0]), int(cnt[1])
        all_walls = [step_wall]
        if dist is not None and os.environ.get("KMC_BENCH_STEPS") == "1":
            all_walls = [None] * self.world
            dist.all_gather_object(all_walls, step_wall)

        # ---- parity: a small shard per rank through the same path, against the oracle on the concatenated input
        parity = None
        if not self.args.no_cpu:
            pn = min(PARITY_BASES, wl["bases"])
            pwl = dict(wl, bases=pn)
            pb, po = device_input(torch, np, self.gen, dc.kc, name, pwl, self.rank + 100, pn, self.dev)
            dc.reset()
            dc.submit_device(pb.data_ptr(), po.data_ptr(), pb.numel(), po.numel() - 1)
            pd, pt = dc.finish()
            got = torch.tensor(np.array([pt, pd, dc.digest()], np.uint64).view(np.int64), device=self.dev)
            if dist is not None:
                dist.all_reduce(got, op=dist.ReduceOp.SUM)
            got = [int(x) for x in got.cpu().numpy().view(np.uint64)]
            if self.rank == 0:
                from oracle import orc  # checker only
                orc.build()
                parts = [host_input(self.gen, np, name, pwl, r + 100, pn) for r in range(self.world)]
                allb = np.concatenate([p[0] for p in parts])
                shift, offs = 0, [np.zeros(1, np.uint64)]
                for p in parts:
                    offs.append(p[1][1:] + np.uint64(shift))
                    shift += len(p[0])
                want = orc.contiguous_mt(allb, np.concatenate(offs), wl["k"], wl["canonical"], threads=os.cpu_count() or 1)
                ok = got == [want.n_total, want.n_distinct, want.digest()]
                parity = {"ok": bool(ok), "bases_per_rank": pn, "ranks": self.world, "n_total": got[0], "n_distinct": got[1],
                          "path": getattr(dc, "path", None) or "single",
                          "check": "(n_total, n_distinct, sum of table digests) == CPU oracle on the concatenated shards"}
                if not ok:
                    parity["want"] = [want.n_total, want.n_distinct, want.digest()]
                    parity["got"] = got
            del pb, po

        res = None
        if self.rank == 0:
            peak, peak_src = hbm_peak()
            wide = wl["k"] > 32
            W = 16 if wide else 8
            dn = n_distinct / max(1, n_total)
            b_alg = 1 + (W + 4) + 4 + (W + 4) * dn  # bytes per k-mer occurrence
            value = n_total / (ms / 1e3) / 1e9
            pipe_gbs = n_total / self.world * b_alg / (ms / 1e3) / 1e9
            # dominant kernel of this rank's last steps (CUDA events inside libkmc, on the launching stream)
            agg = {}
            for st in kstats:
                for kname, v in st.get("kernels", {}).items():
                    a = agg.setdefault(kname, [0, 0.0])
                    a[0] += v["launches"]
                    a[1] += v["ms"]
            per_step_total = sum(v[1] for v in agg.values()) / steps if agg else 0.0
            dom = max(agg.items(), key=lambda kv: kv[1][1]) if agg else (None, [0, 0.0])
            kbytes = dc.algorithmic_bytes(dom[0], n_total // self.world, n_distinct // self.world, n, W) if dom[0] else None
            dom_ms = dom[1][1] / max(1, dom[1][0])
            launches_per_step = dom[1][0] / steps
            k_ach = (kbytes / launches_per_step / (dom_ms / 1e3) / 1e9) if kbytes else None
            roof = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                    "achieved": pipe_gbs, "frac": pipe_gbs / peak, "b_alg_per_kmer": b_alg,
                    "definition": "SURVEY.md 8d: (17 + 12 D/N | 25 + 20 D/N) bytes x k-mers per GPU / step time",
                    "kernel": dom[0], "kernel_achieved": k_ach, "kernel_frac": (k_ach / peak) if k_ach else None,
                    "avg_launch_ms": dom_ms, "launches_per_step": launches_per_step,
                    "share_of_step": (dom[1][1] / steps) / per_step_total if per_step_total else None,
                    "traffic": None}
            prof = os.path.join(REPO, "profiles", "traffic.json")
            if os.path.exists(prof):
                roof["traffic"] = json.load(open(prof)).get(dom[0] or "", None)
            res = {
                "value": value, "ms_per_step": ms, "n_total": n_total, "n_distinct": n_distinct, "digest": int(dg[0]),
                "config": config_of(name, wl, n),
                "run": {"records_per_gpu": n_recs, "first_mib_sha256": first_mib, "strategy": kstats[-1].get("strategy_used"),
                        "variant": kstats[-1].get("fast_variant"),
                        "parallelism": f"{getattr(dc, 'path', None) or 'single'}-partition x{self.world}"},
                "roofline": roof,
                "e2e": None if skip_e2e else {
                    "value": n_total / (e2e_ms / 1e3) / 1e9, "unit": "Gk/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(n + 8 * (n_recs + 1)), "d2h_bytes_per_step": 16,
                    "note": "kmc_submit_host (pinned) + kmc_finish (n_distinct, n_total read back; libkmc also reads ~40 KB "
                            "of histogram and cursors for its plan); wall clock, max over ranks; table digest compared with "
                            "the device-resident run afterwards"},
                "gpu_launches": int(sum(st["kernel_launches"] for st in kstats)),
                "kernels_ms_per_step": {kn: v[1] / steps for kn, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
                "phases_ms": kstats[-1].get("phases_ms"),
                "clocks": clocks,
                "parity": parity,
                "step_wall_ms": all_walls if os.environ.get("KMC_BENCH_STEPS") == "1" else step_wall,
            }
        self.last = (dc, bases, off)
        return res


def run_cfg1(torch, np, K, steps, warmup, no_cpu):
    """BASELINE.json configs[0]: the reference's own job (sample.fasta, L27 + R27, gaps 26..86; main.rs:44-90) on the
    GPU — device FASTA parse + pair extraction + count — against the CPU oracle (orc_gapped_mt, all cores)."""
    path = os.path.join(REPO, "tests", "golden", "sample.fasta")
    text = np.frombuffer(open(path, "rb").read(), np.uint8)
    stream = torch.cuda.current_stream()
    with K.KmerCounter(mode=K.MODE_LR_GAPPED, canonical=False) as kc:
        kc.set_stream(stream.cuda_stream)

        def step():
            kc.reset()
            kc.submit_fasta(text)
            return kc.finish()

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            d, t = step()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / steps
        dig, st = kc.digest(), kc.stats()
    out = {"workload": "cfg1: k-mer-count/sample.fasta (200 records x 400 bases), lr-gapped 27+27, chunk 80..140 (main.rs:48-49,63)",
           "value": t / (ms / 1e3) / 1e9, "unit": "Gk/s", "ms_per_step": ms, "n_total": t, "n_distinct": d,
           "timed": "FASTA text in host memory -> sorted (key,count) table in HBM (parse, H2D and count; wall clock)",
           "strategy": st.get("strategy_used"), "variant": st.get("fast_variant"), "fast_fallbacks": st.get("fast_fallbacks"),
           "phases_ms": st.get("phases_ms")}
    if not no_cpu:
        from oracle import orc  # checker + the CPU arm of this block
        orc.build()
        b, o = orc.parse_fasta(path)
        cores = os.cpu_count() or 1
        orc.gapped_mt(b, o, 27, 27, 80, 140, threads=cores)
        t0 = time.perf_counter()
        for _ in range(steps):
            want = orc.gapped_mt(b, o, 27, 27, 80, 140, threads=cores)
        cpu_ms = 1e3 * (time.perf_counter() - t0) / steps
        out["cpu"] = {"value": want.n_total / (cpu_ms / 1e3) / 1e9, "unit": "Gk/s", "ms_per_step": cpu_ms, "cores": cores,
                      "kind": "port", "what": "oracle orc_gapped_mt, bases in memory -> sorted table in memory"}
        out["parity"] = {"ok": bool((want.n_total, want.n_distinct, want.digest()) == (t, d, dig)),
                         "known_answer": [3550200, 1079497]}
    return out


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
    ap.add_argument("--no-cpu", action="store_true", help="skip everything that runs the CPU oracle (baseline, parity)")
    ap.add_argument("--no-extra", action="store_true", help="skip the k31 and cfg1 blocks")
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
    from kmer_count_b200 import gen

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line (rank 0): whatever libraries print there (NCCL's version banner, ...) is sent
    # to stderr instead, and the line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libkmc has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL's init lines (communicator, nranks, NVLS) go to stderr — stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG", os.environ.get("KMC_NCCL_DEBUG", "INFO"))
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("KMC_KERNEL_TIMING", "1")
    K.build()

    R = Runner(torch, np, K, gen, dist, world, rank, local, args)
    sampler = ClockSampler(local) if rank == 0 else None
    main_res = R.run(args.workload, wl, args.steps, args.warmup, sampler)
    dc, bases, off = R.last
    out = None
    if rank == 0:
        wide = wl["k"] > 32
        out = {"metric": "k-mers counted/sec", "value": main_res["value"], "unit": "Gk/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "u128" if wide else "u64", "data": "synthetic"}
        out.update({k: v for k, v in main_res.items() if k not in ("value", "ms_per_step")})
        if not args.no_cpu:
            from oracle import orc  # cpu_baseline leg only
            orc.build()
            cores = os.cpu_count() or 1
            sample = min(bases.numel(), args.cpu_sample)
            sb, so = host_input(gen, np, args.workload, wl, 0, sample)
            assert np.array_equal(sb[: 1 << 16], bases[: 1 << 16].cpu().numpy())   # host twin == device generator
            v, dt, tab = cpu_baseline(orc, sb, so, wl["k"], wl["canonical"], cores)
            out["cpu_baseline"] = {"value": v, "unit": "Gk/s", "cores": cores, "kind": "port", "seconds": dt,
                                   "sample": f"first {sample:.3g} bases of rank 0's input, one pass, oracle/liborc.so "
                                             f"orc_contiguous_mt (C restatement; the Rust reference cannot be built here)"}
    dc.close()
    del dc, bases, off
    R.last = None
    torch.cuda.empty_cache()
    if not args.no_extra and args.workload == "cfg2":
        k31 = R.run("cfg3", dict(WORKLOADS["cfg3"]), args.steps, args.warmup)
        R.last[0].close()
        R.last = None
        torch.cuda.empty_cache()
        if rank == 0:
            out["k31"] = {k: v for k, v in k31.items() if k not in ("clocks", "step_wall_ms")}
            out["k31"]["unit"] = "Gk/s"
            out["cfg1"] = run_cfg1(torch, np, K, args.steps, args.warmup, args.no_cpu) if world == 1 else None
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
