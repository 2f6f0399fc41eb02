#!/usr/bin/env python
"""bench.py -- throughput of the per-read repeat-count hot path on synthetic HiFi reads.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path, host cores)

A "step" is one pass of the hot path over one batch (loci_per_step loci x 30 reads, BASELINE config 2
distributions: "synthetic genome-wide catalog 1M loci x 30x HiFi reads, loci sharded across GPUs").  The
default K x loci_per_step is the full 1M-locus catalog on one GPU.  Multi-GPU: loci are sharded by catalog
partition, one rank per GPU, NO collective on the data path (weak scaling: every rank gets its own
loci_per_step per step); torch.distributed is used for the barriers and the max-over-ranks timing only.

Printed JSON (one line, rank 0): metric / value = reads x loci per second with the batch resident in HBM;
e2e = the same through the host-buffer C-ABI call (pinned host arenas, H2D + kernels + D2H inside the
timed region); roofline = the DP kernel against the measured integer issue rate; cpu_baseline = the CPU
oracle (a port of the reference's algorithm; the real parasail/strkit_rust_ext are not installable here)
timed on this box's host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "read x locus repeat counts per second"
UNIT = "reads*loci/s"
READS_PER_LOCUS = 30


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loci-per-step", type=int, default=32768)
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic batches cycled through (> L2)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "general"])
    ap.add_argument("--cpu-sample-loci", type=int, default=4096,
                    help="loci of batch 0 the CPU port runs (6-7 s on 16 cores, ~25 s on 8 slow ones)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = the same number of blocks as --steps")
    return ap.parse_args()


def workload_name(args):
    return (f"cfg2: synthetic genome-wide catalog, HiFi-like reads, motif 2-6 bp x 10-60 copies, "
            f"{READS_PER_LOCUS} reads/locus, {args.loci_per_step} loci per step per GPU "
            f"({args.steps * args.loci_per_step} loci per GPU in the timed region)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def ncu_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed
    `ncu --set full` summary (profiles/, a separate run under the profiler; None if the file is not there)."""
    import re

    path = os.path.join(ROOT, "profiles", "r1_full_packed_v9.txt")
    if not os.path.exists(path):
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, kernel = 0.0, None
    for line in open(path):
        if line.startswith("# kernel:"):
            if kernel is not None:
                break  # first kernel of the capture only
            kernel = line.split(":", 1)[1].strip().split("(")[0]
        m = re.match(r"dram__bytes_(read|write)\.sum\s+(\w+)\s+([0-9.]+)", line)
        if m and kernel is not None:
            total += float(m.group(3)) * unit.get(m.group(2), 1.0)
    return (total or None), f"{kernel}, one launch, profiles/r1_full_packed_v9.txt"


def cpu_baseline(batch, n_loci_sample, steps=1):
    """The oracle (port of the reference's CPU algorithm) on the first n_loci_sample loci, all host cores."""
    from tests import oracle_lib

    orc = oracle_lib.load()
    cores = os.cpu_count() or 1
    sub = batch.slice_loci(0, min(n_loci_sample, batch.n_loci))
    best, cells = None, 0.0
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        out, cells = orc.count_loci(sub.arena, sub.seq_off, sub.lens, sub.est_cn, sub.read_begin, sub.motif_off,
                                    sub.motif_len, n_threads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": sub.n_reads / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sub.n_loci} loci ({sub.n_reads} reads) of batch 0, {best:.2f} s, "
                      f"{cells / best / 1e9:.2f} GCUPS reference-equivalent",
            "gcups": cells / best / 1e9}, out, sub


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; parasail and
    strkit_rust_ext cannot be installed here) on the host cores, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch  # noqa: F401  (data synthesis only)

    from strkit_b200 import synth

    t_all = time.perf_counter()
    from tests import oracle_lib

    orc = oracle_lib.load()
    cores = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # Each step is a bounded sample of the workload, sized so that the W + K steps end within ~100 s on this host:
    # a 256-locus probe gives the port's rate here, the sample is at most --cpu-sample-loci loci and at least 32.
    probe = synth.generate(synth.CONFIGS[2], 256, seed=20261018 + 1999, device="cpu").to_host()
    t0 = time.perf_counter()
    orc.count_loci(probe.arena, probe.seq_off, probe.lens, probe.est_cn, probe.read_begin, probe.motif_off,
                   probe.motif_len, n_threads=cores)
    loci_per_s = probe.n_loci / (time.perf_counter() - t0)
    n_sample = int(min(args.cpu_sample_loci, max(32, 100.0 * loci_per_s / (steps + warm))))
    batch = synth.generate(synth.CONFIGS[2], n_sample, seed=20261018 + 2000, device="cpu").to_host()
    times, cells = [], 0.0
    for it in range(warm + steps):
        t0 = time.perf_counter()
        _, cells = orc.count_loci(batch.arena, batch.seq_off, batch.lens, batch.est_cn, batch.read_begin,
                                  batch.motif_off, batch.motif_len, n_threads=cores)
        if it >= warm:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    value = batch.n_reads / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": f"{n_sample} loci x {READS_PER_LOCUS} reads per step"},
            "gcups_reference_equivalent": cells / dt / 1e9,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_sample} loci ({batch.n_reads} reads) per step, {steps} steps"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch

    import strkit_b200
    from strkit_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner off it (the image exports
        # NCCL_DEBUG=VERSION, and NCCL prints the banner to stdout at VERSION and at WARN)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # Multi-rank runs: keep each rank's host threads and pinned arenas on the cores / memory next to its GPU
    # (the e2e path streams ~300 MB per step per GPU from host memory).  Undone before the CPU baseline.
    full_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_note = None
    if world > 1 and full_affinity is not None:
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1} & full_affinity
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa_note = f"rank pinned to the {len(cpus)} host cores local to its GPU"
        except Exception as exc:  # affinity is an optimisation, never a requirement
            numa_note = f"no GPU-local affinity ({type(exc).__name__})"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    eng = strkit_b200.Engine(device=local_rank)
    params = strkit_b200.RepeatCountParams("repalign", 50, 3, 1)  # read-path defaults (params.py:45,157-163)
    kernel = strkit_b200.KERNEL_GENERAL if args.kernel == "general" else strkit_b200.KERNEL_AUTO

    # ---- synthetic catalog partition of this rank: `pool` distinct batches, generated on the GPU
    host_batches, dev_batches = [], []
    for p in range(args.pool):
        sbatch = synth.generate(synth.CONFIGS[2], args.loci_per_step, seed=20261018 + 2000 + 100 * rank + p,
                                device=str(dev), chunk_loci=4096)
        hb = sbatch.to_host(pin=True)
        del sbatch
        host_batches.append(hb)
        dev_batches.append(eng.upload(hb))
    torch.cuda.empty_cache()
    reads_per_step = host_batches[0].n_reads
    pool_bytes = sum(b.nbytes() for b in host_batches)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- integer issue-rate peak (roofline denominator), measured on this GPU now
    peak = eng.measure_int_peak()

    # ---- warm-up
    for w in range(args.warmup):
        eng.run(dev_batches[w % args.pool], params, kernel, stream)

    # ---- timed region: K steps, inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = {k: 0.0 for k in ("executed_cells", "reference_cells", "kernel_launches", "dp_ms", "replay_ms",
                            "widening_passes", "reads_packed_kernel", "reads_general_kernel")}
    ev0.record()
    for k in range(args.steps):
        eng.run(dev_batches[k % args.pool], params, kernel, stream)
        st = eng.stats()
        for key in agg:
            agg[key] += st[key]
    ev1.record()
    barrier()
    elapsed_ms = reduce_max(ev0.elapsed_time(ev1))
    clocks = sampler.stop()
    total_reads = reduce_sum(float(reads_per_step) * args.steps)
    value = total_reads / (elapsed_ms * 1e-3)
    exec_cells = reduce_sum(agg["executed_cells"])
    ref_cells = reduce_sum(agg["reference_cells"])

    # ---- e2e: host buffers through the C-ABI call (H2D + kernels + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or args.steps
        outs = [np.empty((b.n_reads, 4), dtype=np.int32) for b in host_batches]
        for o in outs:
            strkit_b200._native.check(strkit_b200._native.lib.strk_host_register(o.ctypes.data, o.nbytes))
        # (a) one blocking C-ABI call per block; (b) the streamed form: the same calls from two host threads /
        # two native contexts, the H2D copy of block i+1 overlapping the kernels of block i
        eng.count_reads(host_batches[0], params, kernel, out=outs[0])  # warm the recycled device buffers
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            eng.count_reads(host_batches[k % args.pool], params, kernel, out=outs[k % args.pool])
        torch.cuda.synchronize()
        dt_call = time.perf_counter() - t0
        barrier()
        dt_call = reduce_max(dt_call)
        for _ in eng.count_reads_stream(host_batches[:2], params, kernel, outs=outs[:2]):  # warm both contexts
            pass
        barrier()
        t0 = time.perf_counter()
        n_done = 0
        for _ in eng.count_reads_stream((host_batches[k % args.pool] for k in range(e2e_steps)), params, kernel,
                                        outs=(outs[k % args.pool] for k in range(e2e_steps))):
            n_done += 1
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        dt = reduce_max(dt)
        assert n_done == e2e_steps
        e2e_parity = bool(np.array_equal(outs[(e2e_steps - 1) % args.pool],
                                         eng.download(dev_batches[(e2e_steps - 1) % args.pool])))
        e2e = {"value": reduce_sum(float(reads_per_step) * e2e_steps) / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(host_batches[0].nbytes()),
               "d2h_bytes_per_step": int(outs[0].nbytes), "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "how": "Engine.count_reads_stream: host arrays (pinned) -> strk_batch_fill / strk_batch_run / "
                      "strk_batch_download per block, two host threads, copies overlapped with the kernels",
               "blocking_call_value": reduce_sum(float(reads_per_step) * e2e_steps) / dt_call,
               "equals_resident_results": e2e_parity}

    # ---- the reference-genome path of the same loci (get_ref_repeat_count, once per locus; repeats.py:73-192):
    # the first read of every locus of batch 0 stands in for the reference window
    ref_path = None
    if rank == 0:
        from strkit_b200.batcher import ReadBatch

        hb = host_batches[0]
        first = hb.read_begin[:-1]
        lens = hb.lens[first].copy()
        tot = lens.sum(axis=1).astype(np.int64)
        r_off = np.concatenate([[0], np.cumsum(tot)[:-1]]).astype(np.int64)
        src = np.repeat(hb.seq_off[first].astype(np.int64) - r_off, tot) + np.arange(int(tot.sum()))
        ml = hb.motif_len.astype(np.int64)
        m_off = int(tot.sum()) + np.concatenate([[0], np.cumsum(ml)[:-1]]).astype(np.int64)
        msrc = np.repeat(hb.motif_off.astype(np.int64) - m_off, ml) + int(tot.sum()) + np.arange(int(ml.sum()))
        ref = ReadBatch(arena=np.concatenate([hb.arena[src], hb.arena[msrc]]), seq_off=r_off.astype(np.uint64), lens=lens,
                        est_cn=hb.est_cn[first].copy(), read_begin=np.arange(hb.n_loci + 1, dtype=np.int64),
                        motif_off=m_off.astype(np.uint64), motif_len=hb.motif_len)
        rc = np.tile(np.array([250, 3, 1], dtype=np.int32), (hb.n_loci, 1))  # repeat_count_params.py:25-27
        start, ref_size = ref.est_cn.copy(), ref.lens[:, 1].copy()
        for a in (ref.arena, ref.seq_off, ref.lens, ref.motif_off, start, ref_size, rc):  # pinned, like the reads
            strkit_b200._native.check(strkit_b200._native.lib.strk_host_register(a.ctypes.data, a.nbytes))
        eng.ref_counts(ref, start, ref_size, rc, 5)  # grows the recycled device buffers
        t0 = time.perf_counter()
        for _ in range(3):
            eng.ref_counts(ref, start, ref_size, rc, 5)
        dt_ref = (time.perf_counter() - t0) / 3
        ref_path = {"loci_per_s": hb.n_loci / dt_ref, "ms_per_block": dt_ref * 1e3, "loci_per_block": hb.n_loci,
                    "how": "Engine.ref_counts, host arrays in / 8 ints per locus out, one GPU",
                    "share_of_read_path_time": dt_ref * 1e3 / (elapsed_ms / max(1, args.steps))}

    # ---- parity spot-check + CPU baseline on a bounded sample (rank 0)
    cpu = None
    parity = None
    if full_affinity is not None:
        os.sched_setaffinity(0, full_affinity)
    if rank == 0 and not args.no_cpu_baseline:
        # the timed CPU baseline belongs to the N = 1 line; multi-GPU runs keep a small bit-exactness sample only
        cpu, want, sub = cpu_baseline(host_batches[0], args.cpu_sample_loci if world == 1 else min(256, args.cpu_sample_loci))
        if world > 1:
            cpu = None
        got = eng.download(dev_batches[0])[:sub.n_reads] if args.warmup + args.steps > 0 else None
        if args.pool and (args.steps + args.warmup) > 0:
            # batch 0 was last run in the timed loop or warm-up; its device results are still resident
            parity = bool(np.array_equal(got, want))

    if rank == 0:
        dp_s = agg["dp_ms"] * 1e-3
        # dominant kernel = the DP kernel; algorithmic work = 4 int32 ops per executed cell
        achieved = agg["executed_cells"] * 4.0 / dp_s / 1e12 if dp_s > 0 else 0.0
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
        arena_gbs = (host_batches[0].nbytes() * args.steps) / dp_s / 1e9 if dp_s > 0 else 0.0
        traffic, traffic_src = ncu_dram_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32" if args.kernel == "general" else "u16x2/int32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "reads_per_locus": READS_PER_LOCUS,
                       "loci_per_step_per_gpu": args.loci_per_step, "search": "max_iters 50, range 3, step 1",
                       "alignment": "parasail sg (all ends free), match 2 / mismatch -7 / indel 5",
                       "l2": f"inputs larger than L2: pool of {args.pool} distinct batches, "
                             f"{pool_bytes / 1e6:.0f} MB per GPU, cycled",
                       "parallelism": f"catalog partition x{world}, no collective", "host_affinity": numa_note},
            "gcups_executed": exec_cells / (elapsed_ms * 1e-3) / 1e9,
            "gcups_reference_equivalent": ref_cells / (elapsed_ms * 1e-3) / 1e9,
            "roofline": {"bound": "int_alu", "achieved": achieved, "peak": peak["dual_pipe"], "unit": "Tiop/s",
                         "frac": achieved / peak["dual_pipe"] if peak["dual_pipe"] else None, "traffic": traffic,
                         "traffic_note": f"DRAM bytes ({traffic_src}): mostly write-back of the capture scratch; the "
                                         "algorithmic bytes of that launch are ~35 MB of read arena"
                                         if traffic else None,
                         "kernel": "dp_general_kernel" if agg["reads_packed_kernel"] == 0 else "dp_packed_kernel",
                         "ops_per_cell": 4, "kernel_ms_per_step": agg["dp_ms"] / max(1, args.steps),
                         "kernel_share_of_step": agg["dp_ms"] / elapsed_ms if elapsed_ms else None,
                         "peak_how": "measured now on this GPU: lane-level 32-bit integer instructions/s, "
                                     "VIADDMNMX + IMAD chains on both issue pipes (ALU-pipe only: "
                                     f"{peak['alu_pipe']:.2f}, FMA-pipe only: {peak['fma_pipe']:.2f})",
                         "hbm": {"achieved": arena_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "note": "arena streaming only; the path is INT-ALU bound, not HBM bound"}},
            "cpu_baseline": cpu, "e2e": e2e, "ref_path": ref_path, "gpu_launches": int(agg["kernel_launches"]),
            "replay_ms_per_step": agg["replay_ms"] / max(1, args.steps),
            "widening_passes": int(agg["widening_passes"]),
            "reads_packed_kernel": int(agg["reads_packed_kernel"]),
            "reads_general_kernel": int(agg["reads_general_kernel"]),
            # rank 0's share, like the two counters above (identical reads of a locus run the DP once)
            "reads_sharing_an_earlier_reads_table": max(0, int(reads_per_step * args.steps - agg["reads_packed_kernel"]
                                                                - agg["reads_general_kernel"])),
            "parity_sample_bit_exact": parity, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
