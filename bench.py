#!/usr/bin/env python3
"""bench.py -- haplotype-pair*bp/s of the windowed pi / Hudson Fst / Tajima's D hot path.

Workload (BASELINE.json configs[1]): h-fst.py-style Hudson Fst, AFR (140) vs EAS (100) panels inside a
466-haplotype synthetic HPRC-shaped panel, 50 kb windows over a chr2-length graph (4 854 windows; every
window also yields pi, S and Tajima's D in the same pass).  One "step" = one pass of the fused path over
the whole batch; at N > 1 every rank owns its own chromosome-length batch (weak scaling) and the result
rows are all-gathered once per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (see the task contract): value = device-resident throughput, e2e = through the
public API with host buffers, roofline = dominant kernel vs its bound, cpu_baseline = the CPU oracle port
timed on this box's host cores.
"""
from __future__ import annotations

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

N_HAP = 466
WINDOW_BP = 50_000
CHR2_BP = 242_696_752                     # CHM13 v2.0 chr2
WINDOWS = -(-CHR2_BP // WINDOW_BP)        # 4 854
METRIC = "haplotype-pair*bp/s (windowed pi / Hudson Fst / Tajima's D)"
UNIT = "hap-pair*bp/s"
LAB_SUBSET, LAB_A, LAB_B, LAB_SEG = 1, 2, 4, 8


def labels_for(pops: np.ndarray) -> np.ndarray:
    lab = np.full(pops.shape[0], LAB_SUBSET | LAB_SEG, dtype=np.uint8)
    lab[pops == 0] |= LAB_A               # AFR
    lab[pops == 2] |= LAB_B               # EAS
    return lab


def units_per_window(n: int, length: int) -> float:
    return n * (n - 1) / 2.0 * length


def load_peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) and profiles/int8_peak.json (measured by
    tools/measure_int8_peak.py on this pool); else the stated fallbacks."""
    out = {"hbm_gbs": 6650.0, "hbm_src": "fallback", "int8_tops": None, "int8_src": None, "bf16_tflops": 1590.0}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mp = json.load(fh)
        out.update(hbm_gbs=float(mp["hbm_gbs"]), hbm_src="measured", bf16_tflops=float(mp["bf16_tflops"]))
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "int8_peak.json")) as fh:
            ip = json.load(fh)
        out.update(int8_tops=float(ip["int8_tops"]), int8_src="measured (profiles/int8_peak.json, torch._int_mm 8192^3)")
    except Exception:
        out.update(int8_tops=2.0 * out["bf16_tflops"],
                   int8_src="2 x measured bf16 (no int8 entry in MEASURED_PEAKS.json; nominal int8:bf16 = 2:1)")
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_rate(x_bits, node_len, labels, n, m_pad, pitch, length, threads, target_s=12.0, max_windows=None):
    """Time the plain-C oracle port (oracle/csrc/oracle_impop.c, pthreads) on a bounded sample sized to
    ~target_s seconds.  Returns (pair*bp/s, windows in the sample, seconds, stats, counts)."""
    from oracle import clib
    total = x_bits.shape[0] if max_windows is None else min(max_windows, x_bits.shape[0])

    def run(S):
        ar = np.arange(S, dtype=np.int64)
        t0 = time.perf_counter()
        st, ct = clib.batch_stats(np.full(S, n), np.full(S, m_pad), np.full(S, pitch), ar * (n * pitch), ar * m_pad,
                                  np.zeros(S, dtype=np.int64), np.full(S, length), x_bits[:S], node_len[:S], labels, threads)
        return time.perf_counter() - t0, st, ct

    s0 = min(total, max(threads * 2, 8))
    t0, st, ct = run(s0)
    S = int(min(total, max(s0, s0 * target_s / max(t0, 1e-3))))
    if S > s0:
        t0, st, ct = run(S)
    else:
        S = s0
    return S * units_per_window(n, length) / t0, S, t0, st, ct


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (oracle port; the reference itself is
    Python scripts + external odgi/impg binaries, neither of which travels to the GPU box) with all host
    threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch  # noqa: F401  (host tensor plumbing for the generator only)
    from impop_b200 import synth
    threads = len(os.sched_getaffinity(0))
    gen = synth.HostGenerator()
    sample = min(WINDOWS, max(threads * 24, 96))
    x_bits, node_len, pops, m, m_pad = synth.make_windows_device(gen, N_HAP, WINDOW_BP, sample, seed=0xB200 + 1, chunk=128)
    xb = x_bits.numpy().view(np.uint32)
    nl = node_len.numpy().view(np.uint32)
    lab = labels_for(pops)
    pitch = m_pad // 32
    from oracle import clib
    ar = np.arange(sample, dtype=np.int64)

    def step():
        clib.batch_stats(np.full(sample, N_HAP), np.full(sample, m_pad), np.full(sample, pitch), ar * (N_HAP * pitch),
                         ar * m_pad, np.zeros(sample, dtype=np.int64), np.full(sample, WINDOW_BP), xb, nl, lab, threads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample * units_per_window(N_HAP, WINDOW_BP) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64 intersections + f64 statistics", "data": "synthetic",
        "config": {"workload": f"h-fst AFR(140) vs EAS(100), {N_HAP} haplotypes, {WINDOW_BP} bp windows, chr2-length graph",
                   "windows_per_step": sample, "nodes_per_window": int(m_pad)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} of {WINDOWS} windows per step, plain-C oracle (byte-LUT intersections) on {threads} pthreads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from impop_b200 import synth
    from impop_b200.distributed import gather_rows
    from impop_b200.engine import ALGO_SIMT, ALGO_TCGEN05, Context, WindowBatch, NCOUNTS, NSTATS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    dev = ctx.torch_device
    algo = ALGO_SIMT if args.algo == "simt" else ALGO_TCGEN05
    W = args.windows

    # ---------------------------------------------------------------- synthetic batch, resident in HBM
    x_bits, node_len, pops, m, m_pad = synth.make_windows_device(ctx, N_HAP, WINDOW_BP, W, seed=0xB200 + 1 + 1000 * rank)
    pitch = m_pad // 32
    lab_host = labels_for(pops)
    labels = torch.from_numpy(lab_host).to(dev)
    batch = WindowBatch.from_uniform(ctx, x_bits, node_len, labels, WINDOW_BP)
    stats = torch.empty((W, NSTATS), dtype=torch.float64, device=dev)
    counts = torch.empty((W, NCOUNTS), dtype=torch.int64, device=dev)
    bounds = np.arange(world + 1, dtype=np.int64) * W

    def step():
        batch.stats(algo, out_stats=stats, out_counts=counts)
        if world > 1:                     # the single result gather of the north star (W x 20 fp64 per rank)
            return gather_rows(stats, bounds)
        return stats

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    ctx.check()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    launches0 = ctx.launches
    ctx.timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms_step = ev0.elapsed_time(ev1) / args.steps
    launches = ctx.launches - launches0
    pairs_ms, pairs_n = ctx.timing_read("pairs")
    per_kernel = {k: ctx.timing_read(k)[0] / max(args.steps, 1) for k in ("prep", "pairs", "sums", "finalize")}
    ctx.timing(False)
    ctx.check()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    units_step = world * W * units_per_window(N_HAP, WINDOW_BP)
    value = units_step / (ms_step * 1e-3)

    # ---------------------------------------------------------------- e2e: host buffers through the public API
    # Every step: pinned host -> device copies of that step's inputs, batch set-up, the fused kernels and the
    # device -> host read of the result rows, all inside the timed region.  The batch is cut into sub-batches
    # that alternate between two streams so copies overlap kernels (public API: WindowBatch + stream arguments).
    hx = torch.empty(x_bits.shape, dtype=torch.int32, pin_memory=True); hx.copy_(x_bits)
    hl = torch.empty(node_len.shape, dtype=torch.int32, pin_memory=True); hl.copy_(node_len)
    hlab = torch.from_numpy(lab_host).pin_memory()
    hs = torch.empty((W, NSTATS), dtype=torch.float64, pin_memory=True)
    hc = torch.empty((W, NCOUNTS), dtype=torch.int64, pin_memory=True)
    dx, dl = torch.empty_like(x_bits), torch.empty_like(node_len)
    ds, dc = torch.empty_like(stats), torch.empty_like(counts)
    nsub = max(1, min(args.sub_batches, W))
    if args.e2e_plain or nsub < 4:
        cuts = [int(v) for v in np.linspace(0, W, nsub + 1)]
    else:
        # the last two sub-batches are smaller: what is left to compute after the final copy lands is the step's tail
        wts = np.ones(nsub); wts[-2], wts[-1] = 0.6, 0.3
        cuts = [0] + [int(v) for v in np.round(np.cumsum(wts) / wts.sum() * W)]
        cuts[-1] = W
    streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, args.e2e_streams))]
    dlabs = [torch.empty_like(labels) for _ in range(nsub)]

    pending = []          # batches of the previous step: closed while this step's copies and kernels run
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(nsub)]

    def e2e_step():
        live = []
        for k in range(nsub):
            lo, hi = cuts[k], cuts[k + 1]
            if hi <= lo:
                continue
            st = streams[k % len(streams)]
            if args.e2e_copy_stream:
                # every host -> device copy goes through ONE stream, in sub-batch order, so the copy engine (the bottleneck
                # of the step) never waits for a kernel; the set-up's small table upload follows its sub-batch's big copy in
                # the same stream; the kernels run on the two compute streams behind an event
                with torch.cuda.stream(copy_stream):
                    dlabs[k].copy_(hlab, non_blocking=True)
                    dl[lo:hi].copy_(hl[lo:hi], non_blocking=True)
                    dx[lo:hi].copy_(hx[lo:hi], non_blocking=True)
                    b = WindowBatch.from_uniform(ctx, dx[lo:hi], dl[lo:hi], dlabs[k], WINDOW_BP, node_len_host=hl[lo:hi], stream=copy_stream)
                    copied[k].record(copy_stream)
                with torch.cuda.stream(st):
                    st.wait_event(copied[k])
                    b.stats(algo, stream=st, out_stats=ds[lo:hi], out_counts=dc[lo:hi])
                    hs[lo:hi].copy_(ds[lo:hi], non_blocking=True)
                    hc[lo:hi].copy_(dc[lo:hi], non_blocking=True)
                live.append(b)
                continue
            with torch.cuda.stream(st):
                # this sub-batch's copies first, its set-up (host-side tables + their small upload) while they run: the copy
                # engine is the bottleneck of the step and must never wait for the host; the table upload lands in the
                # engine's FIFO right behind this sub-batch's own big copy, ahead of the OTHER stream's next one
                dlabs[k].copy_(hlab, non_blocking=True)
                dl[lo:hi].copy_(hl[lo:hi], non_blocking=True)
                dx[lo:hi].copy_(hx[lo:hi], non_blocking=True)
                b = WindowBatch.from_uniform(ctx, dx[lo:hi], dl[lo:hi], dlabs[k], WINDOW_BP, node_len_host=hl[lo:hi], stream=st)
                b.stats(algo, stream=st, out_stats=ds[lo:hi], out_counts=dc[lo:hi])
                hs[lo:hi].copy_(ds[lo:hi], non_blocking=True)
                hc[lo:hi].copy_(dc[lo:hi], non_blocking=True)
            live.append(b)
        if args.e2e_plain:
            for st in streams:
                st.synchronize()
            for b in live:
                b.close()
            return
        while pending:                      # host work hidden behind the copies just enqueued
            pending.pop().close()
        for st in streams:
            st.synchronize()
        pending.extend(live)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    while pending:                          # inside the timed region: every batch of the timed steps is closed
        pending.pop().close()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    ctx.check()
    # the sub-batched run must reproduce the resident run: same NaN pattern, every statistic within 1e-12 of the column's scale
    # (the per-lane sums of the pairs kernel are plain fp64, so the last bits depend on how a batch is dealt to the CTAs)
    ha, hb = hs.numpy(), stats.cpu().numpy()
    scale = np.nanmax(np.abs(hb), axis=0, keepdims=True)
    scale = np.where(np.isfinite(scale), scale, 0.0)
    with np.errstate(invalid="ignore"):
        close = np.abs(ha - hb) <= 1e-12 * np.maximum(scale, np.maximum(np.abs(ha), np.abs(hb)))
    same = bool(np.array_equal(np.isnan(ha), np.isnan(hb)) and bool(np.all(close | np.isnan(hb)))
                and torch.equal(hc, counts.cpu()))
    bitwise = bool(torch.equal(hs.nan_to_num(7.0), stats.cpu().nan_to_num(7.0)))
    h2d = world * (hx.numel() * 4 + hl.numel() * 4 + hlab.numel() * nsub)      # whole job: every rank copies its own batch
    d2h = world * (hs.numel() * 8 + hc.numel() * 8)

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peaks = load_peaks()
    nl_max = int(node_len.max().item())
    planes = 1 if nl_max < 256 else (2 if nl_max < 65536 else (3 if nl_max < (1 << 24) else 4))
    ops_launch = 2.0 * W * (N_HAP * (N_HAP + 1) / 2.0) * m_pad * planes          # int8 ops, input m, P planes
    bytes_launch = W * (N_HAP * (m_pad // 8) + 4 * m_pad + N_HAP + 8 * 14)       # algorithmic HBM bytes
    pairs_avg_s = (pairs_ms / max(pairs_n, 1)) * 1e-3
    tops = ops_launch / pairs_avg_s / 1e12
    roofline = {"bound": "tensor", "algorithmic_ops_per_launch": ops_launch, "algorithmic_bytes_per_launch": bytes_launch,
                "achieved": tops, "peak": peaks["int8_tops"], "unit": "TOP/s (int8)",
                "frac": tops / peaks["int8_tops"], "traffic": None, "kernel": "window_pairs_tc_kernel" if algo == ALGO_TCGEN05 else "window_pairs_simt_kernel",
                "kernel_ms": pairs_avg_s * 1e3, "peak_source": peaks["int8_src"], "byte_planes": planes,
                "hbm": {"achieved_gbs": bytes_launch / pairs_avg_s / 1e9, "peak_gbs": peaks["hbm_gbs"],
                        "frac": bytes_launch / pairs_avg_s / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["hbm_src"]},
                "fp64_pair_epilogues_per_s": W * N_HAP * (N_HAP - 1) / 2.0 / pairs_avg_s,
                # the co-limit DESIGN.md 4.2 derives: 18 fp64 instructions per haplotype pair (2 correctly rounded divisions:
                # 16, + sums) on a pipe of 64 lanes / clk / SM; clock = the SM clock sampled during the timed region
                "fp64": {"instr_per_pair": 18, "lanes_per_clk_per_sm": 64, "sms": int(torch.cuda.get_device_properties(local).multi_processor_count),
                         "frac": None},
                "step_share": {k: v / ms_step for k, v in per_kernel.items()}}
    traffic_file = os.path.join(ROOT, "profiles", "pairs_traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file))["dram_bytes_per_window"] * W   # ncu dram read+write, scaled to this launch
        except Exception:
            pass

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = len(os.sched_getaffinity(0))
        cap = W
        xb = x_bits[:cap].cpu().numpy().view(np.uint32)
        nl = node_len[:cap].cpu().numpy().view(np.uint32)
        rate, S, secs, st_cpu, ct_cpu = cpu_oracle_rate(xb, nl, lab_host, N_HAP, m_pad, pitch, WINDOW_BP, threads)
        got_s, got_c = stats[:S].cpu().numpy(), counts[:S].cpu().numpy()
        from oracle.compare import rows_close
        ok_counts = bool((got_c == ct_cpu).all())
        ok_stats, why = rows_close(got_s, st_cpu)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{S} of {W} windows of this workload in {secs:.2f} s, plain-C oracle (byte-LUT intersections) on {threads} pthreads",
               "gpu_matches_oracle_on_sample": {"counts_exact": ok_counts, "stats_within_1e-12": ok_stats, "detail": why}}

    if rank == 0:
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        f64 = roofline["fp64"]
        f64["frac"] = roofline["fp64_pair_epilogues_per_s"] * f64["instr_per_pair"] / (f64["lanes_per_clk_per_sm"] * f64["sms"] * mhz * 1e6)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 x u8 -> s32 intersections (tcgen05 kind::i8) + f64 statistics" if algo == ALGO_TCGEN05 else "u8 dp4a -> u32 + f64 statistics",
            "data": "synthetic",
            "config": {"workload": f"h-fst AFR(140) vs EAS(100) Hudson Fst + pi + Tajima's D, {N_HAP} haplotypes, {WINDOW_BP} bp windows, chr2-length graph",
                       "windows_per_gpu": W, "nodes_per_window": int(m_pad), "haplotypes": N_HAP,
                       "parallelism": f"windows sharded over {world} GPU(s), one all-gather of result rows per step",
                       "l2": f"inputs larger than L2 ({(x_bits.numel() * 4 + node_len.numel() * 4) / 1e6:.0f} MB per GPU read every step)",
                       "algo": args.algo},
            "e2e": {"value": units_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "sub_batches": nsub, "matches_resident_run": same, "bitwise_equal_to_resident_run": bitwise},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    batch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="tc", choices=["tc", "simt"])
    ap.add_argument("--windows", type=int, default=WINDOWS, help="windows per GPU (default: chr2 / 50 kb = 4854)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sub-batches", type=int, default=8, help="e2e leg: sub-batches alternating between two streams")
    ap.add_argument("--e2e-streams", type=int, default=2, help="e2e leg: streams the sub-batches rotate over")
    ap.add_argument("--e2e-copy-stream", type=int, default=0, help="e2e leg: 1 = all host->device copies on one stream, kernels behind events")
    ap.add_argument("--e2e-plain", action="store_true", help="e2e leg: equal sub-batches, batches closed at the end of their own step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
