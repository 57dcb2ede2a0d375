#!/usr/bin/env python3
"""bench.py -- haplotype-pair*bp/s of the windowed pi / Hudson Fst / Tajima's D hot path.

Default workload (BASELINE.json configs[1]): h-fst.py-style Hudson Fst, AFR (140) vs EAS (100) panels inside a
466-haplotype synthetic HPRC-shaped panel, 50 kb windows over a chr2-length graph (4 854 windows; every window also
yields pi, S and Tajima's D in the same pass).  One "step" = one pass of the fused path over the whole batch; at N > 1
every rank owns its own chromosome-length batch (weak scaling) and the result rows are all-gathered once per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config {1,2,3,4,5}] [--strong]

  --config 3   tj_d.py genome-wide: 20 kb windows, 466 haplotypes, S from segregating nodes; one GPU's share of the
               155 864 windows of a CHM13-length genome per rank (weak), or the whole genome split over the ranks (--strong)
  --config 5   scale-up: 16 windows of 10 000 haplotypes x 200 kb, the tile grid of every window split over the ranks
               (X broadcast once with NCCL, partial sums all-gathered and added in rank order): strong scaling
  --config 4   af.py per-site allele counts, 10^7 sites x 466 haplotypes x 5 panels (sites sharded over the ranks)
  --config 1   pica2.py on one 90-haplotype similarity table: latency of the drop-in CLI (N = 1 only)

Ingest: the windows are generated as full presence matrices (every node a column); `impop_compact_scan / _fill`
(host, once per window, timed and reported as `ingest`) writes the affine form of include/impop_b200.h: the columns every
haplotype carries go into a window constant, identical columns are merged, the two complementary columns of a bi-allelic
bubble become one (I = acc + C - R_i - R_j), empty columns are dropped, the rest is ordered by weight and weights >= 255 are
spread over column copies; every window keeps its own row pitch.  Every result is unchanged (the line carries the
comparison with the oracle run on the ORIGINAL columns); the roofline counts algorithmic operations from the ORIGINAL
node count.  The CPU oracle port is given the plain compacted form (constant columns merged only).

Prints ONE JSON line (see the task contract): value = device-resident throughput, e2e = through the public API with
host buffers, roofline = dominant kernel vs its bound, cpu_baseline = the CPU oracle port timed on this box's host
cores, other_configs (N = 1, default config only) = short measurements of the other BASELINE configs.
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
CHR2_BP = 242_696_752                     # CHM13 v2.0 chr2
GENOME_BP = 3_117_275_501                 # CHM13 v2.0 total
METRIC = "haplotype-pair*bp/s (windowed pi / Hudson Fst / Tajima's D)"
UNIT = "hap-pair*bp/s"
LAB_SUBSET, LAB_A, LAB_B, LAB_SEG = 1, 2, 4, 8
TILE_M, TILE_N, KCHUNK = 128, 256, 128    # tile geometry of the pairs kernel (impop_b200/csrc/common.cuh)

CONFIGS = {
    2: {"n": N_HAP, "L": 50_000, "W": -(-CHR2_BP // 50_000), "labels": "afr_eas", "seed": 0xB200 + 1,
        "workload": f"h-fst AFR(140) vs EAS(100) Hudson Fst + pi + Tajima's D, {N_HAP} haplotypes, 50000 bp windows, chr2-length graph"},
    3: {"n": N_HAP, "L": 20_000, "W": -(-GENOME_BP // 20_000) // 8 + 1, "W_total": -(-GENOME_BP // 20_000), "labels": "all", "seed": 0xB200 + 2,
        "workload": f"tj_d Tajima's D genome-wide (S = segregating nodes) + pi, {N_HAP} haplotypes, 20000 bp windows, CHM13-length genome"},
    5: {"n": 10_000, "L": 200_000, "W": 16, "labels": "halves", "seed": 0xB200 + 4,
        "workload": "scale-up pi + Hudson Fst, 10000 haplotypes (two panels of 5000), 200000 bp windows, tile grid split over the GPUs"},
}


def units_per_window(n: int, length: int) -> float:
    return n * (n - 1) / 2.0 * length


def labels_for(kind: str, pops: np.ndarray) -> np.ndarray:
    lab = np.full(pops.shape[0], LAB_SUBSET | LAB_SEG, dtype=np.uint8)
    if kind == "afr_eas":
        lab[pops == 0] |= LAB_A               # AFR
        lab[pops == 2] |= LAB_B               # EAS
    elif kind == "halves":
        lab[: pops.shape[0] // 2] |= LAB_A
        lab[pops.shape[0] // 2:] |= LAB_B
    return lab


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def load_peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) and profiles/int8_peak.json (measured by
    tools/measure_int8_peak.py on this pool); else the stated fallbacks."""
    out = {"hbm_gbs": 6650.0, "hbm_src": "fallback", "int8_tops": None, "int8_src": None, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mp = json.load(fh)
        out.update(hbm_gbs=float(mp["hbm_gbs"]), hbm_src="measured", bf16_tflops=float(mp["bf16_tflops"]),
                   sm_max_mhz=float(mp.get("sm_max_mhz", 1965.0)))
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


# ------------------------------------------------------------------------------------------------------------------
# Workload: generated windows -> ingest (column compaction) -> arrays the batch is built from
# ------------------------------------------------------------------------------------------------------------------
def make_workload(gen, cfg: dict, windows: int, seed: int, threads: int, keep_original: int = 0, plain: int = 0, affine: bool = True,
                  ragged: bool = True):
    """Generate `windows` windows of config `cfg` with `gen` (a device Context or synth.HostGenerator), then run the
    ingest-time column compaction on the host: the affine form (bubbles merged, constants in C: impop_compact_fill with
    IMPOP_COMPACT_PAIRS | _REPLICATE) is what the GPU arm takes.  Returns a dict of HOST arrays + meta; `keep_original`
    windows of the uncompacted input are kept for the oracle cross-check, and the first `plain` windows are also
    compacted in the plain form (constant columns merged, nothing else) -- the input of the CPU oracle port, which
    walks node columns like the similarity tools walk path steps."""
    from impop_b200 import ingest, synth
    n, L = cfg["n"], cfg["L"]
    if cfg["labels"] == "halves":
        pops = np.repeat([0, 1], n // 2)
        x, nl, pops, m, m_pad = synth.make_windows_device(gen, n, L, windows, seed=seed, pops=pops, chunk=4)
    else:
        x, nl, pops, m, m_pad = synth.make_windows_device(gen, n, L, windows, seed=seed)
    xh = x.cpu().numpy().view(np.uint32)
    lh = nl.cpu().numpy().view(np.uint32)
    del x, nl
    W = int(xh.shape[0])
    ar = np.arange(W, dtype=np.int64)
    t0 = time.perf_counter()
    if affine:
        # every window keeps its own row pitch (a multiple of 128 columns): fewer bytes to upload than one pitch for all
        c = ingest.compact_batch(np.full(W, n), np.full(W, m_pad), np.full(W, m_pad // 32), ar * (n * (m_pad // 32)), ar * m_pad, xh, lh,
                                 threads, uniform_pitch=not ragged)
        xc, lc, m_out, site_runs, pitch_w = c.x_out, c.len_out, c.m_out, c.site_runs, c.pitch_out
        row_adj, win_const, col_mult = c.row_adj, c.win_const, c.col_mult
    else:
        cu = ingest.compact_uniform(xh, lh, threads=threads, pairs=False)
        xc, lc, m_out, site_runs = cu.x, cu.node_len, cu.m, cu.site_runs
        pitch_w = np.full(W, xc.shape[2] if W else 4, dtype=np.int32)
        row_adj = win_const = col_mult = None
    t1 = time.perf_counter()
    x_off = np.concatenate([[0], np.cumsum(n * pitch_w.astype(np.int64))]).astype(np.int64)
    len_off = np.concatenate([[0], np.cumsum(32 * pitch_w.astype(np.int64))]).astype(np.int64)
    # transfer form of the presence bits: tight rows (ceil(m / 32) words each); impop_repitch_rows pads them to the
    # 16-byte rows the kernels read once they are on the device
    tp_w = np.maximum(1, (m_out.astype(np.int64) + 31) // 32).astype(np.int32)
    xt_off = np.concatenate([[0], np.cumsum((n * tp_w.astype(np.int64) + 31) // 32 * 32)]).astype(np.int64)     # windows start on 128 bytes
    x_tight = None
    if affine:
        xflat = xc.reshape(-1)
        x_tight = np.zeros(int(xt_off[-1]), dtype=np.uint32)
        for pw in np.unique(pitch_w):                      # windows of one pitch at a time (a handful of pitches)
            for tw in np.unique(tp_w[pitch_w == pw]):
                sel = np.flatnonzero((pitch_w == pw) & (tp_w == tw))
                src = (x_off[sel][:, None] + np.arange(n * int(pw))[None, :]).reshape(len(sel), n, int(pw))[:, :, :int(tw)]
                dst = (xt_off[sel][:, None] + np.arange(n * int(tw))[None, :])
                x_tight[dst.reshape(-1)] = xflat[src.reshape(-1)]
    pl = ingest.compact_uniform(xh[:plain], lh[:plain], threads=threads, pairs=False) if (plain and affine) else None
    lflat = lc.reshape(-1).astype(np.int64)
    hv = (lflat // 255 + 254) // 255                                               # heavy-table entries per column
    heavy = np.add.reduceat(hv, len_off[:-1]) if W else np.zeros(0, np.int64)
    k_exec = ((m_out.astype(np.int64) + KCHUNK - 1) // KCHUNK + (heavy + KCHUNK - 1) // KCHUNK) * KCHUNK
    nl_max = int(lh.max()) if lh.size else 0
    planes = 1 if nl_max < 256 else (2 if nl_max < 65536 else (3 if nl_max < (1 << 24) else 4))
    aff_bytes = int(row_adj.nbytes + col_mult.nbytes + win_const.nbytes) if affine else 0
    x_bytes = int(x_tight.nbytes) if affine else int(xc.nbytes)
    return {"x": xc, "len": lc, "m_out": m_out, "site_runs": site_runs, "pops": pops, "labels": labels_for(cfg["labels"], pops), "n": n, "L": L,
            "row_adj": row_adj, "win_const": win_const, "col_mult": col_mult, "pitch_w": pitch_w, "x_off": x_off, "len_off": len_off,
            "x_tight": x_tight, "tp_w": tp_w, "xt_off": xt_off, "heavy": heavy.astype(np.int32),
            "m_in": int(m), "m_pad_in": int(m_pad), "pitch": int(pitch_w.max()) if W else 4, "m_pad": int(pitch_w.max()) * 32 if W else 128,
            "planes": planes, "k_exec": k_exec, "ingest_s": t1 - t0, "ingest_threads": threads,
            "plain_x": pl.x if pl is not None else None, "plain_len": pl.node_len if pl is not None else None,
            "plain_m_out": int(pl.m.max()) if (pl is not None and len(pl.m)) else 0,
            "orig_x": xh[:keep_original].copy() if keep_original else None,
            "orig_len": lh[:keep_original].copy() if keep_original else None,
            "bytes_in": int(xh.nbytes + lh.nbytes), "bytes_out": x_bytes + int(lc.nbytes) + aff_bytes}


class DeviceWindows:
    """The ingested windows of a workload as flat device (or pinned host) arrays + what a WindowBatch over windows [lo, hi)
    needs.  The pinned host copy holds the presence bits in their transfer form (tight rows, `xt`); on the device they
    are padded to the 16-byte rows the kernels read (`x`, impop_repitch_rows)."""

    def __init__(self, torch, wl, where, tight=True):
        pinned = where == "pinned"
        mk = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()) if pinned else \
             (lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(where))
        self.x = None if (pinned and tight) else mk(wl["x"].reshape(-1).view(np.int32))
        self.xt = mk(wl["x_tight"].view(np.int32)) if (pinned and tight) else None
        self._tabs = {}
        self.len = mk(wl["len"].reshape(-1).view(np.int32))
        self.row_adj = mk(wl["row_adj"].reshape(-1))
        self.col_mult = mk(wl["col_mult"].reshape(-1))

    @property
    def nbytes(self):
        return sum(int(t.numel()) * t.element_size() for t in (self.x, self.xt, self.len, self.row_adj, self.col_mult) if t is not None)

    def empty_like(self, torch, dev, xt_words=0):
        out = object.__new__(DeviceWindows)
        for k in ("x", "len", "row_adj", "col_mult"):
            setattr(out, k, torch.empty(getattr(self, k).shape, dtype=getattr(self, k).dtype, device=dev))
        out.xt = torch.empty(int(xt_words), dtype=torch.int32, device=dev) if xt_words else None      # landing area of the tight rows
        out._tabs = {}
        return out

    def copy_from(self, ctx, src, wl, lo, hi):
        """Enqueue the copies of windows [lo, hi) from `src` (pinned host) on the current stream: small arrays first, then the
        presence rows (tight: padded on the device afterwards, see repitch)."""
        n = wl["n"]
        x0, l0, l1 = int(wl["x_off"][lo]), int(wl["len_off"][lo]), int(wl["len_off"][hi])
        t0, t1 = int(wl["xt_off"][lo]), int(wl["xt_off"][hi])
        self.len[l0:l1].copy_(src.len[l0:l1], non_blocking=True)
        self.row_adj[lo * n:hi * n].copy_(src.row_adj[lo * n:hi * n], non_blocking=True)
        self.col_mult[l0:l1].copy_(src.col_mult[l0:l1], non_blocking=True)
        x1 = int(wl["x_off"][hi])
        if src.xt is None:                                  # rows transferred as the kernels read them (16-byte multiples)
            self.x[x0:x1].copy_(src.x[x0:x1], non_blocking=True)
            return
        self.xt[t0:t1].copy_(src.xt[t0:t1], non_blocking=True)

    def repitch(self, ctx, wl, lo, hi):
        """Pad the tight rows of windows [lo, hi) (already copied) to 16-byte rows.  Called AFTER the batch is created: a kernel
        between the big copy and the batch's small table upload would let that upload queue up behind the other stream's
        next big copy on the copy engine."""
        if self.xt is None:
            return
        n = wl["n"]
        x0, x1, t0, t1 = int(wl["x_off"][lo]), int(wl["x_off"][hi]), int(wl["xt_off"][lo]), int(wl["xt_off"][hi])
        tabs = self._tabs.get((lo, hi))
        if tabs is None:                                    # per-window tables of this window range: built once
            tabs = self._tabs[(lo, hi)] = (np.full(hi - lo, n, dtype=np.int32), np.ascontiguousarray(wl["tp_w"][lo:hi], dtype=np.int32),
                                           np.ascontiguousarray(wl["pitch_w"][lo:hi], dtype=np.int32),
                                           np.ascontiguousarray(wl["xt_off"][lo:hi] - t0), np.ascontiguousarray(wl["x_off"][lo:hi] - x0))
        ctx.repitch_rows(self.xt[t0:t1], self.x[x0:x1], *tabs)

    def batch(self, ctx, WindowBatch, wl, labels, lo, hi, stream=None, host=None, with_runs=True):
        n, L = wl["n"], wl["L"]
        Wn = hi - lo
        x0, x1, l0, l1 = int(wl["x_off"][lo]), int(wl["x_off"][hi]), int(wl["len_off"][lo]), int(wl["len_off"][hi])
        return WindowBatch(ctx, np.full(Wn, n), wl["m_out"][lo:hi], wl["pitch_w"][lo:hi], wl["x_off"][lo:hi] - x0, wl["len_off"][lo:hi] - l0,
                           np.zeros(Wn, dtype=np.int64), np.full(Wn, L), self.x[x0:x1], self.len[l0:l1], labels,
                           node_len_host=None if host is None else host.len[l0:l1], stream=stream,
                           site_runs=wl["site_runs"][lo:hi] if with_runs else None, row_adj=self.row_adj[lo * n:hi * n],
                           win_const=wl["win_const"][lo:hi], col_mult=self.col_mult[l0:l1],
                           heavy_entries=None if wl.get("heavy") is None else wl["heavy"][lo:hi])


def item_geometry(n: int):
    """Work items of one n-haplotype window: [(rows of the block, columns of the item), ...] (common.cuh)."""
    out = []
    for bi in range((n + TILE_M - 1) // TILE_M):
        rng = n - bi * TILE_M
        cnt = (rng + TILE_N - 1) // TILE_N
        width = ((rng + cnt - 1) // cnt + 15) & ~15
        out += [(TILE_M, width)] * cnt
    return out


def cpu_oracle(x_bits, node_len, labels, n, m_pad, pitch, length, threads, windows):
    """One pass of the plain-C oracle port (oracle/csrc/oracle_impop.c, pthreads) over `windows` same-shape windows."""
    from oracle import clib
    ar = np.arange(windows, dtype=np.int64)
    t0 = time.perf_counter()
    st, ct = clib.batch_stats(np.full(windows, n), np.full(windows, m_pad), np.full(windows, pitch), ar * (n * pitch),
                              ar * m_pad, np.zeros(windows, dtype=np.int64), np.full(windows, length),
                              x_bits[:windows], node_len[:windows], labels, threads)
    return time.perf_counter() - t0, st, ct


def max_rel_errors(got, want):
    """Largest plain relative error |got - want| / |want| per statistic over the rows (NaN rows must agree)."""
    from impop_b200._native import ST
    out = {}
    for name in ("pi", "pi_per_site", "pi_a", "pi_b", "dxy", "da", "fst", "tajima_d"):
        g, w = got[:, ST[name]], want[:, ST[name]]
        ok = np.isfinite(w) & (w != 0)
        err = np.abs(g[ok] - w[ok]) / np.abs(w[ok]) if ok.any() else np.zeros(1)
        out[name] = float(err.max()) if err.size else 0.0
    out["nan_pattern_equal"] = bool(np.array_equal(np.isnan(got), np.isnan(want)))
    return out


# ------------------------------------------------------------------------------------------------------------------
# Reference arm: the CPU restatement of the reference path on the same workload
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU path for this workload = `odgi/impg similarity` per window + the
    scripts' reductions.  Neither the external binaries nor /root/reference travel to the GPU box, so the arm times the
    plain-C oracle port of that path (oracle/csrc/oracle_impop.c) with every host thread, over ALL windows of the
    workload per step (same inputs as the GPU arm: compacted columns)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch  # noqa: F401  (host tensor plumbing for the generator only)
    from impop_b200 import synth
    cfg = CONFIGS[args.config if args.config in CONFIGS else 2]
    threads = host_threads()
    W = args.windows or cfg["W"]
    if args.config == 5:
        W = min(W, 1)                         # 5e7 pairs x 5 888 nodes per window: one window per step is ~minutes of CPU
    wl = make_workload(synth.HostGenerator(), cfg, W, cfg["seed"], threads, affine=False)       # plain compaction: constant columns merged

    def step():
        return cpu_oracle(wl["x"], wl["len"], wl["labels"], wl["n"], wl["m_pad"], wl["pitch"], wl["L"], threads, W)[0]

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = W * units_per_window(wl["n"], wl["L"]) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.config == 5 or args.strong else "weak", "vs_baseline": None,
        "dtype": "int64 intersections + f64 statistics", "data": "synthetic",
        "config": {"workload": cfg["workload"], "windows_per_step": W, "nodes_per_window": wl["m_in"],
                   "nodes_after_ingest": int(wl["m_out"].max()), "haplotypes": wl["n"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"all {W} windows of the workload per step, plain-C oracle port (byte-LUT intersections) on {threads} pthreads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# Secondary configs (short; embedded in the default line so that the driver's record carries them)
# ------------------------------------------------------------------------------------------------------------------
def _timed(torch, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_sites(ctx, torch, peaks, sites=10_000_000, reps=10, rank=0, world=1):
    """Config 4: per-site allele counts and frequencies, `sites` sites x 466 haplotypes x 5 panels (this rank's share)."""
    from impop_b200 import synth
    dev = ctx.torch_device
    mine = sites // world + (1 if rank < sites % world else 0)
    small, masks = synth.make_site_matrix(1 << 20, N_HAP, seed=0xB200 + 3 + rank)
    ds = torch.from_numpy(small.view(np.int64)).to(dev).repeat((mine + (1 << 20) - 1) // (1 << 20), 1)[:mine].contiguous()
    dm = torch.from_numpy(masks[:5].view(np.int64)).to(dev)
    counts = torch.empty((mine, 5), dtype=torch.int32, device=dev)
    freq = torch.empty((mine, 5), dtype=torch.float64, device=dev)
    ms = _timed(torch, lambda: ctx.site_counts(ds, dm, out_counts=counts, out_freq=freq), reps)
    bytes_alg = mine * (64 + 5 * 4 + 5 * 8)
    return {"sites": sites, "sites_this_gpu": mine, "ms": ms, "sites_per_s_per_gpu": mine / (ms * 1e-3),
            "algorithmic_bytes": bytes_alg, "achieved_gbs": bytes_alg / (ms * 1e-3) / 1e9,
            "hbm_frac": bytes_alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "bound": "hbm", "kernel": "site_counts_w8_kernel"}


def bench_windows_short(ctx, torch, peaks, cfg, windows, reps, threads):
    """Resident timing of the fused path on `windows` windows of a config (this GPU only), with the int8 / fp64 fractions."""
    from impop_b200.engine import ALGO_TCGEN05, WindowBatch
    dev = ctx.torch_device
    wl = make_workload(ctx, cfg, windows, cfg["seed"], threads)
    dw = DeviceWindows(torch, wl, dev)
    lab = torch.from_numpy(wl["labels"]).to(dev)
    b = dw.batch(ctx, WindowBatch, wl, lab, 0, windows)
    ms = _timed(torch, lambda: b.stats(ALGO_TCGEN05), reps)
    ctx.timing(True)
    b.stats(ALGO_TCGEN05)
    torch.cuda.synchronize()
    pairs_ms, prep_ms = ctx.timing_read("pairs")[0], ctx.timing_read("prep")[0]
    ctx.timing(False)
    ctx.check()
    b.close()
    n, L = wl["n"], wl["L"]
    ops = 2.0 * windows * (n * (n + 1) / 2.0) * wl["m_pad_in"] * wl["planes"]
    sms = torch.cuda.get_device_properties(ctx.device).multi_processor_count
    eps = windows * n * (n - 1) / 2.0 / (pairs_ms * 1e-3)
    return {"windows": windows, "haplotypes": n, "window_bp": L, "nodes_per_window": wl["m_in"], "nodes_after_ingest": int(wl["m_out"].max()),
            "ms_per_pass": ms, "pairs_kernel_ms": pairs_ms, "prep_ms": prep_ms,
            "hap_pair_bp_per_s": windows * units_per_window(n, L) / (ms * 1e-3),
            "int8_tops_algorithmic": ops / (pairs_ms * 1e-3) / 1e12, "int8_frac": ops / (pairs_ms * 1e-3) / 1e12 / peaks["int8_tops"],
            "fp64_frac": eps * 18 / (64.0 * sms * peaks["sm_max_mhz"] * 1e6), "ingest_s": wl["ingest_s"]}


def bench_cli_latency(reps=3, sizes=((90, 100_000), (466, 50_000))):
    """Config 1 (and its 466-haplotype sibling): wall time of the drop-in command lines `scripts/pica2.py`, `h-fst.py`, `af.py`
    on one window's all-pairs table (TSV mode: the table the similarity tool prints, written here from the matrix-mode
    dump).  Process start to exit, i.e. what a wrapper's per-window loop pays (run_pica2_impg.sh:175, run_h-fst.sh:74-85).
    The unmodified reference scripts on the same tables, timed in the build container: profiles/r2_reference_scripts.json
    (the reference tree does not travel to the GPU box)."""
    import tempfile
    from impop_b200 import synth
    from impop_b200.engine import Context, WindowBatch
    ref = {}
    try:
        rj = json.load(open(os.path.join(ROOT, "profiles", "r2_reference_scripts.json")))
        for case in rj["cases"].values():
            ref[case["haplotypes"]] = {k.split()[0]: v["single_process_wall_s"] for k, v in case["scripts"].items() if "-r 5" not in k}
    except Exception:
        pass
    ctx = Context(0)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for n, L in sizes:
            ws = synth.make_windows(n, L, 1, seed=0xB200 + (0 if n == 90 else 1))
            names = synth.haplotype_names(n, "chr2", 109_000_000, 109_000_000 + L)
            b = WindowBatch.from_uniform(ctx, ws.x_bits, ws.node_len, np.full(n, 9, dtype=np.uint8), L)
            _, _, pi = b.pairwise(0)
            ident = (1.0 - pi).cpu().numpy()
            b.close()
            tsv = os.path.join(tmp, f"n{n}.sim.tsv")
            with open(tsv, "w") as fh:
                fh.write("group.a\tgroup.b\testimated.identity\n")
                for i in range(n):
                    fh.write("".join(f"{names[i]}\t{names[j]}\t{float(ident[i, j])!r}\n" for j in range(i + 1, n)))
            pops, _ = synth.panel(n)
            asm = synth.assembly_names(range(n))
            fa, fb = os.path.join(tmp, f"n{n}.a.txt"), os.path.join(tmp, f"n{n}.b.txt")
            open(fa, "w").write("\n".join(a for a, p in zip(asm, pops) if p == 0) + "\n")
            open(fb, "w").write("\n".join(a for a, p in zip(asm, pops) if p == 2) + "\n")
            case = {}
            for script, extra in (("pica2.py", [tsv, "-t", "1.0", "-l", str(L), "-d", tmp]),
                                  ("h-fst.py", [tsv, "-a", fa, "-b", fb, "-l", str(L), "-d", tmp]),
                                  ("af.py", ["--input", tsv, "--threshold", "0.9995", "--output", os.path.join(tmp, "af.out")])):
                ts = []
                for _ in range(reps):
                    t0 = time.perf_counter()
                    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *extra], capture_output=True, text=True)
                    ts.append(time.perf_counter() - t0)
                case[script] = {"wall_s_median": float(np.median(ts)), "wall_s_min": float(min(ts)), "rc": r.returncode,
                                "stdout": r.stdout.strip().splitlines()[-1][:80] if r.stdout.strip() else "",
                                "reference_wall_s_build_container": ref.get(n, {}).get(script)}
            out[f"n{n}"] = case
    ctx.close()
    return out


def other_configs(ctx, torch, peaks, threads):
    out = {}
    for key, fn in (("config3", lambda: bench_windows_short(ctx, torch, peaks, CONFIGS[3], CONFIGS[3]["W"], 5, threads)),
                    ("config4", lambda: bench_sites(ctx, torch, peaks)),
                    ("config5", lambda: bench_windows_short(ctx, torch, peaks, CONFIGS[5], 4, 3, threads)),
                    ("config1", bench_cli_latency)):
        try:
            t0 = time.perf_counter()
            out[key] = fn()
            out[key]["measured_in_s"] = time.perf_counter() - t0
        except Exception as exc:  # a secondary measurement must never take the headline down
            out[key] = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------
# Our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from impop_b200.distributed import gather_parts, gather_rows
    from impop_b200.engine import ALGO_SIMT, ALGO_TCGEN05, Context, WindowBatch, NCOUNTS, NSTATS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = None
    if world > 1:
        torch.cuda.set_device(local)
        # host threads (and with them the first-touch placement of this rank's pinned buffers) on the CPUs next to this
        # rank's GPU: all ranks' host-to-device copies otherwise leave from whichever NUMA node the launcher started on
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            numa = sorted(os.sched_getaffinity(0))
            numa = f"{len(numa)} CPUs {numa[0]}-{numa[-1]}"
        except Exception as exc:
            numa = f"not set ({type(exc).__name__})"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    dev = ctx.torch_device
    algo = ALGO_SIMT if args.algo == "simt" else ALGO_TCGEN05
    peaks = load_peaks()
    threads = max(1, host_threads() // max(1, world))          # the ranks of one box share its host cores
    sms = int(torch.cuda.get_device_properties(local).multi_processor_count)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == 4:
        return run_sites(args, ctx, torch, dist, peaks, rank, world, local)
    if args.config == 1:
        if rank == 0:
            r = bench_cli_latency(5)
            print(json.dumps({"metric": "wall seconds of scripts/pica2.py on a 90-haplotype similarity table", "value": r["n90"]["pica2.py"]["wall_s_median"],
                              "unit": "s", "higher_is_better": False, "n_gpus": 1, "steps": 5, "warmup": 0, "data": "synthetic",
                              "config": {"workload": "pica2.py nucleotide diversity, one EDAR-shaped 100 kb window, 90 haplotypes, similarity TSV input", "baseline_config": 1},
                              "command_lines": r}))
        return 0

    cfg = CONFIGS[args.config]
    split = args.config == 5                                   # one batch, tile grid split over the ranks
    strong = split or args.strong
    W_total = args.windows or (cfg.get("W_total", cfg["W"]) if (strong and not split) else cfg["W"])
    n, L = cfg["n"], cfg["L"]

    # ---------------------------------------------------------------- synthetic batch -> ingest -> resident in HBM
    if split:
        # rank 0 holds the windows; X, node lengths and labels are broadcast once (NCCL), every rank builds the
        # same batch and works the items t with t % world == rank
        W = W_total
        keep = min(W, 1)
        if rank == 0:
            wl = make_workload(ctx, cfg, W, cfg["seed"], host_threads(), keep_original=0)
            host = DeviceWindows(torch, wl, "pinned", tight=args.transfer == "tight")
            meta = [wl["m_in"], wl["m_pad_in"], wl["planes"], int(wl["m_out"].max()), int(wl["k_exec"].max()), wl["pitch"],
                    int(wl["x_off"][-1]), int(wl["len_off"][-1])]
            tabs = np.stack([wl["m_out"].astype(np.int64), wl["pitch_w"].astype(np.int64), wl["x_off"][:-1], wl["len_off"][:-1],
                             wl["win_const"], wl["site_runs"]])
        else:
            wl, host, meta, tabs = None, None, [0] * 8, np.zeros((6, W), dtype=np.int64)
        mt, tb = torch.tensor(meta, dtype=torch.int64, device=dev), torch.from_numpy(tabs).to(dev)
        if world > 1:
            dist.broadcast(mt, 0); dist.broadcast(tb, 0)
        m_in, m_pad_in, planes, m_out_max, k_exec_max, pitch, xw, lw = (int(v) for v in mt.tolist())
        tabs = tb.cpu().numpy()
        if rank != 0:                          # the same descriptor tables on every rank; the bulk arrays arrive by broadcast
            wl = {"n": n, "L": L, "m_out": tabs[0].astype(np.int32), "pitch_w": tabs[1].astype(np.int32),
                  "x_off": np.concatenate([tabs[2], [xw]]), "len_off": np.concatenate([tabs[3], [lw]]),
                  "win_const": tabs[4].copy(), "site_runs": tabs[5].copy(), "labels": labels_for(cfg["labels"], np.zeros(n, dtype=np.int64))}
        lab_host = wl["labels"]
        res = object.__new__(DeviceWindows)
        res.x = torch.empty(xw, dtype=torch.int32, device=dev); res.len = torch.empty(lw, dtype=torch.int32, device=dev)
        res.row_adj = torch.empty(W * n, dtype=torch.int32, device=dev); res.col_mult = torch.empty(lw, dtype=torch.uint8, device=dev)
        res.xt = None
        if rank == 0:
            res.x.copy_(torch.from_numpy(wl["x"].reshape(-1).view(np.int32)))
            for k_ in ("len", "row_adj", "col_mult"):
                getattr(res, k_).copy_(getattr(host, k_))
        if world > 1:
            for k_ in ("x", "len", "row_adj", "col_mult"):
                dist.broadcast(getattr(res, k_), 0)
        k_exec = np.full(W, k_exec_max)
        ingest = {"seconds": wl["ingest_s"], "threads": wl["ingest_threads"], "bytes_in": wl["bytes_in"], "bytes_out": wl["bytes_out"]} if rank == 0 else None
        bounds = None
        m_pad = pitch * 32
    else:
        W = W_total // world + (1 if rank < W_total % world else 0) if strong else W_total
        keep = min(W, 96) if (rank == 0 and world == 1) else 0
        wl = make_workload(ctx, cfg, W, cfg["seed"] + 1000 * rank, threads, keep_original=keep,
                           plain=W if (rank == 0 and world == 1 and not args.no_cpu) else 0)
        pitch, m_pad, m_in, m_pad_in, planes = wl["pitch"], wl["m_pad"], wl["m_in"], wl["m_pad_in"], wl["planes"]
        m_out_max, k_exec = int(wl["m_out"].max()), wl["k_exec"]
        lab_host = wl["labels"]
        host = DeviceWindows(torch, wl, "pinned", tight=args.transfer == "tight")     # the ingested windows in pinned host memory (what the e2e leg uploads)
        res = DeviceWindows(torch, wl, dev)                # ... and resident in HBM (the device-timed leg)
        ingest = {"seconds": wl["ingest_s"], "threads": wl["ingest_threads"], "bytes_in": wl["bytes_in"], "bytes_out": wl["bytes_out"]}
        sizes = torch.tensor([W], dtype=torch.int64, device=dev)
        if world > 1:
            allw = [torch.zeros_like(sizes) for _ in range(world)]
            dist.all_gather(allw, sizes)
            bounds = np.concatenate([[0], np.cumsum([int(t.item()) for t in allw])]).astype(np.int64)
        else:
            bounds = np.array([0, W], dtype=np.int64)
    labels = torch.from_numpy(lab_host).to(dev)
    resident_mb = res.nbytes / 1e6
    l2_note = (f"{W} windows of {n * pitch * 4 / 1e6:.1f} MB each ({resident_mb:.0f} MB per GPU), read every step" if split
               else f"inputs larger than L2 ({resident_mb:.0f} MB per GPU read every step)")
    batch = res.batch(ctx, WindowBatch, wl, labels, 0, W, with_runs=not split)
    stats = torch.empty((W, NSTATS), dtype=torch.float64, device=dev)
    counts = torch.empty((W, NCOUNTS), dtype=torch.int64, device=dev)

    def step():
        if split:                              # partial sums of this rank's items, one all-gather, rank-ordered finalize
            sums = batch.window_sums(rank, world, algo)
            st, ct = batch.finalize(gather_parts(sums))
            stats.copy_(st); counts.copy_(ct)
            return stats
        batch.stats(algo, out_stats=stats, out_counts=counts)
        if world > 1:                          # the single result gather of the north star (W x 20 fp64 per rank)
            return gather_rows(stats, bounds)
        return stats

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
    t = torch.tensor([ms_step, pairs_ms / max(pairs_n, 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, pairs_avg_ms = float(t[0].item()), float(t[1].item())
    tot = torch.tensor([float(W)], dtype=torch.float64, device=dev)
    if world > 1 and not split:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    windows_job = int(tot.item())
    units_step = windows_job * units_per_window(n, L)
    value = units_step / (ms_step * 1e-3)

    # ---------------------------------------------------------------- e2e: host buffers through the public API
    # Every step: pinned host -> device copies of that step's inputs (the ingested matrices), batch set-up, the fused kernels
    # and the device -> host read of the result rows, all inside the timed region.  Window-sharded configs cut the batch
    # into sub-batches that alternate between two streams so copies overlap kernels; the split config uploads on rank 0,
    # broadcasts over NVLink and reads the finalized rows back.
    hs = torch.empty((W, NSTATS), dtype=torch.float64, pin_memory=True)
    hc = torch.empty((W, NCOUNTS), dtype=torch.int64, pin_memory=True)
    hlab = torch.from_numpy(lab_host).pin_memory()
    dwin = res.empty_like(torch, dev, xt_words=wl["xt_off"][-1] if (args.transfer == "tight" and (not split or rank == 0)) else 0)   # the e2e leg's own device buffers
    ds, dc = torch.empty_like(stats), torch.empty_like(counts)
    pending = []          # batches of the previous step: closed while this step's copies and kernels run
    host_busy = [0.0]
    if split:
        dlab = torch.empty_like(labels)
        nsub = 1

        def e2e_step():
            if rank == 0:
                dwin.copy_from(ctx, host, wl, 0, W)
                dwin.repitch(ctx, wl, 0, W)
            dlab.copy_(hlab, non_blocking=True)
            if world > 1:
                for k_ in ("x", "len", "row_adj", "col_mult"):
                    dist.broadcast(getattr(dwin, k_), 0)
            b = dwin.batch(ctx, WindowBatch, wl, dlab, 0, W, with_runs=False)
            sums = b.window_sums(rank, world, algo)
            st, ct = b.finalize(gather_parts(sums))
            hs.copy_(st, non_blocking=True); hc.copy_(ct, non_blocking=True)
            torch.cuda.synchronize()
            b.close()
    else:
        nsub = max(1, min(args.sub_batches or (4 if world <= 2 else 8), W))
        if nsub < 4:
            cuts = [int(v) for v in np.linspace(0, W, nsub + 1)]
        else:
            # the last two sub-batches are smaller: what is left to compute after the final copy lands is the step's tail
            wts = np.ones(nsub); wts[-2], wts[-1] = 0.6, 0.3
            cuts = [0] + [int(v) for v in np.round(np.cumsum(wts) / wts.sum() * W)]
            cuts[-1] = W
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        dlabs = [torch.empty_like(labels) for _ in range(nsub)]

        def e2e_step():
            t_in = time.perf_counter()
            live = []
            for k in range(nsub):
                lo, hi = cuts[k], cuts[k + 1]
                if hi <= lo:
                    continue
                st = streams[k % len(streams)]
                with torch.cuda.stream(st):
                    # this sub-batch's copies first, its set-up (host-side tables + their small upload) while they run: the copy
                    # engine is the bottleneck of the step and must never wait for the host
                    dlabs[k].copy_(hlab, non_blocking=True)
                    dwin.copy_from(ctx, host, wl, lo, hi)
                    b = dwin.batch(ctx, WindowBatch, wl, dlabs[k], lo, hi, stream=st, host=host)
                    dwin.repitch(ctx, wl, lo, hi)
                    b.stats(algo, stream=st, out_stats=ds[lo:hi], out_counts=dc[lo:hi])
                    hs[lo:hi].copy_(ds[lo:hi], non_blocking=True)
                    hc[lo:hi].copy_(dc[lo:hi], non_blocking=True)
                live.append(b)
            while pending:                      # host work hidden behind the copies just enqueued
                pending.pop().close()
            host_busy[0] += time.perf_counter() - t_in         # host time spent enqueueing (the rest of a step waits for the GPU)
            for st in streams:
                st.synchronize()
            pending.extend(live)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(3):
        e2e_step()
    barrier()
    host_busy[0] = 0.0
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
    # the e2e run must reproduce the resident run: same NaN pattern, every statistic within 1e-12 of the column's scale
    # (the per-lane sums of the pairs kernel are plain fp64, so the last bits depend on how a batch is dealt to the CTAs)
    ha, hb = hs.numpy(), stats.cpu().numpy()
    scale = np.nanmax(np.abs(hb), axis=0, keepdims=True)
    scale = np.where(np.isfinite(scale), scale, 0.0)
    with np.errstate(invalid="ignore"):
        close = np.abs(ha - hb) <= 1e-12 * np.maximum(scale, np.maximum(np.abs(ha), np.abs(hb)))
    same = bool(np.array_equal(np.isnan(ha), np.isnan(hb)) and bool(np.all(close | np.isnan(hb)))
                and torch.equal(hc, counts.cpu()))
    bitwise = bool(torch.equal(hs.nan_to_num(7.0), stats.cpu().nan_to_num(7.0)))
    if split:
        h2d = res.nbytes + 8 * W + world * hlab.numel()          # uploaded once (rank 0), broadcast over NVLink
        d2h = world * (hs.numel() * 8 + hc.numel() * 8)
    else:
        # whole job: every rank copies its own windows (presence bits, weights, row terms, multiplicities, constants, labels)
        hb = torch.tensor([float(host.nbytes + 8 * W + hlab.numel() * nsub)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(hb, op=dist.ReduceOp.SUM)
        h2d = int(hb.item())
        d2h = windows_job * (NSTATS * 8 + NCOUNTS * 8)

    # ---------------------------------------------------------------- roofline of the dominant kernel
    # algorithmic int8 ops from the INPUT node count and byte planes; executed ops = what the tensor pipe really
    # multiplies (128 x item width x virtual columns after ingest, every item); fp64 = 18 instructions per pair on a pipe of
    # 64 lanes / clk / SM, the binding pipe of configs 2 and 3 (an fp64 instruction also holds the sub-partition's issue port
    # for two cycles: DESIGN.md 4.2).
    w_launch = W                                                                     # windows one launch of this rank touches
    share = (1.0 / world) if split else 1.0                                          # ... and the share of their items it works
    pairs_s = pairs_avg_ms * 1e-3
    ops_launch = 2.0 * w_launch * (n * (n + 1) / 2.0) * m_pad_in * planes * share
    geom = item_geometry(n)
    exec_ops = 2.0 * float(sum(r * c for r, c in geom)) * float(np.sum(k_exec)) * share
    bytes_launch = w_launch * (n * (m_pad_in // 8) + 4 * m_pad_in + n + 8 * 14) * (1.0 if not split else 1.0)
    eps = w_launch * n * (n - 1) / 2.0 * share / pairs_s
    mhz_max = peaks["sm_max_mhz"]
    fp64_peak_eps = 64.0 * sms * mhz_max * 1e6 / 18.0
    tensor = {"algorithmic_ops_per_launch": ops_launch, "executed_ops_per_launch": exec_ops, "achieved": ops_launch / pairs_s / 1e12,
              "achieved_executed": exec_ops / pairs_s / 1e12, "peak": peaks["int8_tops"], "unit": "TOP/s (int8)",
              "frac": ops_launch / pairs_s / 1e12 / peaks["int8_tops"], "frac_executed": exec_ops / pairs_s / 1e12 / peaks["int8_tops"],
              "peak_source": peaks["int8_src"], "byte_planes": planes}
    fp64 = {"achieved": eps, "peak": fp64_peak_eps, "unit": "pair epilogues/s", "frac": eps / fp64_peak_eps, "instr_per_pair": 18,
            "lanes_per_clk_per_sm": 64, "sms": sms, "clock_mhz": mhz_max,
            "peak_source": "64 fp64 lanes / clk / SM (tools/micro/pi_bench.cu: 2.0 cycles per warp DFMA and sub-partition) x SMs x max SM clock / 18 instructions per pair"}
    hbm = {"algorithmic_bytes_per_launch": bytes_launch, "achieved_gbs": bytes_launch / pairs_s / 1e9, "peak_gbs": peaks["hbm_gbs"],
           "frac": bytes_launch / pairs_s / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["hbm_src"]}
    # the binding pipe: fp64 (pair epilogues) unless the tensor pipe is the busier one on what it really executes (before the
    # affine compaction config 5 was tensor-bound; with a fifth of its columns left it is fp64-bound like the others)
    bound = "tensor" if tensor["frac_executed"] > fp64["frac"] else "fp64"
    head = tensor if bound == "tensor" else fp64
    roofline = {"bound": bound, "achieved": head["achieved"], "peak": head["peak"], "unit": head["unit"], "frac": head["frac"],
                "traffic": None, "kernel": "window_pairs_tc_kernel" if algo == ALGO_TCGEN05 else "window_pairs_simt_kernel",
                "kernel_ms": pairs_avg_ms, "tensor": tensor, "fp64": fp64, "hbm": hbm,
                "frac_round1_formula": tensor["frac"],          # round 1 reported the algorithmic int8 fraction as `frac`
                "step_share": {k: v / ms_step for k, v in per_kernel.items()}}
    traffic_file = os.path.join(ROOT, "profiles", "pairs_traffic.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            roofline["traffic"] = tj["dram_bytes_per_window"] * w_launch * share
            roofline["traffic_source"] = "static: " + tj.get("source", "ncu dram__bytes_read.sum + dram__bytes_write.sum of the pairs kernel, scaled to this launch")
            if "step_dram_bytes_per_window" in tj:
                roofline["traffic_whole_step"] = tj["step_dram_bytes_per_window"] * w_launch
        except Exception:
            pass

    # ---------------------------------------------------------------- CPU baseline + parity on a sample (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not split:
        from oracle.compare import rows_close
        ht = host_threads()
        S0 = min(W, max(ht * 2, 8))
        # the oracle port takes the same windows in the PLAIN compacted form (constant columns merged): it walks node
        # columns, as the similarity tools walk path steps, and knows nothing of the affine form the GPU arm is given
        px, plen = wl["plain_x"], wl["plain_len"]
        p_pitch, p_mpad = int(px.shape[2]), int(plen.shape[1])
        dt, _, _ = cpu_oracle(px, plen, lab_host, n, p_mpad, p_pitch, L, ht, S0)
        S = int(min(W, max(S0, S0 * 12.0 / max(dt, 1e-3))))
        dt, st_cpu, ct_cpu = cpu_oracle(px, plen, lab_host, n, p_mpad, p_pitch, L, ht, S)
        rate = S * units_per_window(n, L) / dt
        got_s, got_c = stats[:S].cpu().numpy(), counts[:S].cpu().numpy()
        ok_counts = bool((got_c == ct_cpu).all())
        ok_stats, why = rows_close(got_s, st_cpu)
        # the same windows with their ORIGINAL columns (no ingest-time compaction) through the oracle
        K = min(keep, S)
        _, st_orig, ct_orig = cpu_oracle(wl["orig_x"], wl["orig_len"], lab_host, n, m_pad_in, m_pad_in // 32, L, ht, K)
        errs = max_rel_errors(got_s[:K], st_orig)
        strict = bool(errs["nan_pattern_equal"] and max(errs[k] for k in ("pi", "pi_per_site", "pi_a", "pi_b", "dxy", "da", "fst", "tajima_d")) <= 1e-12)
        cpu = {"value": rate, "unit": UNIT, "cores": ht, "kind": "port",
               "sample": f"{S} of {W} windows of this workload (constant columns merged: {wl['plain_m_out']} of {m_in} nodes) in {dt:.2f} s, plain-C oracle port (byte-LUT intersections) on {ht} pthreads",
               "gpu_matches_oracle_on_sample": {"counts_exact": ok_counts, "stats_within_1e-12": ok_stats, "detail": why},
               "gpu_vs_oracle_on_original_columns": {"windows": K, "counts_exact": bool((got_c[:K] == ct_orig).all()),
                                                     "variant_sites_exact": bool((got_s[:K, 19] == st_orig[:, 19]).all()),
                                                     "max_plain_relative_error": errs, "strict_1e-12_without_scale_policy": strict}}

    others = None
    if rank == 0 and world == 1 and args.config == 2 and not args.no_others:
        batch.close()
        del res, dwin, host
        torch.cuda.empty_cache()
        others = other_configs(ctx, torch, peaks, host_threads())

    if rank == 0:
        f64 = roofline["fp64"]
        mhz = (clocks or {}).get("sm_mhz")
        if mhz:
            f64["frac_at_sampled_clock"] = f64["achieved"] * 18 / (64.0 * sms * mhz * 1e6)
        scaling = "strong" if strong else "weak"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "u8 x u8 -> s32 intersections (tcgen05 kind::i8) + f64 statistics" if algo == ALGO_TCGEN05 else "u8 dp4a -> u32 + f64 statistics",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "baseline_config": args.config,
                       "windows_per_gpu": W, "windows_job": windows_job, "nodes_per_window": m_in, "nodes_after_ingest": m_out_max,
                       "haplotypes": n,
                       "parallelism": (f"tile grid of every window split over {world} GPU(s): X broadcast once (NCCL), partial sums all-gathered, rank-ordered finalize"
                                       if split else f"windows sharded over {world} GPU(s), one all-gather of result rows per step"),
                       "l2": l2_note,
                       "algo": args.algo, "ingest": ingest, "rank0_cpu_affinity": numa},
            "e2e": {"value": units_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "sub_batches": nsub, "transfer": args.transfer,
                    "host_enqueue_ms_per_step": host_busy[0] / e2e_steps * 1e3, "matches_resident_run": same,
                    "bitwise_equal_to_resident_run": bitwise,
                    "h2d_bytes_per_step_without_ingest": int(windows_job * (n * (m_pad_in // 8) + m_pad_in * 4)) if not split else None},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        if others is not None:
            line["other_configs"] = others
        print(json.dumps(line))
    if others is None:
        batch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_sites(args, ctx, torch, dist, peaks, rank, world, local):
    """--config 4 as the headline: sites sharded over the ranks, no exchange (each rank keeps its own counts)."""
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = bench_sites(ctx, torch, peaks, reps=max(args.steps, 3), rank=rank, world=world)
    t = torch.tensor([r["ms"]], dtype=torch.float64, device=ctx.torch_device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        print(json.dumps({"metric": "variant sites/s (per-site allele counts and frequencies, 5 panels)", "value": r["sites"] / (ms * 1e-3),
                          "unit": "sites/s", "n_gpus": world, "steps": max(args.steps, 3), "warmup": 3, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 popcounts -> i32 counts + f64 frequencies",
                          "data": "synthetic", "config": {"workload": "af per-site allele frequencies, 10^7 sites x 466 haplotypes x 5 panels", "baseline_config": 4},
                          "roofline": {"bound": "hbm", "achieved": r["achieved_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": r["hbm_frac"],
                                       "traffic": None, "kernel": r["kernel"], "algorithmic_bytes_per_launch": r["algorithmic_bytes"]},
                          "gpu_launches": max(args.steps, 3) + 3, "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json config (1-based); default 2 = the metric's config")
    ap.add_argument("--strong", action="store_true", help="config 3: the whole genome's windows split over the ranks (fixed total work)")
    ap.add_argument("--algo", default="tc", choices=["tc", "simt"])
    ap.add_argument("--windows", type=int, default=0, help="windows per GPU (weak) / in total (strong); default: the config's own")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-others", action="store_true", help="skip the short measurements of the other BASELINE configs")
    ap.add_argument("--sub-batches", type=int, default=0,
                    help="e2e leg: sub-batches alternating between two streams; default 4 on one or two GPUs (with more, the host's enqueue time per step, "
                         "e2e.host_enqueue_ms_per_step, exceeds the copy time) and 8 on four or eight (the copies share the host's memory bandwidth and are the slower side)")
    ap.add_argument("--transfer", default="tight", choices=["tight", "aligned"],
                    help="e2e leg: presence rows uploaded tight (ceil(m / 32) words, padded on the device by impop_repitch_rows) or as the kernels read them")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
