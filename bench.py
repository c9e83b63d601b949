#!/usr/bin/env python
"""bench.py — one JSON line for the MeShClust2 hot path on B200 (contract in the task prompt / DESIGN.md §Measurement).

A step = one pass of the hot path over one batch of synthetic mutated-template DNA:
    K1  k-mer histograms of every sequence (Loader<T>::get_point), then
    K2  the all-pairs feature + GLM + cutoff sweep with the reference's length prefilter (fastcar work() /
        the candidate scans of Trainer::get_close) -> survivor list.
`value` = pairs scored per second over the whole step with the packed sequences already resident in HBM;
`e2e`   = the same metric through the C ABI from HOST buffers (codes + segments in, survivors out), copies timed.
N > 1: one process per GPU (torchrun); sequences sharded for K1, histogram shards all-gathered over NCCL, the sweep
split by folded query-row blocks; total work fixed ("strong" scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2] [--n-seqs N]
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

METRIC = "sequence pairs scored/sec (features+GLM)"
UNIT = "pairs/s"
WEIGHTS = os.path.join(ROOT, "tests", "golden", "weights_cfg1_id90.txt")
WORKLOADS = {
    # BASELINE.json configs[2]: 100k sequences 1 kb, k=5, uint8, all-pairs feature+GLM sweep at 1/2/4/8 B200
    "cfg3": dict(synth="cfg3", desc="BASELINE configs[2]: 100k x 1 kb, k=5, uint8, all-pairs feature+GLM sweep (id 0.9)"),
    # BASELINE.json configs[1] shape: 10k 16S-like 1.5 kb (hot-path content of the train+cluster run: K1 + candidate scans)
    "cfg2": dict(synth="cfg2", desc="BASELINE configs[1] shape: 10k x 1.5 kb 16S-like, k=5, uint8, K1 + all-pairs candidate sweep (id 0.9)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def traffic_per_pair(key):
    """DRAM bytes per pair of a kernel from the committed ncu --set full capture (profiles/r1_traffic.json)"""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        d = json.load(open(p))
        for k, v in d.items():
            if k.startswith(key):
                return float(v["dram_bytes_per_pair"]), v["source"]
    except Exception:
        pass
    return None, None


def instr_per_pair(key):
    """executed warp-instructions per pair of a kernel from the committed ncu capture (profiles/r1_traffic.json)"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        for k, v in d.items():
            if k.startswith(key) and "warp_instr_per_pair" in v:
                return float(v["warp_instr_per_pair"]), v["source"]
    except Exception:
        pass
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML every 200 ms
    (the B200_PROFILING.md clocks line without spawning nvidia-smi, whose polling loop perturbs short steps)."""

    def __init__(self, device, interval=0.5):
        self.device, self.rows, self.stop_flag, self.t, self.interval = device, [], False, None, interval
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(device))
        except Exception:
            self.nv = None

    @staticmethod
    def _physical_index(device):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[device])
            except Exception:
                pass
        return device

    def _loop(self):
        nv = self.nv
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                # two light NVML reads per sample; NVML queries serialise with CUDA calls in the driver, so keep them sparse
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), mx, reasons(self.h)))
            except Exception:
                pass
            for _ in range(int(self.interval / 0.05)):
                if self.stop_flag:
                    break
                time.sleep(0.05)

    def start(self):
        if self.nv is None:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        if self.nv is None or self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = [r[0] for r in self.rows]
        smax = self.rows[-1][1] if self.rows else None
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in self.rows))
        busy = [x for x in sm if smax and x > 0.3 * smax] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": float(smax) if smax else None,
                "reasons": reasons, "samples": len(sm)}


def load_workload(name, n_override, lo, hi):
    from meshclust2_b200 import synth
    t0 = time.time()
    seqs, _, k, eb = synth.make_config_range(WORKLOADS[name]["synth"], lo=lo, hi=hi, n=n_override)
    log("[bench] generated %d synthetic sequences [%d,%d) in %.1fs" % (len(seqs), lo, hi, time.time() - t0))
    return seqs, k, eb


def host_bytes(enc):
    return int(enc["codes"].nbytes + enc["seq_off"].nbytes + enc["segs"].nbytes + enc["seg_off"].nbytes)


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own implementation compiled in place (oracle/_ref), else the C port
# ---------------------------------------------------------------------------------------------------------------
def cpu_leg(seqs, k, eb, n_total, n_scored_total, cutoff, hist_sample, pair_sample, seed=0):
    """Times Loader<T>::get_point (omp over sequences) and Predictor<T>::close (omp over pairs) on all host cores on a
    bounded sample of the workload; returns dict with the workload-equivalent pairs/s."""
    from oracle import port, ref
    threads = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    model_txt = open(WEIGHTS).read()
    hs = min(hist_sample, len(seqs))
    sample = [seqs[i] for i in rng.choice(len(seqs), hs, replace=False)] if hs < len(seqs) else list(seqs)
    if ref.available():
        kind = "reference"
        H, t_hist = ref.count_batch(sample, k, eb, threads=threads, want_hist=True)
        ln = np.array([len(s) for s in sample], dtype=np.uint64)   # synthetic ACGT: effective size == length
    else:
        kind = "port"
        encs = [port.encode(s) for s in sample]
        codes = np.concatenate([e[0] for e in encs])
        seq_off = np.concatenate([[0], np.cumsum([len(e[0]) for e in encs])])
        segs = np.concatenate([e[1].reshape(-1, 2) for e in encs])
        seg_off = np.concatenate([[0], np.cumsum([len(e[1].reshape(-1, 2)) for e in encs])])
        H, t_hist = port.count_batch(codes, seq_off, segs, seg_off, k, eb, threads=threads)
        ln = np.array([e[2] for e in encs], dtype=np.uint64)
    hist_rate = hs / t_hist
    # pairs inside the length window, as the sweep scores them
    order = np.argsort(ln, kind="stable")
    ia = rng.integers(0, hs, pair_sample * 2)
    ib = rng.integers(0, hs, pair_sample * 2)
    lo = (ln[ib].astype(np.float64) * cutoff).astype(np.uint64)
    hi = (ln[ib].astype(np.float64) / cutoff).astype(np.uint64)
    ok = (ln[ia] >= lo) & (ln[ia] <= hi)
    ia, ib = ia[ok][:pair_sample], ib[ok][:pair_sample]
    if kind == "reference":
        rm = ref.RefModel(model_txt, eb, cutoff)
        r = rm.score_pairs(H, None, ln, ia, ib, mode=1, threads=threads)      # Predictor<T>::close
        t_pairs, n_close = r["seconds"], int(r["close"].sum())
    else:
        m = port.Model.from_text(model_txt)
        mag = H.sum(axis=1, dtype=np.uint64)
        r = port.score_pairs(m, H, mag, ln, ia, ib, threads=threads, want_cache=False)
        t_pairs, n_close = r["seconds"], int(r["close"].sum())
    pair_rate = len(ia) / t_pairs
    t_step = n_total / hist_rate + n_scored_total / pair_rate
    del order
    return dict(value=n_scored_total / t_step, unit=UNIT, cores=threads, kind=kind,
                sample="%d sequences through Loader<T>::get_point (%.2fs) + %d in-window pairs through Predictor<T>::close "
                       "(%.2fs), %d OpenMP threads; extrapolated to the step's %d histograms + %d pairs" % (
                           hs, t_hist, len(ia), t_pairs, threads, n_total, n_scored_total),
                hist_per_s=hist_rate, pairs_per_s_kernel=pair_rate, seconds=t_hist + t_pairs, n_close_sample=n_close)


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_total = args.n or {"cfg3": 100000, "cfg2": 10000}[args.workload]
    nsamp = min(n_total, args.cpu_hist_sample)
    seqs, k, eb = load_workload(args.workload, args.n, 0, nsamp)
    n_scored_total = n_total * (n_total - 1) // 2          # synthetic lengths are within the 0.9 window (L +/- 5 %)
    vals, secs = [], []
    leg = None
    for s in range(args.warmup + args.steps):
        leg = cpu_leg(seqs, k, eb, n_total, n_scored_total, 0.9, nsamp, args.cpu_pair_sample, seed=s)
        if s >= args.warmup:
            vals.append(leg["value"])
            secs.append(leg["seconds"])
    v = float(np.mean(vals))
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "n_sequences": n_total, "k": k, "elem_bytes": eb,
                       "pairs_per_step": n_scored_total, "model": os.path.basename(WEIGHTS),
                       "note": "each step is a bounded sample; value is the workload-equivalent rate"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": leg["cores"], "kind": leg["kind"], "sample": leg["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from meshclust2_b200 import capi, dist as mdist
    import torch
    tdist = None
    if world > 1:
        import torch.distributed as tdist_mod
        torch.cuda.set_device(local_rank)
        tdist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        tdist = tdist_mod
    comm = mdist.Comm(tdist)
    n_total = args.n or {"cfg3": 100000, "cfg2": 10000}[args.workload]
    per, bounds = mdist.shard_bounds(n_total, world)
    lo, hi = bounds[rank]
    seqs, k, eb = load_workload(args.workload, args.n, lo, hi)
    t0 = time.time()
    enc = capi.encode_batch(seqs)
    log("[bench] rank %d host encode %.2fs" % (rank, time.time() - t0))
    ctx = capi.Context(local_rank)
    pin_inputs(capi, enc)
    model = ctx.model_from_file(WEIGHTS)
    cutoff = model.meta["id"]
    eng = mdist.GpuEngine(capi, ctx, torch, model, k, eb, local_rank)
    eng.set_local_sequences(ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"]), hi - lo, per)
    max_out = 1 << 22

    def step_resident():
        ctx.flush_l2(256 << 20)
        return mdist.all_pairs_step(eng, comm, torch, n_total, cutoff, upper_only=True, blocks_per_rank=args.blocks_per_rank,
                                    max_out=max_out)

    # end to end starts from RAW sequence text in page-locked host memory: H2D, N-run segmentation, letter coding and
    # packing (mc2_seqs_from_text_into), K1, exchange, sweep, D2H of the survivors are all inside the timed region
    text = np.frombuffer(bytearray(b"".join(seqs)), dtype=np.uint8) if seqs else np.zeros(1, dtype=np.uint8)
    text_off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    text_off[1:] = np.cumsum([len(s) for s in seqs])
    for arr in (text, text_off):
        try:
            capi.host_register(arr)
        except capi.Mc2Error as e:
            log("[bench] could not page-lock the text: %s" % e)

    def step_e2e():
        ctx.flush_l2(256 << 20)
        ctx.seqs_from_text_into(eng.seqs, text, text_off)
        return mdist.all_pairs_step(eng, comm, torch, n_total, cutoff, upper_only=True, blocks_per_rank=args.blocks_per_rank,
                                    max_out=max_out)

    def timed(fn, steps, sample_clocks=False):
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()          # before the barrier: rank 0 must not enter the timed region later than the others
        comm.barrier()
        torch.cuda.synchronize()
        ctx.sync()
        l0 = ctx.launches
        ctx.profile(True)
        ctx.timer_start()
        res = None
        for _ in range(steps):
            res = fn()
        ms = ctx.timer_stop()
        comm.barrier()
        torch.cuda.synchronize()
        ktime = {kind: ctx.kernel_time(kind) for kind in range(6)}
        ctx.profile(False)
        clocks = sampler.stop() if sampler else None
        ms = comm.all_reduce_max(ms, torch, eng.device)
        return ms, res, ctx.launches - l0, ktime, clocks

    for _ in range(args.warmup):
        res = step_resident()
    ms, res, launches, ktime, clocks = timed(step_resident, args.steps, sample_clocks=True)
    n_scored = res["n_scored"]
    value = n_scored * args.steps / (ms * 1e-3)
    log("[bench] rank %d resident: %.1f ms/step, %d pairs scored, %d close, %d launches" % (
        rank, ms / args.steps, n_scored, res["n_close"], launches))
    # end to end from host buffers
    log("[bench] rank %d e2e phase" % rank)
    for _ in range(min(args.warmup, 3)):
        step_e2e()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e, res_e, _, _, _ = timed(step_e2e, e2e_steps)
    e2e_value = res_e["n_scored"] * e2e_steps / (ms_e * 1e-3)
    log("[bench] rank %d e2e: %.1f ms/step" % (rank, ms_e / e2e_steps))
    h2d = int(text.nbytes + text_off.nbytes)
    d2h = int(len(res_e["survivors"]) * 24 + 16 * len(res_e["blocks"]))

    # roofline of the dominant kernel (the sweep): algorithmic bytes = candidate form, N*w + 24 + 9 per pair
    N = 4 ** k
    bytes_per_pair = N * eb + 24 + 9
    sweep_ms, sweep_n = ktime[3]
    count_ms, count_n = ktime[1]
    peak, peak_src = peaks()
    local_pairs = res["local_scored"] * args.steps
    achieved = local_pairs * bytes_per_pair / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0
    tpp, tsrc = traffic_per_pair("sweep_kernel")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": tpp * local_pairs / max(1, sweep_n) if tpp else None, "traffic_source": tsrc,
                "kernel": "sweep_kernel<u8,NEED_DOT|NEED_EMD>",
                "algorithmic_bytes_per_unit": bytes_per_pair, "units_per_launch": local_pairs / max(1, sweep_n),
                "avg_launch_ms": sweep_ms / max(1, sweep_n), "launches": sweep_n,
                "kernel_share_of_step": sweep_ms / ms if ms > 0 else None, "peak_source": peak_src,
                "note": "candidate-form bytes (each scored pair streams one database row, query row resident); at this "
                        "size the 4^k x n set is L2-resident, so the kernel is ALU-issue bound, see DESIGN.md"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "n_sequences": n_total, "k": k, "elem_bytes": eb,
                       "pairs_scored_per_step": n_scored, "pairs_close_per_step": res["n_close"],
                       "model": os.path.basename(WEIGHTS), "parallelism": "row-block x%d, NCCL all-gather of histograms" % world,
                       "l2": "L2 flushed between steps (256 MB memset inside the timed region)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": ms_e / e2e_steps,
                    "from": "raw sequence text in page-locked host memory (mc2_seqs_from_text_into -> mc2_count_kmers_into -> "
                            "mc2_all_pairs; survivors copied back)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "k1": {"hist_per_s": (hi - lo) * args.steps / (count_ms * 1e-3) if count_ms > 0 else None,
                   "avg_launch_ms": count_ms / max(1, count_n),
                   "achieved_gbs": (hi - lo) * args.steps * (250 + N * eb + 40) / (count_ms * 1e-3) / 1e9 if count_ms > 0 else None}}
    try:
        # the roofline that actually binds the L2-resident sweep ("not HBM -- CUDA-core integer issue rate"):
        # executed warp-instructions per second against the issue rate of the SMs at the clock sampled in the timed region
        ipp, isrc = instr_per_pair("sweep_kernel")
        sm_mhz = (clocks or {}).get("sm_mhz")
        if ipp and sm_mhz and sweep_ms > 0:
            issue_peak = ctx.sm_count * 4 * sm_mhz * 1e6
            issue_ach = local_pairs / (sweep_ms * 1e-3) * ipp
            line["roofline_issue"] = {"bound": "issue", "achieved": issue_ach / 1e9, "peak": issue_peak / 1e9,
                                      "unit": "Gwarp-instr/s", "frac": issue_ach / issue_peak, "warp_instr_per_pair": ipp,
                                      "source": isrc, "kernel": roofline["kernel"],
                                      "note": "secondary: issue slots of 4 schedulers x SMs at the sampled SM clock; the "
                                              "instruction count per pair is the committed ncu figure, pairs/s is live"}
    except Exception as e:
        log("[bench] issue roofline skipped: %r" % (e,))
    if world == 1 and not args.no_extras:
        line["roofline_candidates"] = candidates_roofline(ctx, capi, model, k, eb, peak, peak_src)
        if args.workload != "cfg2":
            line["also_cfg2"] = small_workload(ctx, capi, mdist, torch, model, cutoff)
    if rank == 0 and world == 1 and not args.no_cpu:
        t0 = time.time()
        cb = cpu_leg(seqs, k, eb, n_total, n_scored, cutoff, args.cpu_hist_sample, args.cpu_pair_sample)
        log("[bench] cpu baseline leg %.1fs" % (time.time() - t0))
        line["cpu_baseline"] = {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        emit(line)
    ctx.close()
    if tdist is not None:
        tdist.destroy_process_group()


def pin_inputs(capi, enc):
    """page-lock the step's host inputs (the encoded batch) so the timed H2D copies are DMA transfers from pinned memory"""
    for key in ("codes", "seq_off", "segs", "seg_off"):
        enc[key] = np.ascontiguousarray(enc[key])
        if enc[key].nbytes:
            try:
                capi.host_register(enc[key])
            except capi.Mc2Error as e:          # e.g. a locked-memory limit: the copies still work, just through pageable memory
                log("[bench] could not page-lock %s: %s" % (key, e))


def small_workload(ctx, capi, mdist, torch, model, cutoff, steps=20):
    """The BASELINE configs[1] shape in the same run (10k x 1.5 kb 16S-like, k=5, uint8): K1 + all-pairs candidate sweep,
    resident and end to end; the whole histogram set (10 MB) is L2-resident, steps are ~15 ms."""
    from meshclust2_b200 import synth
    seqs, _, k, eb = synth.make_config_range("cfg2", 0, None)
    enc = capi.encode_batch(seqs)
    pin_inputs(capi, enc)
    eng = mdist.GpuEngine(capi, ctx, torch, model, k, eb, 0)
    n = len(seqs)
    eng.set_local_sequences(ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"]), n, n)
    comm = mdist.Comm(None)
    text = np.frombuffer(bytearray(b"".join(seqs)), dtype=np.uint8)
    text_off = np.zeros(n + 1, dtype=np.uint64)
    text_off[1:] = np.cumsum([len(s) for s in seqs])
    for arr in (text, text_off):
        try:
            capi.host_register(arr)
        except capi.Mc2Error:
            pass
    out = {}
    for name in ("resident", "e2e"):
        def step():
            ctx.flush_l2(256 << 20)
            if name == "e2e":
                ctx.seqs_from_text_into(eng.seqs, text, text_off)
            return mdist.all_pairs_step(eng, comm, torch, n, cutoff, upper_only=True, blocks_per_rank=1, max_out=1 << 22)
        for _ in range(3):
            res = step()
        ctx.sync()
        ctx.timer_start()
        for _ in range(steps):
            res = step()
        ms = ctx.timer_stop() / steps
        out[name] = {"value": res["n_scored"] / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms}
    try:
        out["update_stage"] = update_stage(ctx, capi, model, eng.full, cutoff)
    except Exception as e:                              # an extra, never the reason a bench line is lost
        log("[bench] update-stage extra failed: %r" % (e,))
    out["workload"] = WORKLOADS["cfg2"]["desc"]
    out["pairs_scored_per_step"] = res["n_scored"]
    out["pairs_close_per_step"] = res["n_close"]
    out["hist_per_step"] = n
    return out


def update_stage(ctx, capi, model, hs, cutoff, delta=5, per_cluster=5, passes=5):
    """One pass of the mean-shift update stage (ClusterFactory.cpp:636-653) over the configs[1]-shaped point set: clusters of
    5 consecutive points (the synthetic set keeps a template's variants together), center = the cluster's first point,
    members of center j = the points of clusters j-delta..j+delta, merge candidates = the next delta centers.  Batched
    (mc2_update_centers + mc2_merge_centers: one device call each per pass) next to what the reference's per-center loops
    issue through the same C ABI (mc2_filter_as + mc2_mean_closest + mc2_merge per center).  Host wall clock, since the
    difference IS the per-call overhead; both forms return the same choices."""
    n = len(hs)
    got = hs.download()
    mag, ln = got["mag"].astype(np.uint64), got["len"].astype(np.uint64)
    nc = n // per_cluster
    rows = np.arange(nc, dtype=np.uint64) * per_cluster
    off = np.zeros(nc + 1, dtype=np.uint64)
    mem = []
    for j in range(nc):
        lo, hi = max(0, j - delta) * per_cluster, min(nc, j + delta + 1) * per_cluster
        mem.append(np.arange(lo, hi, dtype=np.uint64))
        off[j + 1] = off[j] + np.uint64(hi - lo)
    members = np.concatenate(mem)
    sc = ctx.hset_from_host(np.ones((nc, 4 ** hs.k), dtype=np.uint8), hs.k, length=np.ones(nc, dtype=np.uint64))
    idx = np.arange(nc, dtype=np.uint64)

    def batched():
        sc.assign_rows(idx, hs, rows, mag=mag[rows], length=ln[rows])
        nxt, ng = ctx.update_centers(model, sc, nc, hs, off, members, cutoff)
        mg = ctx.merge_centers(model, sc, nc, delta, cutoff)
        return nxt, mg

    def per_center(limit):
        nxt = np.full(limit, -1, dtype=np.int64)
        mg = np.zeros(limit, dtype=np.int64)
        sc.assign_rows(idx, hs, rows, mag=mag[rows], length=ln[rows])
        for j in range(limit):
            keep = ctx.filter_as(model, hs, int(rows[j]), int(mag[rows[j]]), int(ln[rows[j]]), hs, mem[j], cutoff).astype(bool)
            if keep.any():
                nxt[j] = np.flatnonzero(keep)[ctx.mean_closest(hs, mem[j][keep])[0]]
            mg[j] = ctx.merge(model, sc, idx, j, j + 1, min(nc - 1, j + delta), cutoff)
        return nxt, mg

    batched()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(passes):
        b = batched()
    ctx.sync()
    t_b = (time.perf_counter() - t0) / passes
    limit = min(nc, 500)                                 # a bounded sample of the per-center form, scaled to the pass
    per_center(8)
    t0 = time.perf_counter()
    p = per_center(limit)
    ctx.sync()
    t_p = (time.perf_counter() - t0) * nc / limit
    same = bool(np.array_equal(b[0][:limit], p[0]) and np.array_equal(b[1][:limit], p[1]))
    sc.free()
    return {"centers": nc, "member_pairs_per_pass": int(off[-1]), "batched_ms_per_pass": t_b * 1e3,
            "per_center_calls_ms_per_pass": t_p * 1e3, "speedup": t_p / t_b, "same_choices": same,
            "sample": "per-center form timed on the first %d centers and scaled to %d" % (limit, nc)}


def candidates_roofline(ctx, capi, model, k, eb, peak, peak_src, n=1 << 20):
    """The HBM-streaming form (Trainer::get_close over a long candidate range, BASELINE configs[4] shape):
    one query vs 2^20 candidate histograms (1 GiB > L2), device-timed per launch."""
    rng = np.random.default_rng(1)
    N = 4 ** k
    base = rng.integers(1, 7, size=(4096, N), dtype=np.uint8)
    H = base[rng.integers(0, 4096, n)]
    ln = rng.integers(950, 1050, n).astype(np.uint64)
    hs = ctx.hset_from_host(H, k, length=ln)
    del H
    ms, nclose = ctx.bench_score_pairs(model, hs, hs, n_pairs=n, a_begin=0, b_begin=3, b_bc=1, iters=20, flush_l2=False,
                                       len_filter=1, anchor_is_b=1, cutoff=0.9)
    ms, nclose = ctx.bench_score_pairs(model, hs, hs, n_pairs=n, a_begin=0, b_begin=3, b_bc=1, iters=20, flush_l2=False,
                                       len_filter=1, anchor_is_b=1, cutoff=0.9)
    hs.free()
    bpp = N * eb + 24 + 9
    ach = n * bpp / (ms * 1e-3) / 1e9
    tpp, tsrc = traffic_per_pair("pair_fast_kernel")
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": tpp * n if tpp else None, "traffic_source": tsrc,
            "kernel": "pair_fast_kernel<u8> one query vs 2^20 candidates (1 GiB streamed, > L2)",
            "algorithmic_bytes_per_unit": bpp, "units_per_launch": n, "avg_launch_ms": ms, "pairs_per_s": n / (ms * 1e-3),
            "peak_source": peak_src}


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else libraries print to fd 1 (e.g. NCCL's version banner)
    has been redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--n-seqs", dest="n", type=int, default=None, help="override the number of sequences (smoke runs)")
    ap.add_argument("--blocks-per-rank", type=int, default=2, help="folded query-row block pairs per rank and step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cpu-hist-sample", type=int, default=100000)
    ap.add_argument("--cpu-pair-sample", type=int, default=20000000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: W >= 3 warm-up steps are required for a valid number")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
