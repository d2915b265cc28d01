#!/usr/bin/env python
"""Benchmark of the LDWeaver hot path on B200: weighted SNP-pair MI/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm's CPU path (oracle port)

A "step" is one full weighted pairwise-MI scan (all make_blocks blocks: GEMM + fused MI epilogue + sr/lr link
filter + per-block exact LR selection + link-column materialisation) of the synthetic 616 x 100k alignment
(SURVEY.md 8d, config C2), blocks dealt by cost over the ranks.  `value` is pairs/s with the packed operands
already resident in HBM (device time from CUDA events on the library's stream, max over ranks); `e2e` is the same
metric through the C ABI from host buffers: host->device upload + operand packing + scan + device->host copy of
every link column, inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

SR_DIST = 20000.0
LR_RETAIN = 1e6
MAX_BLK = 10000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--nsnp", type=int, default=0, help="override the number of SNPs (debug only; invalidates the metric)")
    ap.add_argument("--cpu-sample", type=int, default=4000, help="block edge of the bounded CPU-baseline sample")
    ap.add_argument("--nrate", type=float, default=-1.0, help="override the per-cell N rate (debug only; invalidates the metric)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-post", action="store_true", help="skip the timing of the post-scan host steps")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d.get("hbm_gbs"), "bf16_tflops": d.get("bf16_tflops"),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  The sampler is
    started before the warm-up (nvidia-smi needs a few hundred ms to produce its first line) and only the samples whose
    timestamps fall inside the timed window are used; if the window is shorter than the sampling period the nearest
    samples (within 0.5 s, i.e. still under the same back-to-back load) are reported and flagged."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def window_begin(self):
        self.t0 = time.time()

    def window_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)  # let the sample that covers the end of the window land
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        import datetime
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        samples = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rs = [nm for k, nm in enumerate(names) if "Active" in r[4 + k] and "Not" not in r[4 + k]]
                samples.append((ts, float(r[1]), float(r[2]), rs))
            except Exception:
                pass
        t0, t1 = self.t0 or 0.0, self.t1 or 1e18
        inside = [s for s in samples if t0 <= s[0] <= t1]
        note = None
        if not inside:
            inside = [s for s in samples if t0 - 0.5 <= s[0] <= t1 + 0.5]
            note = f"timed window of {1e3 * (t1 - t0):.0f} ms is shorter than the sampling period: nearest samples (+-0.5 s, same back-to-back load)"
        out = {"sm_mhz": float(np.median([s[1] for s in inside])) if inside else None,
               "sm_max_mhz": max(s[2] for s in inside) if inside else None,
               "reasons": sorted({r for s in inside for r in s[3]}), "samples": len(inside)}
        if note:
            out["note"] = note
        return out


def cpu_reference_sample(sy, hdw, edge: int):
    """The reference algorithm's CPU path (oracle C port, all host threads) on a bounded sample of the same workload:
    one off-diagonal `edge` x `edge` block (25 weighted products + temporaries + fastHadamard, pair enumeration, len,
    type-7 quantile, filter).  Returns (pairs, seconds, threads)."""
    import c_oracle as CO
    n = sy.codes.shape[0]
    edge = min(edge, n // 2)
    f = np.arange(0, edge)
    t = np.arange(n - edge, n)
    table = np.stack([(sy.codes == a).any(axis=1) for a in range(5)], axis=1).astype(np.float64)
    r = table.sum(axis=1)
    t0 = time.perf_counter()
    MI = CO.block_mi(sy.codes, hdw, r, table, f, t)
    L = CO.block_links(MI, sy.POS.astype(np.float64), f, t, float(sy.g), SR_DIST, LR_RETAIN, 4.9e9)
    dt = time.perf_counter() - t0
    return len(L["MI"]), dt, CO.lib().ldwo_num_threads()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from ldweaver_b200 import synth

    S, n_cfg, seed, probs, nrate = synth.CONFIGS[args.config]
    n = args.nsnp or n_cfg
    if args.nrate >= 0:
        nrate = args.nrate
    workload = f"{args.config}: synthetic {S} seqs x {n} SNPs (seed {seed}), Hamming-weighted MI, sr_dist {int(SR_DIST)}, " \
               f"lr_retain_links {int(LR_RETAIN)}, max_blk_sz {MAX_BLK}"
    config = {"workload": workload, "nseq": S, "nsnp": n, "sr_dist": SR_DIST, "lr_retain_links": LR_RETAIN,
              "max_blk_sz": MAX_BLK, "partition": f"make_blocks blocks dealt by cost (pairs, largest first, least-loaded rank) over {world} rank(s)",
              "l2": "inputs larger than L2: operand planes + records are 170 MB (> 126 MB L2) and every step writes 2.9 GB of link columns, so no input survives in L2 between timed steps; no explicit flush"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        import c_oracle as CO
        sy = synth.generate(S, min(n, 4 * args.cpu_sample), seed, probs, nrate)
        hdw, _ = CO.hdw(sy.codes, 0.1)
        for _ in range(min(args.warmup, 1)):
            cpu_reference_sample(sy, hdw, min(500, args.cpu_sample))
        pairs = 0
        secs = 0.0
        for _ in range(args.steps):
            p_, dt, thr = cpu_reference_sample(sy, hdw, args.cpu_sample)
            pairs += p_
            secs += dt
        val = pairs / secs
        sample = f"one off-diagonal {args.cpu_sample}x{args.cpu_sample} block of the workload per step (reference-shaped C/OpenMP port)"
        line = {"impl": "reference", "metric": "weighted SNP-pair MI/sec", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": thr, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ldweaver_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL writes its version banner / debug lines to stdout; rank 0's stdout must hold the JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import ldweaver_b200 as ldw
    from ldweaver_b200 import api

    sy = synth.generate(S, n, seed, probs, nrate)
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    # population-structure weights on rank 0, broadcast over NVLink with NCCL
    hdw_t = torch.empty(S, dtype=torch.float64, device="cuda")
    t_hdw = None
    if rank == 0:
        t0 = time.perf_counter()
        hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1, device=local_rank)
        t_hdw = time.perf_counter() - t0
        hdw_t.copy_(torch.from_numpy(hdw))
    if world > 1:
        dist.broadcast(hdw_t, src=0)
    hdw = hdw_t.cpu().numpy()
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, SR_DIST)
    blk = api.round_half_even_thousands(MAX_BLK)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # pinned host copies of the inputs (what an R shim would hand over)
    codes_pin = torch.from_numpy(sy.codes).pin_memory().numpy()
    plan = ldw.MIPlan(ldw.snp_dat_from_codes(codes_pin, sy.POS, sy.g), hdw, sy.paint, blk, device=local_rank)
    flags_dev = api.SCAN_NO_D2H
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        plan.scan(sy.g, SR_DIST, LR_RETAIN, lra, flags_dev, world, rank, copy=False)
    barrier()
    sampler.window_begin()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    kern_ms = 0.0
    step_ms = []
    stats = None
    agg = {"n_pairs": 0, "n_launches": 0, "n_scan_launches": 0, "exec_int8_ops": 0.0, "n_tiles": 0}
    for _ in range(args.steps):
        *_, stats = plan.scan(sy.g, SR_DIST, LR_RETAIN, lra, flags_dev, world, rank, copy=False)
        dev_ms += stats["t_scan_ms"] + stats["t_select_ms"]
        step_ms.append(stats["t_scan_ms"] + stats["t_select_ms"])
        kern_ms += stats["t_kernel_ms"]
        for k in agg:
            agg[k] += stats[k]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.window_end()
    clocks = sampler.stop()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dev_ms_max = allmax(dev_ms)
    wall_max = allmax(t_wall)
    pairs_all = allsum(float(agg["n_pairs"]))  # over all ranks and steps
    launches_all = allsum(float(agg["n_launches"]))
    value = pairs_all / (dev_ms_max * 1e-3)

    # ------------------------------------------------------------------ end to end through the C ABI from host buffers
    e2e = None
    if not args.no_e2e:
        snp_pin = ldw.snp_dat_from_codes(codes_pin, sy.POS, sy.g)
        e_steps = max(1, min(args.steps, 3))
        for _ in range(1):  # warm-up: pinned output buffers get allocated once
            p2 = ldw.MIPlan(snp_pin, hdw, sy.paint, blk, device=local_rank)
            p2.scan(sy.g, SR_DIST, LR_RETAIN, lra, 0, world, rank, copy=False)
            p2.close()
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        pe = 0
        t_plan = t_scan = t_close = 0.0
        for _ in range(e_steps):
            ta = time.perf_counter()
            p2 = ldw.MIPlan(snp_pin, hdw, sy.paint, blk, device=local_rank)
            tb = time.perf_counter()
            sr, lr, bd, thr, prob, st2 = p2.scan(sy.g, SR_DIST, LR_RETAIN, lra, 0, world, rank, copy=False)
            tc = time.perf_counter()
            d2h += (int(sr.n) + int(lr.n)) * 32
            pe += st2["n_pairs"]
            p2.close()
            td = time.perf_counter()
            t_plan += tb - ta; t_scan += tc - tb; t_close += td - tc
        barrier()
        te = allmax(time.perf_counter() - t0)
        # what follows the scan in perform_MI_computation (R/computePairwiseMI.R:118-143), native host code, on the
        # short-range table the last step left in pinned host memory: reported beside the metric, not part of it
        post = None
        if rank == 0 and world == 1 and not args.no_post:
            try:
                tp0 = time.perf_counter()
                sp = api.mergeNsort_sr_links(ldw.CdsVar(sy.paint, 3), sr, SR_DIST, None, 3.0)
                tp1 = time.perf_counter()
                ar = api.runARACNE({k: sp.df[k][sp.red] for k in ("pos1", "pos2", "MI")}, {k: sp.df[k][sp.chk] for k in ("pos1", "pos2", "MI")})
                tp2 = time.perf_counter()
                post = {"mergeNsort_sr_links_s": tp1 - tp0, "runARACNE_s": tp2 - tp1, "n_sr_links": int(sr.n), "n_df": int(len(sp.df["row"])),
                        "n_red": int(len(sp.red)), "n_aracne_check": int(len(sp.chk)), "aracne_kept": int(ar.sum()),
                        "beta_shapes": [f["shape"].tolist() for f in sp.fits], "nm_evals": [f["nm_evals"] for f in sp.fits],
                        "host_threads": os.cpu_count()}
            except Exception as ex:  # noqa: BLE001 -- a failed fit must not take the metric down with it
                post = {"error": str(ex)}
        pe_all = allsum(float(pe))
        h2d = codes_pin.nbytes + hdw.nbytes + sy.POS.nbytes + sy.paint.nbytes
        e2e = {"value": pe_all / te, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h // e_steps),
               "ms_per_step": 1e3 * te / e_steps, "steps": e_steps,
               "breakdown_ms": {"plan_create": 1e3 * t_plan / e_steps, "scan_call": 1e3 * t_scan / e_steps,
                                "plan_destroy": 1e3 * t_close / e_steps, "scan_device": st2["t_scan_ms"] + st2["t_select_ms"],
                                "scan_kernels": st2["t_kernel_ms"], "select_tail": st2["t_select_ms"], "host_prep": st2["t_host_prep_ms"], "d2h_tail": st2["t_d2h_ms"]},
               "includes": "host->device upload, operand packing, scan, link materialisation, device->host copy of all link columns"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ roofline of the dominant kernel (mi_scan_kernel)
    peaks = load_peaks()
    peak_tf = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
    n_launch = max(1, agg["n_scan_launches"])
    avg_launch_s = kern_ms * 1e-3 / n_launch
    alg_flop_per_launch = 50.0 * S * (agg["n_pairs"] / n_launch)  # 50*S flop per SNP pair
    achieved_tf = alg_flop_per_launch / avg_launch_s / 1e12
    exec_tops = agg["exec_int8_ops"] / (kern_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "mi_scan_traffic.json")
    if os.path.exists(tpath) and args.config == "C2" and not args.nsnp:
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]  # from one ncu --set full capture
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                "kernel": "mi_scan_kernel", "avg_launch_ms": 1e3 * avg_launch_s, "launches": n_launch,
                "kernel_share_of_step": kern_ms / dev_ms if dev_ms else None,
                "executed_int8_tops": exec_tops,
                "executed_frac_of_int8_peak": exec_tops / (2.0 * (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"])),
                "note": "achieved = ALGORITHMIC 50*S flop/pair over the kernel's CUDA-event time; the kernel executes fewer "
                        "tensor ops than that (only r-1 allele planes per site enter the GEMM, 4 int8 K-passes), see DESIGN.md"}

    cpu = None
    if not args.no_cpu and world == 1:
        pr, dt, thr = cpu_reference_sample(sy, hdw, args.cpu_sample)
        cpu = {"value": pr / dt, "unit": "pairs/s", "cores": thr, "kind": "port",
               "sample": f"one off-diagonal {min(args.cpu_sample, n // 2)}x{min(args.cpu_sample, n // 2)} block of the same workload "
                         f"({pr} pairs, {dt:.1f} s), reference-shaped C/OpenMP port of R/computePairwiseMI.R + src/computeMI.cpp"}

    line = {"metric": "weighted SNP-pair MI/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int8 (tcgen05 kind::i8 with int32 accumulation of 28-bit fixed-point weights; fp32 MI epilogue, fp64 refinement of long-range links)",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_all),
            "roofline": roofline, "cpu_baseline": cpu,
            "detail": {"wall_ms_per_step": 1e3 * wall_max / args.steps, "post_scan_host": post if not args.no_e2e else None,
                       "step_ms": {"min": min(step_ms), "median": sorted(step_ms)[len(step_ms) // 2], "max": max(step_ms)}, "hdw_s": t_hdw, "pack_ms": stats["t_pack_ms"], "host_prep_ms": stats["t_host_prep_ms"],
                       "pairs_per_step": pairs_all / args.steps, "n_sr": stats["n_sr"], "n_lr_kept": stats["n_lr_kept"],
                       "n_reruns": stats["n_reruns"], "n_candidates": stats["n_candidates"], "tiles_per_step_rank0": agg["n_tiles"] / args.steps,
                       "lr_links_approx": lra}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
