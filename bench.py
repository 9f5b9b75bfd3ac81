#!/usr/bin/env python
"""bench.py -- B-SHOT front-end benchmark (contract: see the task prompt / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--gpus N] ...               # CPU oracle port, host cores
  torchrun ... bench.py --gpus N ...                            # one rank per GPU (N > 1)

Workload (config.workload = "C2"): BASELINE.json configs[1] -- frame-to-frame B-SHOT odometry front
end over a synthetic HDL-32E sequence (69 440 rays/frame, mm): voxel build -> seg-ratio detector ->
top-K -> normals -> SHOT LRF + 352-bin histogram -> B-SHOT bits -> Hamming mutual-NN match against
the previous frame.  One step = one frame.  metric = B-SHOT descriptors/s through that whole path.
 * value : cloud already resident in HBM when the timed region starts (bshot_process_frame_dev).
 * e2e   : the C-ABI call a reference maintainer would make (bshot_process_frame) on pinned HOST
           buffers, H2D + D2H inside the timed region.
 * N > 1 : frame extraction does not shard (replicas only) -> every rank processes a replica of
           the frame stream, no data-path collective, weak scaling.  The part of the path that
           DOES shard -- frame-to-map Hamming search against a map split over the ranks, per-rank
           top-2 candidates exchanged by stores into symmetric peer memory + flag barriers (or one
           NCCL all-gather + all-reduce, BSHOT_EXCHANGE=nccl) -- is timed in the same run and
           reported under "map_match" (C4: Q = 10 000 queries vs T = 1 048 576 map descriptors).
 * extra objects at N = 1: "c3" (BASELINE.json configs[2]: HDL-64E 120 k-point frame, K = 10 000, and the
           SHOT-radius sweep with FULL normals), "exact_mode" (the same C2 frames with BSHOT_EXACT_SUMS=1).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_bshot, load_oracle, load_sharded, load_synth  # noqa: E402  (loaders only, no pytest needed)

HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
N_FRAMES = 12               # distinct synthetic frames per rank, cycled
TOP_K = 2048                # C1/C2 "~2k keypoints" (reference default 600: --top-k 600)
L2_FLUSH_BYTES = 256 << 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_front_end(oracle, scans, top_k, threads, detector_threads=None):
    """the oracle port of the reference path on host cores; returns seconds per frame (mean)"""
    prev = None
    t0 = time.perf_counter()
    for xyz in scans:
        c = oracle.Cloud(xyz)
        ratio = c.seg_ratio(3000.0, 300, oracle.SR_CV, threads=detector_threads or threads)
        idx, _ = oracle.select_keypoints(ratio, top_k, oracle.TIE_STDSORT)
        d = c.compute_descriptors(xyz[idx], 3000.0, 300, oracle.MODE_REFERENCE, threads=threads)
        tgt = d["bits"] if prev is None else prev
        m = oracle.match(d["bits"], tgt, want_right=True, threads=detector_threads or threads)
        oracle.mutual(m["left_idx"], m["right_idx"])
        prev = d["bits"]
    return (time.perf_counter() - t0) / len(scans)


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port; the PCL reference cannot be built
    here) on all host cores, same workload / metric."""
    if rank != 0:
        return
    oracle, synth = load_oracle(), load_synth()
    cores = os.cpu_count() or 1
    scans = [synth.make_scan("hdl32e", f) for f in range(min(3, max(1, args.steps)))]
    for _ in range(min(args.warmup, 1)):
        cpu_front_end(oracle, scans[:1], args.top_k, cores)
    steps = max(1, min(args.steps, 6))
    per = []
    for s in range(steps):
        per.append(cpu_front_end(oracle, [scans[s % len(scans)]], args.top_k, cores))
    sec = float(np.mean(per))
    val = args.top_k / sec
    line = {
        "impl": "reference", "metric": "bshot_frontend_descriptors_per_s", "value": val, "unit": "descriptors/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64+u32", "data": "synthetic",
        "config": {"workload": "C2", "sensor": "hdl32e", "points_per_frame": int(np.mean([len(s) for s in scans])),
                   "top_k": args.top_k, "radius_mm": 3000, "normals": "REFERENCE"},
        "cpu_baseline": {"value": val, "unit": "descriptors/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} full frames (detector+normals+SHOT+B-SHOT+match), OpenMP over {cores} threads"},
        "e2e": {"value": val, "unit": "descriptors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--top-k", dest="top_k", type=int, default=TOP_K)
    ap.add_argument("--sensor", default="hdl32e")
    ap.add_argument("--normals", default="reference", choices=["reference", "full"])
    ap.add_argument("--map-q", type=int, default=10000)
    ap.add_argument("--map-t", type=int, default=1 << 20)
    ap.add_argument("--map-steps", type=int, default=10)
    ap.add_argument("--no-map", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the HDL-64E frame / SHOT-radius sweep (C3) object")
    ap.add_argument("--map-only", action="store_true", help="debug: print only the map_match object")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    bs, synth = load_bshot(), load_synth()
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    hbm_peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs ---------------------------------------------------------------------------------
    # replicas: every rank runs the SAME frames (identical work per GPU keeps the weak-scaling figure clean)
    frames = [synth.make_scan(args.sensor, f) for f in range(N_FRAMES)]
    npts = [len(f) for f in frames]
    max_n = max(npts)
    mode = bs.NORMALS_REFERENCE if args.normals == "reference" else bs.NORMALS_FULL
    params = bs.default_params(top_k=args.top_k, normals_mode=mode)
    ctx = bs.Context(local_rank, max_points=max_n + 1024, max_keypoints=max(args.top_k, args.map_q),
                     max_targets=max(args.top_k, (args.map_t + world - 1) // world))
    st = torch.cuda.ExternalStream(ctx.stream)
    d_frames = [torch.from_numpy(f).cuda() for f in frames]                 # resident in HBM
    h_frames = [torch.from_numpy(f).pin_memory() for f in frames]           # pinned host copies
    h_kp = torch.empty(args.top_k, dtype=torch.int32).pin_memory()
    h_bits = torch.empty((args.top_k, 6), dtype=torch.int64).pin_memory()
    h_pairs = torch.empty((args.top_k, 2), dtype=torch.int32).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device="cuda")

    def l2_flush():
        with torch.cuda.stream(st):
            flush.fill_(1.0)

    def step_resident(i):
        f = i % N_FRAMES
        ctx.process_frame_dev(d_frames[f].data_ptr(), npts[f], 12, params)

    def step_e2e(i):
        f = i % N_FRAMES
        return ctx.process_frame_raw(h_frames[f].data_ptr(), npts[f], 12, params, h_kp.data_ptr(),
                                     h_bits.data_ptr(), h_pairs.data_ptr())

    # ---- HBM-resident timing (value) ---------------------------------------------------------------
    ctx.enable_timing(True)
    for i in range(W):
        step_resident(i)
    ctx.sync()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    stage_acc, counters_acc = {}, {}
    sampler = ClockSampler(local_rank) if rank == 0 else None   # runs across the value and e2e timed regions
    launches0 = ctx.launch_count()
    barrier()
    for i in range(K):
        l2_flush()
        ev[i][0].record(st)
        step_resident(W + i)
        ev[i][1].record(st)
        for k, v in ctx.stage_times().items():      # synchronises; outside the event pair
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        for k, v in ctx.frame_counters().items():
            counters_acc[k] = counters_acc.get(k, 0) + v
    barrier()
    launches = ctx.launch_count() - launches0
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / K
    n_desc = counters_acc["keypoints"] / K                       # descriptors actually produced per frame
    value = world * n_desc / (ms_step * 1e-3)
    stages = {k: v / K for k, v in stage_acc.items()}

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    kern = {"seg_ratio": ("seg_ratio_kernel", 16.0 * counters_acc["detector_neighbours"] / K + 12.0 * np.mean(npts)),
            "shot_bshot": ("shot_kernel", 16.0 * counters_acc["shot_neighbours"] / K + 48.0 * n_desc),
            "normals": ("normals_kernel", 16.0 * counters_acc["normals_neighbours"] / K + 16.0 * n_desc),
            "voxel_build": ("grid_build (9 kernels)", 2 * 16.0 * np.mean(npts))}
    dom = max(kern, key=lambda k: stages[k])
    alg_bytes = kern[dom][1]
    achieved = alg_bytes / (stages[dom] * 1e-3) / 1e9
    traffic, ncu_extra = None, None
    try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))
        traffic = tj.get(kern[dom][0])
        ncu_extra = tj.get("_ncu", {}).get(kern[dom][0])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kern[dom][0], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": stages[dom], "ncu": ncu_extra,
                "note": "cloud (<2 MB) is L2 resident: algorithmic bytes are re-read from L2/L1, not HBM; the kernel is "
                        "bound by dependent latency / instruction issue (see ncu.issue_active_pct), DESIGN.md section 7"}

    # ---- end-to-end through the host-buffer C ABI (e2e) ---------------------------------------------
    ctx.enable_timing(False)
    for i in range(W):
        step_e2e(i)
    barrier()
    t_e2e = 0.0
    for i in range(K):
        l2_flush()
        ctx.sync()
        t0 = time.perf_counter()
        nk, _ = step_e2e(W + i)
        t_e2e += time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    t_e2e = max_over_ranks(t_e2e)
    e2e_val = world * n_desc / (t_e2e / K)
    e2e = {"value": e2e_val, "unit": "descriptors/s", "ms_per_step": t_e2e / K * 1e3,
           "h2d_bytes_per_step": int(np.mean(npts) * 12),
           "d2h_bytes_per_step": int(args.top_k * (4 + 48 + 12) + 8),
           "api": "bshot_process_frame (C ABI, pinned host buffers, synchronous)"}

    # ---- sharded frame-to-map matching (the part of the path that shards) -----------------------------
    map_match = None
    if not args.no_map:
        T, Q = args.map_t, args.map_q
        per = (T + world - 1) // world
        lo, hi = rank * per, min(T, (rank + 1) * per)
        tfull = synth.random_descriptors(T, seed=7)                # same global map on every rank, own shard kept
        ctx.map_reset()
        ctx.map_append(tfull[lo:hi])
        q = synth.random_descriptors(Q, seed=8)
        planted = np.random.default_rng(9).permutation(T)[:64]     # self-check: 64 queries are exact copies of map entries
        q[:64] = tfull[planted]
        del tfull
        dq = torch.from_numpy(q.view(np.int64)).cuda()
        # the per-call protocol (shard search, record exchange, merge, sharded reverse pass) lives in the package:
        # sharded.DeviceShardedMatcher; BSHOT_EXCHANGE=nccl forces the NCCL exchange, default = peer memory if possible
        sharded = load_sharded()
        matcher = sharded.DeviceShardedMatcher(ctx, world, rank, Q, mode="nccl" if os.environ.get("BSHOT_EXCHANGE") == "nccl" else "auto",
                                               device=torch.device("cuda", local_rank))

        def map_calls(n):
            out = None
            for _ in range(n):
                out = matcher.match(dq.data_ptr(), lo)
            return out

        last = map_calls(3)
        barrier()
        rec = last[:64].cpu().numpy().view(bs.CAND_DTYPE).reshape(64)
        chk = bs.unpack_cands(rec)
        if not (np.array_equal(chk["idx1"], planted) and (chk["dist1"] == 0).all() and np.array_equal(chk["rq"], np.arange(64))):
            raise SystemExit("bench.py: sharded map match self-check failed (planted duplicates not recovered)")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(st)
        map_calls(args.map_steps)
        e1.record(st)
        barrier()
        mm_ms = max_over_ranks(e0.elapsed_time(e1)) / args.map_steps
        matcher.check()
        popc_peak = ctx.popc_peak()
        pairs = float(Q) * float(T)
        map_match = {"workload": "C4", "Q": Q, "T": T, "shards": world, "ms_per_call": mm_ms,
                     "pairs_per_s": pairs / (mm_ms * 1e-3), "target_GBps": T * 48 / (mm_ms * 1e-3) / 1e9,
                     "roofline": {"bound": "popc", "achieved": 11 * pairs / (mm_ms * 1e-3) / 1e12,
                                  "peak": world * popc_peak / 1e12, "unit": "TPOPC32/s",
                                  "frac": 11 * pairs / (mm_ms * 1e-3) / (world * popc_peak),
                                  "peak_source": "measured live (bshot_popc_peak microbenchmark) x shards"},
                     "collective": matcher.describe(),
                     "gpu_launches_per_call": (ctx.launch_count() - l0) // args.map_steps}


    # ---- C3: HDL-64E-shaped 120 k-point scans, K = 10 000 (rank 0, N = 1 only) --------------------------
    # (a) the north_star frame: full extraction + match of one 120 k-point scan, REFERENCE normals;
    # (b) extraction throughput (normals + LRF + SHOT352 + B-SHOT) over the SHOT radius, FULL normals.
    c3 = None
    if rank == 0 and world == 1 and not args.no_c3 and not args.map_only:
        K64, NF64 = 10000, 3
        f64 = [synth.make_scan("hdl64e", f) for f in range(NF64)]
        n64 = [len(f) for f in f64]
        d64 = [torch.from_numpy(f).cuda() for f in f64]
        ctx64 = bs.Context(local_rank, max_points=max(n64) + 1024, max_keypoints=K64, max_targets=K64)
        st64 = torch.cuda.ExternalStream(ctx64.stream)
        ctx64.enable_timing(True)

        def run64(params64, reps):
            acc, nbr = {}, 0
            for i in range(2):                                        # warm-up (also fills prev-frame descriptors)
                ctx64.process_frame_dev(d64[i % NF64].data_ptr(), n64[i % NF64], 12, params64)
            ctx64.sync()
            for i in range(reps):
                with torch.cuda.stream(st64):
                    flush.fill_(1.0)
                ctx64.process_frame_dev(d64[i % NF64].data_ptr(), n64[i % NF64], 12, params64)
                for k, v in ctx64.stage_times().items():
                    acc[k] = acc.get(k, 0.0) + v / reps
                nbr += ctx64.frame_counters()["shot_neighbours"] / reps
            return acc, nbr

        st_ref, _ = run64(bs.default_params(top_k=K64), 12)
        sweep = []
        for R in (500.0, 1000.0, 2000.0, 3000.0, 4000.0):
            ctx64.reset()
            st_r, nbr = run64(bs.default_params(top_k=K64, normals_mode=bs.NORMALS_FULL, normal_radius=R, shot_radius=R), 6)
            ext_ms = st_r["normals"] + st_r["shot_bshot"]
            sweep.append({"radius_mm": R, "normals_ms": st_r["normals"], "shot_bshot_ms": st_r["shot_bshot"],
                          "descriptors_per_s": K64 / (ext_ms * 1e-3), "shot_neighbours_per_keypoint": nbr / K64,
                          "shot_algorithmic_GBps": 32.0 * nbr / (st_r["shot_bshot"] * 1e-3) / 1e9})
        c3 = {"workload": "C3", "sensor": "hdl64e", "points_per_frame": int(np.mean(n64)), "top_k": K64,
              "frame_reference_normals": {"ms_per_frame": st_ref["frame"], "stages_ms": st_ref,
                                          "north_star_target_ms": 2.0},
              "radius_sweep_full_normals": sweep,
              "note": "stage times from CUDA events on the context stream, L2 flushed before every frame"}
        ctx64.close()

    # ---- exact-sums mode (reference summation order, bit-identical seg-ratios / keypoints): same C2 frames ----
    exact = None
    if rank == 0 and world == 1 and not args.map_only:
        os.environ["BSHOT_EXACT_SUMS"] = "1"
        try:
            ctxe = bs.Context(local_rank, max_points=max_n + 1024, max_keypoints=args.top_k, max_targets=args.top_k)
        finally:
            os.environ.pop("BSHOT_EXACT_SUMS", None)
        ctxe.enable_timing(True)
        acc = {}
        for i in range(3 + 10):
            ctxe.process_frame_dev(d_frames[i % N_FRAMES].data_ptr(), npts[i % N_FRAMES], 12, params)
            if i >= 3:
                for k, v in ctxe.stage_times().items():
                    acc[k] = acc.get(k, 0.0) + v / 10
        exact = {"env": "BSHOT_EXACT_SUMS=1", "ms_per_frame": acc["frame"], "stages_ms": acc,
                 "note": "fp32 running sums replayed in neighbour order: seg-ratios and keypoints bit-identical to the oracle"}
        ctxe.close()

    # ---- CPU baseline (rank 0, N = 1 only; bounded sample) --------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        oracle = load_oracle()
        cores = os.cpu_count() or 1
        sample = frames[:3]
        cpu_front_end(oracle, sample[:1], args.top_k, cores)      # warm-up (page-in, OpenMP pool)
        sec = cpu_front_end(oracle, sample, args.top_k, cores)
        sec_ref = cpu_front_end(oracle, sample[:2], args.top_k, min(12, cores), detector_threads=1)
        cpu_baseline = {"value": args.top_k / sec, "unit": "descriptors/s", "cores": cores, "kind": "port",
                        "ms_per_frame": sec * 1e3,
                        "sample": "3 full frames of this workload (detector+normals+SHOT+B-SHOT+match), all cores",
                        "reference_threading": {"value": args.top_k / sec_ref, "ms_per_frame": sec_ref * 1e3,
                                                "detector_match_threads": 1, "normals_shot_threads": min(12, cores),
                                                "sample": "2 frames"}}

    if rank == 0 and args.map_only:
        print(json.dumps(map_match))
    elif rank == 0:
        line = {
            "metric": "bshot_frontend_descriptors_per_s", "value": value, "unit": "descriptors/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32/f64+u32", "data": "synthetic",
            "config": {"workload": "C2", "sensor": args.sensor, "points_per_frame": int(np.mean(npts)),
                       "rays_per_frame": 69440 if args.sensor == "hdl32e" else 120000, "top_k": args.top_k,
                       "descriptors_per_frame": n_desc, "radius_mm": 3000, "max_nn": 300,
                       "normals": args.normals.upper(), "frames_cycled": N_FRAMES,
                       "parallelism": "replicas (one frame stream per GPU)" if world > 1 else "single GPU",
                       "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB write), per-step CUDA events on the context stream"},
            "stages_ms": stages, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "map_match": map_match, "c3": c3, "exact_mode": exact,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
