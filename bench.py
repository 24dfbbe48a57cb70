#!/usr/bin/env python
"""bench.py -- MSCKF updates/sec on the B200 engine (and the reference-arm / CPU baseline beside it).

A "step" is one pass of the hot path over one batch of synthetic input: ONE marginalize() call
(msckf.h:336-449 -> measurementUpdate :1325-1423) on N_feat tracks x N_clones observations.
Workload = BASELINE.json configs[1]: synthetic 300 features x 30 camera clones, float32, produced through
the public MSCKF<_S> surface so that the timed update processes exactly 300 x 30.

  value        device-timed (CUDA events on the engine's stream) kernels of one update, inputs resident in HBM
  e2e          the same update through the drop-in class's marginalize() with HOST buffers: packing, H2D of the
               track batch, all kernels, D2H of the per-track report and of the corrected state (getImuState)
  roofline     dominant kernel: SURVEY 8d algorithmic bytes of one update / that kernel's mean device time
  cpu_baseline the oracle (CPU restatement of the reference's Eigen path) timed on this box's host, 1 core

Every rank runs the same per-GPU work (weak scaling; independent filters shard with no collective on the data
path, NCCL is only used for the barrier / max-over-ranks).  `--impl reference` times the CPU oracle instead.
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_FEAT, N_CLONES = 300, 30
DTYPE = np.float32


def algorithmic_bytes(nf, nc, b):
    """2*b*m*6L (H_o non-zero columns written + read) + 2*b*m (r_o) + 3*b*n^2 (P) + inputs."""
    L = M = nc
    m = nf * (2 * L - 3)
    n = 15 + 6 * M
    return 2 * b * m * 6 * L + 2 * b * m + 3 * b * n * n + nf * L * 2 * b + M * 7 * b


def algorithmic_flops(nf, nc):
    """SURVEY.md 8d reference-path flops (projection + gating + QR + R_n + n^3 tail)."""
    L = M = nc
    rho, n, c = 2 * L - 3, 15 + 6 * M, 6 * M
    m = nf * rho
    return (nf * 2 * rho * (2 * L) * (6 * L) + nf * (2 * rho * (6 * L) ** 2 + 2 * rho ** 2 * (6 * L) + rho ** 3 / 3)
            + (2 * (m - 15) * c ** 2 - 2 / 3 * c ** 3) + 2 * (nf * L) * n ** 2 + 12 * n ** 3)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.01)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def ready_filter(dtype, seq, device, max_clones=40):
    """a filter driven through the public API up to (not including) the marginalize() that processes N x M."""
    from msckf_mono_b200 import engine_filter, synth
    wl = synth.make_window_workload(n_features=N_FEAT, n_clones=N_CLONES, seq=seq)
    f = engine_filter(dtype, device=device, max_clones=max_clones, max_tracks=512, max_obs=512 * N_CLONES)
    synth.drive(f, wl, marginalize_last=False)
    return f


def cpu_oracle_updates(seconds_target, threads, dtype):
    """time the CPU oracle's marginalize() on the same workload; returns (updates/s, n_updates, wall seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from msckf_mono_b200 import synth
    from msckf_mono_b200.cview import CFilter
    lib = ROOT / "oracle" / "libmsckf_oracle.so"
    if not lib.exists():
        raise RuntimeError("oracle/libmsckf_oracle.so missing (run __graft_entry__.build())")

    def prepare(seq):
        o = CFilter(lib, "msckf_oracle_", dtype)
        synth.drive(o, synth.make_window_workload(n_features=N_FEAT, n_clones=N_CLONES, seq=seq), marginalize_last=False)
        return o

    def one(o):
        t0 = time.perf_counter()
        o.marginalize()  # ctypes releases the GIL: threads run truly in parallel
        return time.perf_counter() - t0

    warm = prepare(0)
    t_one = one(warm)
    rounds = max(1, int(round(seconds_target / max(t_one, 1e-3))))
    done, wall = 0, 0.0
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for r in range(rounds):
            filts = [prepare(1 + r * threads + i) for i in range(threads)]  # untimed: builds the pre-update state
            t0 = time.perf_counter()
            list(ex.map(one, filts))
            wall += time.perf_counter() - t0
            done += threads
    return done / wall, done, wall


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path (here: its restatement, the reference cannot be built on
    this image) on the host cores.  The filter is single-threaded like the reference; independent filters use
    all cores (one update per core per step)."""
    if rank != 0:
        return
    # the oracle's working set (three 17100 x 195 matrices per filter) makes it memory-bandwidth bound: throughput
    # saturates around 16 concurrent filters (measured on the 128-core GPU box: 8.1/s with 128 threads vs 2.0/s with 1),
    # and a step must stay bounded, so a step = one update on each of min(cores, 16) threads
    cores = min(os.cpu_count() or 1, 16)
    from concurrent.futures import ThreadPoolExecutor
    from msckf_mono_b200 import synth
    from msckf_mono_b200.cview import CFilter
    lib = ROOT / "oracle" / "libmsckf_oracle.so"

    def prepare(seq):
        o = CFilter(lib, "msckf_oracle_", DTYPE)
        synth.drive(o, synth.make_window_workload(n_features=N_FEAT, n_clones=N_CLONES, seq=seq), marginalize_last=False)
        return o

    def one(o):
        o.marginalize()

    total, wall = 0, 0.0
    with ThreadPoolExecutor(max_workers=cores) as ex:
        for step in range(args.warmup + args.steps):
            filts = [prepare(1000 + step * cores + i) for i in range(cores)]
            t0 = time.perf_counter()
            list(ex.map(one, filts))
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                total += cores
                wall += dt
    val = total / wall
    line = {"impl": "reference", "metric": "msckf_updates_per_sec", "value": val, "unit": "updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{N_FEAT} features x {N_CLONES} clones, float32, one marginalize() per update; "
                                   f"a step = {cores} independent updates, one per host core", "n_features": N_FEAT, "n_clones": N_CLONES},
            "cpu_baseline": {"value": val, "unit": "updates/s", "cores": cores, "kind": "port",
                             "sample": f"{total} updates of the {N_FEAT}x{N_CLONES} fp32 workload, {cores} concurrent single-threaded oracle filters"},
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference (Eigen 3 + Boost + ROS) is not buildable on this image; this is oracle/ (its line-by-line CPU restatement, thin-Q form)"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=8, help="filters pipelined per GPU in the batched side measurement")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--skip-e2e", action="store_true", help="profiling aid: only the device-timed steps (use under ncu)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from msckf_mono_b200 import capi, shard

    K, W = args.steps, args.warmup
    # ---------------------------------------------------------------- setup (untimed)
    tmpl = ready_filter(DTYPE, seq=100 + rank, device=local_rank)
    off, obs, idx = tmpl.packQueued()
    batch = capi.TrackBatch(off, obs, idx, DTYPE)
    tmpl_eng = capi.Engine(DTYPE, borrowed=tmpl.engineHandle())
    work = capi.Engine(DTYPE, device=local_rank, max_clones=40, max_tracks=512, max_obs=512 * N_CLONES)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def flush_l2():
        flush.add_(1.0)
        torch.cuda.synchronize()

    def device_step():
        work.copy_state_from(tmpl_eng)
        work.stage(capi.MARGINALIZE, batch)
        work.synchronize()
        flush_l2()
        ms = work.launch_timed()
        rep = work.fetch(batch.n_tracks)
        return ms, rep

    def barrier():
        shard.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-timed steps
    for _ in range(W):
        device_step()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    barrier()
    l0 = work.launch_count()
    t_wall0 = time.perf_counter()
    times = []
    for _ in range(K):
        ms, rep = device_step()
        times.append(ms)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches_timed = work.launch_count() - l0
    assert rep["m"] == N_FEAT * (2 * N_CLONES - 3), rep
    dev_ms = float(np.sum(times))

    if args.skip_e2e:
        sampler.stop_flag = True
        if rank == 0:
            print(json.dumps({"profiling_only": True, "ms_per_step": dev_ms / K, "gpu_launches": int(launches_timed)}))
        return
    # ---------------------------------------------------------------- end-to-end through the class surface
    filts = [ready_filter(DTYPE, seq=200 + rank * 1000 + i, device=local_rank) for i in range(K + W)]
    e2e_times, e2e_parts = [], []
    for i, f in enumerate(filts):
        flush_l2()
        t0 = time.perf_counter()
        f.marginalizeLaunch()      # marginalize() = launch (host pack + one H2D copy + graph launch) ...
        t1 = time.perf_counter()
        f.marginalizeCollect()     # ... + collect (wait for the kernels, one D2H report copy, host bookkeeping)
        t2 = time.perf_counter()
        st = f.getImuState()       # D2H of the corrected state (the step's result)
        dt = time.perf_counter() - t0
        if i >= W:
            e2e_times.append(dt)
            e2e_parts.append((t1 - t0, t2 - t1, t0 + dt - t2))
    assert np.isfinite(st["p_I_G"]).all()
    e2e_s = float(np.sum(e2e_times))
    if rank == 0:
        pl, pc, ps = (1e6 * float(np.mean([p[j] for p in e2e_parts])) for j in range(3))
        print(f"e2e breakdown (us per update): launch {pl:.0f} (pack + H2D + graph launch), collect {pc:.0f} (kernels + D2H + bookkeeping), "
              f"getImuState {ps:.0f}", file=sys.stderr)
    up16 = lambda x: (x + 15) & ~15  # the engine's packed report block: (m, rank) | 5 int flags per track | p_f_G | gamma, one D2H copy
    rep_bytes = 16 + 5 * up16(4 * N_FEAT) + up16(4 * 3 * N_FEAT) + up16(4 * N_FEAT)
    state_bytes = 1192 + 8 * 4 * N_CLONES  # sizeof(DevState<float>) + clone poses
    # ---------------------------------------------------------------- batched side measurement (multi-stream pipelining)
    B = args.batch
    bfilts = [ready_filter(DTYPE, seq=5000 + rank * 1000 + i, device=local_rank) for i in range(B * 4)]
    batched_s = []
    for r in range(4):
        grp = bfilts[r * B:(r + 1) * B]
        flush_l2()
        t0 = time.perf_counter()
        for f in grp:
            f.marginalizeLaunch()
        for f in grp:
            f.marginalizeCollect()
        dt = time.perf_counter() - t0
        if r >= 1:
            batched_s.append(dt)
    # 32 filters in flight (2 rounds, the first warms up)
    B2 = 32
    b2filts = [ready_filter(DTYPE, seq=9000 + rank * 1000 + i, device=local_rank) for i in range(B2 * 2)]
    batched2_s = []
    for r in range(2):
        grp = b2filts[r * B2:(r + 1) * B2]
        flush_l2()
        t0 = time.perf_counter()
        for f in grp:
            f.marginalizeLaunch()
        for f in grp:
            f.marginalizeCollect()
        if r >= 1:
            batched2_s.append(time.perf_counter() - t0)
    # the same 32 filters' worth of work through the batched C entry point (host work on several threads)
    from msckf_mono_b200.cview import marginalize_batch
    host_threads = max(1, min(8, (os.cpu_count() or 1)))
    b3filts = [ready_filter(DTYPE, seq=13000 + rank * 1000 + i, device=local_rank) for i in range(B2 * 2)]
    batched3_s = []
    for r in range(2):
        grp = b3filts[r * B2:(r + 1) * B2]
        flush_l2()
        t0 = time.perf_counter()
        marginalize_batch(grp, threads=host_threads)
        if r >= 1:
            batched3_s.append(time.perf_counter() - t0)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # ---------------------------------------------------------------- per-kernel profile (CUDA events between kernels)
    work.set_option(1, 1.0)
    per = {}
    for _ in range(6):
        device_step()
        for name, ms in work.kernel_times():
            per.setdefault(name, []).append(ms)
    if os.environ.get("MSCKF_TAIL_PROFILE"):
        import ctypes as C
        buf = (C.c_ulonglong * 80)()
        capi.lib().msckf_b200_tail_profile(work.h, buf, 80)
        st = [int(x) for x in list(buf)[:40] if x]
        print("tail stamps (us since start):", [round((x - st[0]) / 1e3, 1) for x in st], file=sys.stderr)
        sj = []
        for x in list(buf)[40:64]:
            if not x:
                break
            sj.append(int(x))
        dd = [int(x) for x in list(buf)[64:80]]
        if dd[0]:
            print("tail diag-block fine stamps, A factor, pivots 20 and 21 (ns since pivot start: shuffles + FMA, decide, next bracket + rsqrt, stores, publish):",
                  [[dd[6 * q + u] - dd[6 * q] for u in range(1, 6)] for q in range(2)], "pivot-to-pivot:", dd[6] - dd[0],
                  "| block 0: loaded -> workers done (us):", round((dd[13] - dd[12]) / 1e3, 2), file=sys.stderr)
        print("jac stamps (us since start):", [round((x - sj[0]) / 1e3, 1) for x in sj], file=sys.stderr)
    work.set_option(1, 0.0)
    kern_ms = {k: float(np.mean(v[1:])) for k, v in per.items()}
    dom = max(kern_ms, key=kern_ms.get)

    # ---------------------------------------------------------------- reduce over ranks
    dev_ms, e2e_s, bsum, b2sum, b3sum = shard.max_over_ranks(
        [dev_ms, e2e_s, float(np.sum(batched_s)), float(np.sum(batched2_s)), float(np.sum(batched3_s))], device="cuda")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "of measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "of fallback (6.65 TB/s)"
    bytes_alg = algorithmic_bytes(N_FEAT, N_CLONES, 4)
    achieved = bytes_alg / (kern_ms[dom] * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(dom)
    except Exception:
        pass
    cpu_val, cpu_n, cpu_wall = cpu_oracle_updates(args.cpu_seconds, 1, DTYPE)

    value = world * K / (dev_ms * 1e-3)
    line = {
        "metric": "msckf_updates_per_sec", "value": value, "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{N_FEAT} features x {N_CLONES} camera clones, float32, single update() (= one marginalize(): "
                               "triangulation + Jacobian/null-space + gating + compression + Kalman/covariance update) per step, "
                               "one filter per GPU, state restored and L2 flushed (256 MiB write) between timed steps",
                   "n_features": N_FEAT, "n_clones": N_CLONES, "stacked_rows_m": N_FEAT * (2 * N_CLONES - 3), "state_dim_n": 15 + 6 * N_CLONES,
                   "timing": "CUDA events on the engine stream around the update's kernels, summed over steps, max over ranks",
                   "l2": "flushed between timed iterations", "parallelism": f"independent filters, {world} GPU(s), no data-path collective"},
        "clocks": sampler.result(),
        "e2e": {"value": world * K / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": batch.h2d_bytes(),
                "d2h_bytes_per_step": rep_bytes + state_bytes, "ms_per_step": 1e3 * e2e_s / K,
                "path": "msckf_mono::MSCKF<float>::marginalize() + getImuState() via the C view, host wall clock"},
        "gpu_launches": int(launches_timed),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                     "kernel": dom, "kernel_ms": kern_ms[dom], "peak_source": peak_src, "algorithmic_bytes_per_update": bytes_alg,
                     "whole_update_gbs": bytes_alg / (dev_ms / K * 1e-3) / 1e9,
                     "reference_path_gflop_per_update": algorithmic_flops(N_FEAT, N_CLONES) / 1e9,
                     "note": "the update is latency-bound (11 short kernels, 10 on the critical path); see DESIGN.md for the per-kernel table",
                     "kernel_ms_all": kern_ms},
        "cpu_baseline": {"value": cpu_val, "unit": "updates/s", "cores": 1, "kind": "port",
                         "sample": f"{cpu_n} marginalize() calls of the same {N_FEAT}x{N_CLONES} fp32 workload in {cpu_wall:.1f} s, "
                                   "oracle/ (CPU restatement of the reference's Eigen path, thin-Q form), single thread like the reference"},
        "batched": {"filters_per_gpu_in_flight": B, "value": world * B * len(batched_s) / bsum, "unit": "updates/s",
                    "path": "marginalizeLaunch() on all filters, then marginalizeCollect() (one stream per filter), host wall clock incl. copies",
                    "value_32_in_flight": world * B2 * len(batched2_s) / b2sum,
                    "value_32_batch_api": world * B2 * len(batched3_s) / b3sum, "batch_api_host_threads": host_threads,
                    "batch_api": "msckf_mono_marginalize_batch (C view) = msckf_b200_update_batch semantics: launch all, collect all, host work on threads"},
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
