#!/usr/bin/env python
"""Benchmark of the feature-matching hot path (BASELINE.json metric: pairs/sec at 8192 SIFT/img).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W # the CPU matcher on the host cores

A "step" is one pass of the hot path over one batch: every image pair of the sequential window
(BASELINE.json configs[1]: 100 synthetic images x 8192 descriptors, overlap=10, cross_check on -> 855
pairs per GPU).  With N > 1 (torchrun, one rank per GPU) every rank owns 100 consecutive images of a
100*N-image sequence (weak scaling) and fetches the overlap-1 = 9 halo images it needs from the next
rank's HBM over NCCL/NVLink inside the timed step.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "pairs/sec at 8192 SIFT/img"
UNIT = "pairs/s"
IMAGES_PER_GPU = 100
N_DESC = 8192
OVERLAP = 10
INT8_SPEC_TOPS = 4500.0  # B200 dense INT8, NVIDIA datasheet
WORKLOAD = (f"{IMAGES_PER_GPU} images x {N_DESC} descriptors per GPU, sequential overlap={OVERLAP}, cross_check on "
            "(BASELINE.json configs[1])")


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


def _physical_gpu_index(local_index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU matcher (oracle port of COLMAP's MatchSiftFeaturesCPU) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(num_pairs: int, threads: int = 0):
    """Time ``num_pairs`` pairs of the bench workload (adjacent 8192-descriptor images) on the CPU oracle."""
    from oracle import oracle
    from scanner_colmap_b200 import synth
    n_img = max(2, min(num_pairs + 1, OVERLAP))
    imgs = synth.make_images(n_img, N_DESC)
    pairs = [(k % (n_img - 1), k % (n_img - 1) + 1) for k in range(num_pairs)]
    t0 = time.perf_counter()
    res, used = oracle.match_many(imgs, pairs, num_threads=threads)
    dt = time.perf_counter() - t0
    return num_pairs / dt, used, dt, sum(len(r) for r in res)


def run_reference(args):
    rank = _env_int("RANK", 0)
    if rank != 0:
        return 0
    import __graft_entry__ as g
    from oracle import oracle
    oracle.build()
    cores = oracle.num_procs()
    total_steps = args.steps + args.warmup
    sample = max(1, min(cores, int(cores * 16 / max(total_steps, 1))))
    for _ in range(args.warmup):
        cpu_sample(sample)
    t = 0.0
    for _ in range(args.steps):
        _, used, dt, _ = cpu_sample(sample)
        t += dt
    value = sample * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"each step = {sample} of the workload's 855 pairs (bounded CPU sample)",
                   "note": "CPU restatement of COLMAP MatchSiftFeaturesCPU (oracle/sift_match_oracle.c, gcc -O3 "
                           "-march=x86-64-v3, OpenMP over pairs); the Eigen/COLMAP/Scanner binary is unbuildable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"{sample} pairs of 8192x8192 per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
class _DevView:
    """Zero-copy torch view of library-owned device memory."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def int8_gemm_peak(torch, dev):
    """Measured dense int8 GEMM rate of this GPU through cuBLASLt (torch._int_mm), for context."""
    try:
        n = 8192
        a = torch.randint(-4, 4, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-4, 4, (n, n), dtype=torch.int8, device=dev).t().contiguous().t()
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch._int_mm(a, b)
            e.record()
            e.synchronize()
            best = max(best, 2.0 * n ** 3 / (s.elapsed_time(e) * 1e-3) / 1e12)
        return best
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs

    world = _env_int("WORLD_SIZE", 1)
    rank = _env_int("RANK", 0)
    local = _env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: 100 * world images in one sequence; this rank's contiguous anchor window + its halo
    from scanner_colmap_b200 import sharding
    sizes = [N_DESC] * (IMAGES_PER_GPU * world)
    sp = sharding.plan(sizes, OVERLAP, world, rank)
    own_ids = list(range(*sp.own))
    halo_ids = [row for row, _ in sp.recv]
    imgs = [torch.from_numpy(synth.make_image(i, N_DESC)).pin_memory() for i in own_ids]
    imgs_np = [t.numpy() for t in imgs]
    pairs = sp.pairs                      # table row == image id here
    h2d = sum(a.nbytes for a in imgs_np)

    m = SiftMatcher(device=local, profile=True)
    stream = torch.cuda.ExternalStream(m.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    halo_src = {}                                     # one contiguous receive buffer per owning rank
    for row, src in sp.recv:
        halo_src[src] = halo_src.get(src, 0) + N_DESC * 128
    halo_src = {src: torch.empty(nb, dtype=torch.uint8, device=dev) for src, nb in halo_src.items()}

    def halo_exchange():
        """Halo rows come straight out of the owning rank's descriptor pool over NCCL, one message per peer
        (a zero-copy view of the pool when the rows are adjacent there, which they are for whole 256-row images).
        Nothing here waits on the host: the receives are queued on torch's stream and adopted with
        smb_put_images_device_async, so the next match call starts on the pairs that need no halo image and only
        its last sub-batch waits (on the device) for the exchange."""
        if world == 1:
            return
        def send_span(rows):
            spans = [m.image_device_ptr(r) for r in rows]
            if all(spans[k][0] + spans[k][1] * 128 == spans[k + 1][0] for k in range(len(spans) - 1)):
                return torch.as_tensor(_DevView(spans[0][0], sum(n for _, n in spans) * 128), device=dev)
            return torch.cat([torch.as_tensor(_DevView(ptr, n * 128), device=dev) for ptr, n in spans])
        got = sharding.exchange_halo_packed(sp, lambda row: N_DESC * 128, send_span, lambda src, nb: halo_src[src])
        m.put_images_device_async(halo_ids, [got[row].data_ptr() for row in halo_ids], [N_DESC] * len(halo_ids),
                                  torch.cuda.current_stream().cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    # ---- kernel-resident measurement: descriptors already in HBM when the timed region starts
    m.put_images(own_ids, imgs_np)
    halo_exchange()
    matches_per_step = 0
    for _ in range(args.warmup):
        matches_per_step = m.match_pairs_count(pairs)
    sampler = ClockSampler(_physical_gpu_index(local), period=0.01)
    sampler.start()
    barrier()
    dev_ms, score_ms, launches, score_launches, ops = 0.0, 0.0, 0, 0, 0
    for _ in range(args.steps):
        flush_l2()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the halo (NCCL) is issued on torch's stream, the matcher runs on the library's stream: the first event
        # goes where the step's first device work goes, the second after its last
        e0.record(torch.cuda.current_stream() if world > 1 else stream)
        halo_exchange()
        m.match_pairs_count(pairs)
        e1.record(stream)
        e1.synchronize()
        step_ms = e0.elapsed_time(e1)
        t = m.timing()
        dev_ms += step_ms
        score_ms += t["score_ms"]
        launches += t["total_launches"]
        score_launches += t["score_launches"]
        ops = t["ops"]
    barrier()
    clocks = sampler.stop()

    # ---- end to end through the public API: host (pinned) descriptors in, matches out, every step.
    # The images are uploaded asynchronously in 3 chunks (a short first one, so matching can start early) and
    # ONE match_pairs call follows: the library takes the pairs in the order their images land, one sub-batch
    # per upload ticket, each waiting on the device for its own ticket only -- so the copy of chunk k+1 overlaps
    # the matching of chunk k without any host round trip in between.
    n_own = len(own_ids)
    # chunk sizes grow so that the upload of chunk k+1 (~55 GB/s) hides under the matching of chunk k: 16 / 24 / 60
    # images of 100 measured best (tools/e2e_chunks.py)
    # (with 8 ranks uploading at once the host link is slower and a shorter first chunk + four chunks measured better)
    if world == 1:
        first = min(n_own, OVERLAP + 6)
        bounds = sorted(set([0, first, max(first, (2 * n_own) // 5), n_own]))
    else:
        first = min(n_own, OVERLAP + 2)
        bounds = sorted(set([0, first] + [first + ((n_own - first) * c) // 3 for c in (1, 2, 3)]))
    n_chunks = len(bounds) - 1

    def e2e_step():
        m.clear_images()
        for c in range(n_chunks):
            lo, hi = bounds[c], bounds[c + 1]
            m.put_images_async(own_ids[lo:hi], imgs_np[lo:hi])
            if c == 0 and world > 1:
                m.synchronize()       # the halo this rank SENDS is its first overlap-1 images
                halo_exchange()
        return m.match_pairs_count(pairs)

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    e2e_ms = 0.0
    for _ in range(args.steps):
        flush_l2()
        barrier()
        t0 = time.perf_counter()
        total = e2e_step()
        m.synchronize()
        e2e_ms += (time.perf_counter() - t0) * 1e3
    barrier()
    d2h = int(total) * 8 + len(pairs) * 8 + 16

    times = torch.tensor([dev_ms, e2e_ms, score_ms], dtype=torch.float64, device=dev)
    npairs = torch.tensor([float(len(pairs))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(npairs, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, score_ms_max = [float(x) for x in times.tolist()]
    total_pairs = float(npairs.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_sus = peaks.get("bf16_tflops_sustained")
        bf16_burst = peaks.get("bf16_tflops")
        if bf16_sus:
            peak, peak_src = 2.0 * bf16_sus, ("2 x bf16_tflops_sustained of MEASURED_PEAKS.json (kind::i8 issues at twice the "
                                              "kind::f16 rate; kernel timed inside a long step)")
        else:
            peak, peak_src = 2.0 * 1400.0, "2 x 1.4 PFLOP/s sustained bf16 fallback of B200_PROFILING.md"
        launch_ms = score_ms / max(score_launches, 1)
        achieved = ops / (launch_ms * 1e-3) / 1e12 if launch_ms > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        int8_meas = int8_gemm_peak(torch, dev)

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle
            oracle.build()
            cores = oracle.num_procs()
            sample = 4 * cores                     # ~10 s of CPU work on the 16-core box
            v, used, dt, _ = cpu_sample(sample)
            cpu = {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                   "sample": f"{sample} of the workload's 8192x8192 pairs, {dt:.1f} s on {used} threads "
                             f"(oracle/sift_match_oracle.c: COLMAP MatchSiftFeaturesCPU restated, gcc -O3 x86-64-v3)"}

        line = {
            "metric": METRIC, "value": total_pairs * args.steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "pairs_per_step": int(total_pairs), "matches_per_step_rank0": int(matches_per_step),
                "l2": "flushed between timed steps (256 MiB write)",
                "timing": "CUDA events on the library stream around the whole smb_match_pairs call (plan upload, kernels, "
                          "result copy); max over ranks",
                "multi_gpu": "contiguous, cost-balanced image windows (sharding.plan); the overlap-1 halo images are received "
                             "from the next rank's descriptor pool over NCCL send/recv inside every timed step; no "
                             "other collective" if world > 1 else "single GPU",
            },
            "clocks": clocks,
            "e2e": {"value": total_pairs * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "note": "wall clock around clear_images + put_images_async (pinned host descriptors, 3-4 chunks) + "
                            "one match_pairs call (sub-batches wait on the device for their own upload; matches land "
                            "in pinned host memory)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "score_tcgen05_kernel",
                         "algorithmic_ops_per_launch": int(ops), "launch_ms": launch_ms,
                         "peak_source": peak_src, "frac_of_spec_4500": achieved / INT8_SPEC_TOPS,
                         "int8_gemm_measured_here": int8_meas, "bf16_burst": bf16_burst},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
