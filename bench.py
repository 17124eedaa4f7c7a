#!/usr/bin/env python3
"""bench.py -- RTjpeg 720x576 YUV420 decoded frames/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...       # the reference's CPU decoder
    torchrun --nproc-per-node N ... bench.py --gpus N ...             # one rank per GPU

A "step" is one pass of the hot path (K1 scan -> K3 resolve -> K2 idct -> K2b hard blocks) over one batch of
4096 synthetic 720x576 frames (BASELINE.json configs[1]: intra-only, Q=128).  The batch is
produced once, outside every timed region, by the reference's own RTjpeg_compress
(oracle/_ref, as north_star prescribes for the synthetic streams) from the seeded source in
oracle/ref_driver.c.  Multi-GPU is weak scaling: every rank decodes its own 4096-frame
shard (its own seed); shards are independent, there is no data-path collective.

value     whole-job frames/s, packets and descriptors already resident in HBM, frames left in HBM.
e2e       same metric through the C ABI's host entry point (rtjgpu_decode_host): packets in pinned
          host memory, frames returned to pinned host memory, both copies inside the timed region.
roofline  dominant kernel: algorithmic bytes (payload read + planes written) / its mean device
          time over the timed region (CUDA events recorded by the library on the launch stream).
cpu_baseline  the unmodified reference decoder (oracle/_ref) on this box's host cores.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, QUALITY, FRAMES = 720, 576, 128, 4096
METRIC = "RTjpeg 720x576 YUV420 decoded frames/sec"
WORKLOAD = "configs[1]: RTjpeg 720x576 YUV420 intra-only Q=128, 4096-frame synthetic batch per GPU"


def make_workload(frames: int, seed: int):
    """Synthetic clip -> packets by the reference compressor (not timed)."""
    from oracle import oracle as O
    clip = O.make_clip(W, H, QUALITY, key_rate=-1, noise_y=2, noise_c=0, seed=seed)
    t0 = time.time()
    stream, offsets = O.encode_clip(clip, frames, threads=min(os.cpu_count() or 1, 64))
    return stream, offsets, time.time() - t0


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_decode_fps(stream, offsets, frames: int, threads: int, reps: int):
    """Reference RTjpeg_decompress over `frames` frames, `threads` private decoders, frames striped
    (every intra frame is its own segment).  Returns (best fps, all fps)."""
    from oracle import oracle as O
    sub = offsets[:frames + 1]
    seg = np.arange(frames + 1, dtype=np.int32)
    fps = []
    for _ in range(reps):
        secs, _ = O.ref_decode_threaded(stream, sub, seg, W, H, threads, zero_init=False, keep=False)
        fps.append(frames / secs)
    return max(fps), fps


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            getattr(nv, "nvmlClocksEventReasonApplicationsClocksSetting", 0x2): "applications_clocks_setting",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu summary, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(kernel, {}).get("dram_bytes_per_launch_configs1")
    except Exception:  # noqa: BLE001
        return None


def run_reference_arm(args, rank, world):
    if rank != 0:
        return                                           # rank 0 alone runs and prints the CPU arm
    stream, offsets, _ = make_workload(args.frames, seed=1)
    threads = host_threads()
    sample = min(args.frames, args.cpu_sample)
    for _ in range(args.warmup):
        reference_decode_fps(stream, offsets, min(sample, 256), threads, 1)
    t0 = time.time()
    _, fps = reference_decode_fps(stream, offsets, sample, threads, args.steps)
    wall = time.time() - t0
    value = sample * args.steps / sum(sample / f for f in fps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": f"{sample} of the {args.frames} frames per step on the host CPU"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "reference",
                         "sample": f"{args.steps} x {sample} frames, unmodified lib/RTjpeg.c RTjpeg_decompress, "
                                   f"{threads} threads each with a private decoder and plane set"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per GPU per step (default: the metric's 4096)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="host-path steps (default: min(steps, 10))")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="frames per CPU-baseline pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import gmerlin_avdecoder_b200 as g
    from gmerlin_avdecoder_b200 import device as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the decoder has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    F = args.frames
    fsz = W * H * 3 // 2
    stream, offsets, gen_s = make_workload(F, seed=1 + rank)
    desc, _ = g.plan(stream, offsets)
    ctx = g.BatchContext(local)
    batch = D.upload(stream, desc, W, H, device=local)
    algo_bytes = batch.payload_bytes + F * fsz            # payload read + planes written

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps --------------------------------------------------------------
    ctx.enable_timing(True)
    for _ in range(max(args.warmup, 3)):
        D.decode(ctx, batch)
    barrier()
    info = ctx.batch_info()
    assert info.bad_frames == 0
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        D.decode(ctx, batch)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    stage = {"scan": [], "resolve": [], "idct": []}
    for i in range(min(args.steps, 256)):
        t = ctx.timing_at(i)
        stage["scan"].append(t.scan_ms); stage["resolve"].append(t.resolve_ms); stage["idct"].append(t.idct_ms)
    stage_ms = {k: sum(v) / len(v) for k, v in stage.items()}
    ctx.enable_timing(False)

    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * F * args.steps / (ms_max * 1e-3)

    # ---- end to end through the host entry point -----------------------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    h_in = torch.from_numpy(stream).pin_memory()
    h_out = torch.empty((F, fsz), dtype=torch.uint8).pin_memory()
    flags = g.HOST_IN_PINNED | g.HOST_OUT_PINNED
    ctx.decode_host(h_in.numpy(), offsets, h_out.numpy(), flags=flags)          # warm-up (allocations)
    ctx.decode_host(h_in.numpy(), offsets, h_out.numpy(), flags=flags)
    barrier()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.decode_host(h_in.numpy(), offsets, h_out.numpy(), flags=flags)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_launches = ctx.launch_count() - l0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * F * e2e_steps / float(t_e.item())
    # the host path's last chunk must equal the device path's frames (same bytes, two routes)
    assert torch.equal(h_out[-1].cuda(), batch.out[-1]), "host path and device path disagree"

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        dom = max(stage_ms, key=stage_ms.get)
        kname = {"scan": "rtj_scan_chunk_kernel", "resolve": "rtj_resolve_kernel", "idct": "rtj_idct_kernel"}[dom]
        achieved = algo_bytes / (stage_ms[dom] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "frames_per_gpu_per_step": F, "quality": QUALITY,
                "payload_bytes_per_frame": batch.payload_bytes / F, "planar_bytes_per_frame": fsz,
                "l2": "inputs larger than L2: every step reads %.0f MB and writes %.0f MB"
                      % (batch.payload_bytes / 1e6, F * fsz / 1e6),
                "stream_source": "reference RTjpeg_compress on a seeded synthetic clip (%.1f s, untimed)" % gen_s,
            },
            "clocks": clocks,
            "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "frames/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": int(stream.size + desc.nbytes), "d2h_bytes_per_step": int(F * fsz),
                    "gpu_launches": e2e_launches,
                    "api": "rtjgpu_decode_host, pinned host buffers in and out"},
            "roofline": {
                "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(kname),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": stage_ms[dom],
                "stage_ms": stage_ms,
                "whole_path_frac": algo_bytes / (ms_max / args.steps * 1e-3) / 1e9 / peak,
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            sample = min(F, args.cpu_sample)
            best, fps = reference_decode_fps(stream, offsets, sample, threads, 3)
            line["cpu_baseline"] = {
                "value": best, "unit": "frames/s", "cores": threads, "kind": "reference",
                "sample": f"best of 3 passes over the first {sample} frames of the same batch, unmodified "
                          f"lib/RTjpeg.c RTjpeg_decompress, {threads} threads each with a private decoder",
                "all": fps,
            }
        _emit(line)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def _emit(line: dict) -> None:
    """The one JSON line, on the process's real stdout."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


_REAL_STDOUT = sys.stdout

if __name__ == "__main__":
    # Libraries write banners to file descriptor 1 (NCCL prints "NCCL version ..." there when its communicator comes up):
    # everything that is not the JSON line goes to stderr instead.
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    main()
