#!/usr/bin/env python
"""bench.py -- MPEG-1 video decode throughput (frames/s, Gpix/s) on N B200s vs the reference on host cores.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own C on the host cores

Workload (BASELINE.json configs[3], weak-scaled to configs[4]): STREAMS_PER_GPU = 64 independent
1920x1080 MPEG-1 I+P elementary streams of PICTURES = 60 pictures each per GPU (demuxed from
synthetic MPEG-TS clips made by tools/gen_streams.py: seeded content encoded by cv2's FFmpeg;
DISTINCT clips are encoded and replicated into separate buffers to reach 64 streams).
One STEP = every picture of every stream once: start-code scan + VLC parse (stage 1) +
IDCT/MC reconstruction (stage 2).

  value  frames/s with the elementary streams already resident in HBM (jsmpeg_b200_batch_rewind
         forgets the start-code index and all parsed records, so scan + parse + reconstruct are
         all inside the timed region), planes left in HBM.
  e2e    the same through the reference-facing call sequence with HOST buffers: per step every
         stream is written again from host memory (get_write_ptr/memcpy/did_write -> H2D) and
         every decoded picture's Y/Cr/Cb planes are copied back to pinned host memory (D2H).
         The 64 streams are driven as E2E_GROUPS independent BatchDecoders from host threads, so
         that one group's PCIe copy-out overlaps another group's write + parse.
  roofline  stage-2 kernel: algorithmic bytes / CUDA-event time of its launches.
  cpu_baseline  oracle/_ref (the unmodified reference C, compiled in place) on all host threads.

Timing: wall clock between torch.cuda.synchronize() + barrier on both sides of exactly K steps
(every step ends host-synchronised), max over ranks; per-kernel times are CUDA events recorded on
the launching stream inside the library.  Inputs are much larger than L2 (ES ~0.8 GB, records
~20 GB per step), so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

WIDTH = int(os.environ.get("BENCH_WIDTH", 1920))    # the env overrides exist for the CPU smoke test of this file
HEIGHT = int(os.environ.get("BENCH_HEIGHT", 1080))
STREAMS_PER_GPU = int(os.environ.get("BENCH_STREAMS", 64))
PICTURES = int(os.environ.get("BENCH_PICTURES", 60))
DISTINCT = int(os.environ.get("BENCH_DISTINCT", 8))
NOISE = 9
E2E_GROUPS = int(os.environ.get("BENCH_E2E_GROUPS", 16))
VALUE_GROUPS = int(os.environ.get("BENCH_VALUE_GROUPS", 1))


def env_int(name, default):
    return int(os.environ.get(name, default))


def _encode_one(seed):
    import gen_streams
    gen_streams.make_clip_ts(WIDTH, HEIGHT, PICTURES, seed=seed, noise=NOISE)
    return seed


def load_streams(rank, world):
    """DISTINCT elementary streams (bytes).  Rank 0 encodes (in parallel processes), the others
    read the cache."""
    import gen_streams
    seeds = [1234 + i for i in range(DISTINCT)]
    if rank == 0:
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=min(len(seeds), os.cpu_count() or 1)) as ex:
            list(ex.map(_encode_one, seeds))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    out = []
    for s in seeds:
        packets = gen_streams.make_clip_es(WIDTH, HEIGHT, PICTURES, seed=s, noise=NOISE)
        out.append(b"".join(p for _, p in packets))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.t_mark = 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Samples that arrive from now on belong to the timed region."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], 0, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t_arrival, line in self.lines:
            if t_arrival < self.t_mark:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
                power = max(power, float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "power_w_max": power or None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per stage-2 launch from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "recon_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline

def run_reference_cpu(clips, threads, loops):
    """The reference's own C decoder (oracle/_ref, built from /root/reference/src/wasm/*.c in the
    build container) on `threads` host threads; falls back to our CPU port of it (oracle/) when the
    reference build is absent.  Returns (frames, seconds, kind)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libjsmpeg_ref.so")
    n = len(clips)
    if os.path.exists(ref):
        lib = ctypes.CDLL(ref)
        lib.ref_bench_run.restype = ctypes.c_long
        lib.ref_bench_run.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_uint), ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        arr = (ctypes.c_char_p * n)(*clips)
        lens = (ctypes.c_uint * n)(*[len(c) for c in clips])
        sec = ctypes.c_double()
        frames = lib.ref_bench_run(arr, lens, n, threads, loops, ctypes.byref(sec))
        return frames, sec.value, "reference"
    # port: the oracle through ctypes from Python threads (ctypes releases the GIL in the C calls)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    from jsmpeg_b200 import capi
    lib = capi.load_library(os.path.join(ROOT, "oracle", "liboracle.so"))
    counts = [0] * threads

    def work(t):
        for _ in range(loops):
            for c in range(t, n, threads):
                d = lib.mpeg1_decoder_create(len(clips[c]) + 16, 2)
                ctypes.memmove(lib.mpeg1_decoder_get_write_ptr(d, len(clips[c])), clips[c], len(clips[c]))
                lib.mpeg1_decoder_did_write(d, len(clips[c]))
                while lib.mpeg1_decoder_decode(d):
                    counts[t] += 1
                lib.mpeg1_decoder_destroy(d)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return sum(counts), time.perf_counter() - t0, "port"


def truncate_pictures(es, n_pictures):
    """The first n_pictures pictures of an elementary stream (cut at the next picture start code)."""
    pos, count = 0, 0
    while True:
        pos = es.find(b"\x00\x00\x01\x00", pos)
        if pos < 0:
            return es
        if count == n_pictures:
            return es[:pos]
        count += 1
        pos += 4


def cpu_sample(clips, threads, pictures=None):
    """Bounded sample for the reference timing: one clip per host thread (replicated round-robin
    from the distinct clips), optionally only its first `pictures` pictures."""
    if pictures is not None:
        clips = [truncate_pictures(c, pictures) for c in clips]
    return [clips[i % len(clips)] for i in range(threads)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="value leg only (used for the ncu launch list)")
    args = ap.parse_args()

    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    pix = WIDTH * HEIGHT
    base_config = {
        "workload": f"{STREAMS_PER_GPU} independent {WIDTH}x{HEIGHT} MPEG-1 I+P elementary streams x {PICTURES} pictures per GPU "
                    f"(BASELINE configs[3]; weak-scaled x N GPUs = configs[4]), GOP 12, one slice per picture",
        "streams_per_gpu": STREAMS_PER_GPU, "pictures_per_stream": PICTURES, "distinct_clips": DISTINCT,
        "parallelism": f"streams sharded {STREAMS_PER_GPU} per GPU, no collective",
        "l2": "inputs larger than L2 (ES + records per step >> 126 MB); no explicit flush",
    }

    if args.impl == "reference":
        if rank != 0:
            return 0
        clips = load_streams(0, 1)
        threads = os.cpu_count() or 1
        ref_pictures = min(PICTURES, env_int("BENCH_REF_PICTURES", 30))  # bounded sample: keeps a step at ~5 s
        sample = cpu_sample(clips, threads, ref_pictures)
        for _ in range(min(args.warmup, 1)):
            run_reference_cpu(sample[:threads], threads, 1)
        frames = 0
        seconds = 0.0
        kind = "reference"
        for _ in range(args.steps):
            f, s, kind = run_reference_cpu(sample, threads, 1)
            frames += f
            seconds += s
        fps = frames / seconds
        desc = (f"{threads} threads x 1 clip x first {ref_pictures} of {PICTURES} pictures per step "
                f"({len(clips)} distinct {WIDTH}x{HEIGHT} clips, same streams as the GPU arm)")
        print(json.dumps({
            "impl": "reference", "metric": "MPEG-1 video decode frames/s", "value": fps, "unit": "frames/s",
            "gpix_per_s": fps * pix / 1e9, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * seconds / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": base_config,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": desc},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return 0

    # NCCL is used for the barrier / counter reduction only; keep its banner off stdout (one JSON line)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    from jsmpeg_b200.batch import OUT_DEVICE, OUT_HOST, BatchDecoder

    clips = load_streams(rank, world)
    streams = [clips[(i + rank) % len(clips)] for i in range(STREAMS_PER_GPU)]
    # value: VALUE_GROUPS decoders share the 64 streams; with more than one, each is driven by its own
    # host thread and the groups run free (one group's reconstruct/expand overlaps another's walk)
    vgroups = [list(range(g, STREAMS_PER_GPU, VALUE_GROUPS)) for g in range(VALUE_GROUPS)]
    value_decoders = [BatchDecoder(len(g), device=local_rank, max_slots=len(g) * PICTURES + 8) for g in vgroups]
    for dec, g in zip(value_decoders, vgroups):
        for j, i in enumerate(g):
            dec.write(j, streams[i])
        dec.upload()  # elementary streams resident in HBM before the timed region
    bd = value_decoders[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        bd.rewind()
        return bd.decode(PICTURES, OUT_DEVICE)

    # e2e: E2E_GROUPS smaller decoders, one host thread each (ctypes releases the GIL in the C calls)
    groups = [list(range(g, STREAMS_PER_GPU, E2E_GROUPS)) for g in range(E2E_GROUPS)]
    e2e_decoders = [BatchDecoder(len(g), device=local_rank, max_slots=len(g) * PICTURES + 8) for g in groups]
    for dec, g in zip(e2e_decoders, groups):   # sequence headers parsed once, like a decoder that has seen its stream start
        for j, i in enumerate(g):
            dec.write(j, streams[i])

    def run_e2e(n_steps):
        """n_steps whole-workload steps: every group's thread runs its n_steps back to back (no
        per-step join), so one group's PCIe copy-out overlaps the others' write + parse."""
        counts = [0] * E2E_GROUPS

        def work(k):
            dec = e2e_decoders[k]
            for _ in range(n_steps):
                dec.reset()
                for j, i in enumerate(groups[k]):
                    dec.write(j, streams[i])
                counts[k] += dec.decode(PICTURES, OUT_HOST)

        threads = [threading.Thread(target=work, args=(k,)) for k in range(E2E_GROUPS)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        return sum(counts)

    def timed(run, steps, warmup, decoders):
        """run(n) performs n steps and returns the pictures decoded."""
        sampler = ClockSampler(local_rank)
        sampler.start()  # started before the warm-up so that it is already streaming samples
        run(warmup)
        barrier()
        for d in decoders:
            d.reset_stats()
        sampler.mark()
        t0 = time.perf_counter()
        frames = run(steps)
        barrier()
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        st = {}
        for d in decoders:
            for k, v in d.stats().items():
                st[k] = st.get(k, 0) + v
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fr = torch.tensor([frames], dtype=torch.float64, device="cuda")
            dist.all_reduce(fr, op=dist.ReduceOp.SUM)
            dt, frames = float(t.item()), int(fr.item())
        return frames, dt, st, clocks

    def run_device(n_steps):
        if VALUE_GROUPS == 1:
            return sum(step_device() for _ in range(n_steps))
        counts = [0] * VALUE_GROUPS

        def work(k):
            dec = value_decoders[k]
            for _ in range(n_steps):
                dec.rewind()
                counts[k] += dec.decode(PICTURES, OUT_DEVICE)

        threads = [threading.Thread(target=work, args=(k,)) for k in range(VALUE_GROUPS)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        return sum(counts)

    frames, dt, st, clocks = timed(run_device, args.steps, args.warmup, value_decoders)
    e_steps = args.steps
    if args.no_e2e:
        e_frames, e_dt, e_st, e_clocks = 0, 1.0, {"h2d_bytes": 0, "d2h_bytes": 0}, None
    else:
        e_frames, e_dt, e_st, e_clocks = timed(run_e2e, e_steps, 3, e2e_decoders)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    fps = frames / dt
    e_fps = e_frames / e_dt
    peak, peak_src = measured_peak()
    recon_s = st["recon_ms"] / 1e3
    achieved = st["algorithmic_bytes"] / recon_s / 1e9 if recon_s > 0 else 0.0
    traffic = recorded_traffic()
    out = {
        "metric": "MPEG-1 video decode frames/s", "value": fps, "unit": "frames/s",
        "gpix_per_s": fps * pix / 1e9,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": base_config,
        "clocks": clocks,
        "e2e": {"value": e_fps, "unit": "frames/s", "gpix_per_s": e_fps * pix / 1e9,
                "h2d_bytes_per_step": e_st["h2d_bytes"] // e_steps,
                "d2h_bytes_per_step": e_st["d2h_bytes"] // e_steps,
                "ms_per_step": 1e3 * e_dt / e_steps, "steps": e_steps, "warmup": 3,
                "host_threads": E2E_GROUPS, "clocks": e_clocks},
        "gpu_launches": st["kernel_launches"], "value_host_threads": VALUE_GROUPS,
        "roofline": {
            "kernel": "reconstruct_kernel (stage 2: IDCT + motion compensation + add/clamp)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if peak else None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": st["algorithmic_bytes"] / max(1, st["recon_launches"]),
            "avg_launch_ms": st["recon_ms"] / max(1, st["recon_launches"]),
            "launches": st["recon_launches"],
            "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
            "traffic_source": traffic["source"] if traffic else None,
        },
        "stage_ms_per_step": {"scan": st["scan_ms"] / args.steps, "parse": st["parse_ms"] / args.steps,
                              "reconstruct": st["recon_ms"] / args.steps},
        "stage1": {"es_mbit_per_s": st["es_bytes"] * 8 / (st["parse_ms"] / 1e3) / 1e6 if st["parse_ms"] else None,
                   "pictures_per_s": st["pictures"] / (st["parse_ms"] / 1e3) if st["parse_ms"] else None,
                   "parse_errors": st["parse_errors"],
                   "walk": os.environ.get("JSMPEG_B200_WALK") or "lanes",
                   "lane_walk_pictures_per_step": st.get("lane_walk_pictures", 0) / args.steps},
    }
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        sample = cpu_sample(clips, threads)
        loops = 1
        f, s, kind = run_reference_cpu(sample, threads, loops)
        out["cpu_baseline"] = {"value": f / s, "unit": "frames/s", "gpix_per_s": f / s * pix / 1e9, "cores": threads,
                               "kind": kind,
                               "sample": f"{threads} threads x 1 clip x {PICTURES} pictures x {loops} loops of the same {WIDTH}x{HEIGHT} clips"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
