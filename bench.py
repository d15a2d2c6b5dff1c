#!/usr/bin/env python
"""bench.py -- throughput of the pbrs path-tracing inner loop on B200s, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1..c5] [--impl reference]

A "step" is one full frame of the workload (every pixel, every sample, the bounce loop).  The
scene is synthetic and procedurally generated (pbrs_b200/scenes.py), built once before timing.

  value   Msamples/s, whole job, film left on the device (pbrs_render_device on torch's stream,
          timed with CUDA events, max over ranks).  N > 1: the frame's 64x64 tiles (or its sample
          indices for C5) are split over the ranks and the partial films summed with an NCCL
          reduce inside the timed region -- total work fixed ("strong").
  e2e     the same metric through the public API with a HOST film: pbrs_render into a page-locked
          film (N = 1), pbrs_b200.dist.render_sharded into one shared host film (N > 1: every
          rank copies its own tiles straight into it; C5: NCCL reduce, then one copy out).  The
          per-step host->device traffic and the device->host film copy are inside the timed region.
  film_crc32  CRC-32 of rank 0's film after the last timed step (device leg) and of the host film
          (e2e leg): for a tile split both are the same value at N = 1, 2, 4, 8.
  roofline  the closest-hit traversal kernel (k_extend): algorithmic bytes (SURVEY.md 8d formula,
          counters from one untimed PBRS_FLAG_COUNT_TRAVERSAL frame) / its summed launch time
          (CUDA events between launches, PBRS_FLAG_TIME_STAGES frames) vs the measured HBM copy peak.
  cpu_baseline  the C++ oracle (a port: the Rust reference cannot be built here) on all host
          threads, on a bounded row subset of the same frame.

`--impl reference` times that CPU path alone and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Msamples/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md


def workload(name):
    from pbrs_b200 import scenes
    gen, integrator, msaa = scenes.CONFIGS[name]
    return gen, integrator, msaa, scenes.WORKLOAD_NAMES[name]


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def ncu_evidence(workload):
    """Launch-weighted means of the closest-hit traversal kernel over all bounces of two whole batches,
    from the committed ncu capture of this workload at this round's kernels (profiles/r2_traffic.json,
    written from the .ncu-rep by tools/ncu_traffic.py), or {}."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            d = json.load(f).get(workload, {})
        return dict(d.get("extend", {}), source=d.get("source"))
    except Exception:
        return {}


def extend_bytes(st):
    """32 (ray read) + 64/inner node + 48/triangle + 16/sphere + 128/instance + 32 (hit write)."""
    n, t, s, i = st["trav_extend"]
    return 64 * st["n_rays_extend"] + 64 * n + 48 * t + 16 * s + 128 * i


def shadow_bytes(st):
    n, t, s, i = st["trav_shadow"]
    return 64 * st["n_rays_shadow"] + 64 * n + 48 * t + 16 * s + 128 * i


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["sm_max_mhz"] = float(max(mx)); out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def cpu_sample(handle, integrator, msaa, target_s=24.0):
    """Times the oracle on rows y0, y0+step, ... of the frame, step chosen for ~target_s of CPU work."""
    from oracle import oracle_ffi
    H, W = handle.height, handle.width
    t0 = time.time()
    probe_rows = max(1, H // 256)
    _, st = oracle_ffi.render_rows(handle, max(1, H // probe_rows), integrator=integrator, msaa=msaa)
    probe = max(time.time() - t0, 1e-4)
    per_row = probe / max(1, len(range(0, H, max(1, H // probe_rows))))
    rows = int(min(H, max(1, target_s / per_row)))
    step = max(1, H // rows)
    t0 = time.time()
    _, st = oracle_ffi.render_rows(handle, step, integrator=integrator, msaa=msaa)
    dt = time.time() - t0
    n_rows = len(range(0, H, step))
    return st, dt, n_rows, step


def run_reference(args):
    """The CPU path alone (oracle port of the reference, all host threads), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_ffi
    gen, integrator, msaa, wname = workload(args.workload)
    api = oracle_ffi.load()
    h = gen().realize(api)
    cores = api["get_threads"]()
    H, W = h.height, h.width
    # size the per-step sample for ~8 s
    st, dt, n_rows, step = cpu_sample(h, integrator, msaa)  # the same bounded sample as the cpu_baseline leg of the GPU arm
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        _, st = oracle_ffi.render_rows(h, step, integrator=integrator, msaa=msaa)
        if i >= args.warmup:
            times.append(time.time() - t0)
    dt = float(np.mean(times))
    value = st["n_samples"] / dt / 1e6
    rays = (st["n_rays_extend"] + st["n_rays_shadow"]) / dt / 1e6
    sample = f"rows 0,{step},.. ({n_rows} of {H}) of the frame at full spp, per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wname, "integrator": integrator, "spp": msaa * msaa, "max_depth": 5, "sample": sample},
        "mrays_per_s": rays,
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "C++ restatement of the pbrs CPU path (oracle/); the Rust reference cannot be built here"},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # default: C4, the BASELINE config quoted "over 1/2/4/8 B200" (it fits one GPU); see DESIGN.md section 6
    ap.add_argument("--workload", default=os.environ.get("PBRS_BENCH_WORKLOAD", "c4"), choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--paths-in-flight", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--frame-scale", type=float, default=1.0, help="shrink the frame (profiling runs only; 1.0 = the BASELINE size)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    # stdout must carry exactly ONE JSON line.  Native libraries (NCCL prints its version banner
    # on fd 1) write there too, so fd 1 is pointed at stderr for the whole run and the JSON line
    # goes to a private duplicate of the original stdout.
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    from pbrs_b200 import _ffi
    from pbrs_b200.dist import SharedHostFilm, film_reduce, render_sharded, split_for

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pbrs_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    gen, integrator, msaa, wname = workload(args.workload)
    api = _ffi.load()
    sd = gen(args.frame_scale)
    if args.frame_scale != 1.0:
        wname += f" [frame scaled x{args.frame_scale}: NOT the BASELINE size]"
    h = sd.realize(api)
    W, H = h.width, h.height
    spp = msaa * msaa
    split = split_for(args.workload)
    kw = dict(integrator=integrator, msaa=msaa, max_depth=5, rank=rank, world_size=world, split=split, paths_in_flight=args.paths_in_flight)
    film = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()

    def step_device():
        h.render_device(film.data_ptr(), stream=stream.cuda_stream, flags=(8 if (world > 1 and split == "samples") else 0), **kw)
        if world > 1:
            film_reduce(film, spp if split == "samples" else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # untimed: traversal counters of this rank's share (for the roofline bytes)
    st_count = h.render_device(film.data_ptr(), stream=stream.cuda_stream, want_stats=True, flags=1, **kw)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # sampled from the warm-up on, so that short frames still see samples under load
    for _ in range(args.warmup):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    film_crc = zlib.crc32(film.cpu().numpy().tobytes()) if rank == 0 else 0  # the whole frame, as the last timed step left it
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    # roofline leg: per-stage launch times of this rank's share
    st_time = None
    for _ in range(1 if ms_step > 2000.0 else max(1, min(args.steps, 3))):
        s = h.render_device(film.data_ptr(), stream=stream.cuda_stream, want_stats=True, flags=2, **kw)
        if st_time is None:
            st_time = s
        else:
            for k in ("ms_generate", "ms_extend", "ms_shade", "ms_shadow", "ms_accumulate", "ms_total"):
                st_time[k] = min(st_time[k], s[k])

    # e2e leg: one host film shared by the ranks (page-locked), copies inside the timed region
    host = SharedHostFilm(H, W, api)
    e2e_kw = dict(integrator=integrator, msaa=msaa, max_depth=5, split=split, paths_in_flight=args.paths_in_flight, host_film=host,
                  device_film=film if (world > 1 and split == "samples") else None)
    for _ in range(1 if ms_step > 2000.0 else args.warmup):
        render_sharded(h, **e2e_kw)
    host.array[...] = 0.0
    barrier()
    t0 = time.perf_counter()
    step_times = []
    for _ in range(args.steps):
        t_step = time.perf_counter()
        render_sharded(h, **e2e_kw)
        step_times.append((time.perf_counter() - t_step) * 1e3)
    barrier()
    e2e_s = time.perf_counter() - t0
    if os.environ.get("PBRS_BENCH_DEBUG"):
        print("e2e per-step ms:", " ".join(f"{t:.2f}" for t in step_times), file=sys.stderr)
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_step = float(te.item()) / args.steps
    e2e_crc = zlib.crc32(host.array.tobytes()) if rank == 0 else 0
    host_pinned = host.registered
    host.close()

    # the sampler ran from the warm-up through the timed, roofline and e2e legs (all under load)
    clk = clocks.stop() if rank == 0 else None

    # whole-job counts
    cnt = torch.tensor([st_count["n_samples"], st_count["n_rays_extend"], st_count["n_rays_shadow"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt)
    n_samples, n_ext, n_sh = [float(x) for x in cnt.tolist()]

    if rank == 0:
        import ctypes as C
        from pbrs_b200 import _capi as K
        sizeof_opts = C.sizeof(K.RenderOpts)  # what crosses host -> device per step: the launch parameters (the tile list is cached on the device)
        peak, peak_src = hbm_peak()
        ext_b = extend_bytes(st_count)
        n_launch = max(1, st_time["launches_extend"])
        ext_ms = st_time["ms_extend"]
        achieved = ext_b / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        ev = ncu_evidence(args.workload) if args.frame_scale == 1.0 else {}
        limiter = "not profiled for this workload"
        if ev.get("lanes_per_inst"):
            limiter = (f"the SM's L1 load pipe, issue slots and SIMT efficiency, not DRAM (ncu: L1 data-pipe wavefronts {ev.get('l1_pipe_pct', 0):.0f} % of peak, DRAM at "
                       f"{ev.get('dram_pct_of_peak', 0):.1f} % of peak, issue slots {ev.get('issue_active_pct', 0):.0f} % busy, "
                       f"{ev.get('lanes_per_inst', 0):.1f} of 32 lanes per instruction, {ev.get('occupancy_pct', 0):.0f} % occupancy)")
        # thread instructions per ray of the closest-hit walk: ncu's instruction rate (warp instructions x lanes per millisecond of kernel
        # time, launch-weighted over two whole batches) x this run's kernel time per ray
        thread_inst_per_ray = None
        if ev.get("warp_inst_per_launch") and ev.get("lanes_per_inst") and ev.get("ms_per_launch") and st_count["n_rays_extend"]:
            thread_inst_per_ray = ev["warp_inst_per_launch"] * ev["lanes_per_inst"] / ev["ms_per_launch"] * ext_ms / st_count["n_rays_extend"]
        sh_ms = st_time["ms_shadow"]
        line = {
            "metric": METRIC, "value": n_samples / (ms_step * 1e-3) / 1e6, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wname, "integrator": integrator, "spp": spp, "max_depth": 5, "resolution": [W, H],
                       "split": split if world > 1 else "none", "l2": "path state streamed per batch exceeds L2 (inputs larger than L2)",
                       "paths_in_flight": args.paths_in_flight or (1 << 24)},
            "mrays_per_s": (n_ext + n_sh) / (ms_step * 1e-3) / 1e6,
            "frame_ms": ms_step,
            "clocks": clk,
            "film_crc32": f"{film_crc:08x}",
            "e2e": {"value": n_samples / e2e_step / 1e6, "unit": "Msamples/s", "ms_per_step": e2e_step * 1e3,
                    "h2d_bytes_per_step": int(sizeof_opts), "d2h_bytes_per_step": int(W * H * 12), "film_crc32": f"{e2e_crc:08x}",
                    "host_film": "one page-locked film shared by the ranks" if host_pinned else "pageable (registration failed)",
                    "path": "pbrs_render" if world == 1 else ("render_sharded: own tiles -> shared host film" if split == "tiles"
                                                              else "render_sharded: NCCL reduce -> one copy out")},
            "gpu_launches": int(st_time["launches"]) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_extend (closest-hit TLAS/BLAS walk)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": ev.get("dram_bytes_per_launch"), "limiter": limiter,
                         "l2_bytes_per_launch": ev.get("l2_bytes_per_launch"), "local_mem_bytes_per_launch": ev.get("local_mem_bytes_per_launch"),
                         "lanes_per_instruction": ev.get("lanes_per_inst"), "occupancy_pct": ev.get("occupancy_pct"), "issue_active_pct": ev.get("issue_active_pct"),
                         "l1_data_pipe_pct": ev.get("l1_pipe_pct"), "thread_instructions_per_ray": thread_inst_per_ray,
                         "ncu_source": ev.get("source"),
                         "algorithmic_bytes_per_launch": ext_b / n_launch, "ms_per_launch": ext_ms / n_launch, "launches_per_step": int(n_launch),
                         "bytes_per_ray": ext_b / max(1.0, st_count["n_rays_extend"]),
                         "note": "frac = algorithmic bytes (SURVEY 8d formula: 64 B/node, 48 B/triangle, 16 B/sphere, 128 B/instance entry, 64 B/ray) / summed launch "
                                 "time / the measured HBM copy peak -- the contract's fraction.  traffic, l2_bytes, local_mem_bytes, lanes, occupancy and issue "
                                 "utilisation are launch-weighted ncu means over all five bounces of two whole batches (profiles/): the BVH is largely "
                                 "L2-resident, so DRAM traffic << algorithmic bytes and what limits the kernel is `limiter`.  compute-sanitizer is closed on "
                                 "this pool: memory safety rests on the commit-time depth bound and parity against the oracle"},
            "stages_ms": {k: st_time[k] for k in ("ms_generate", "ms_extend", "ms_shade", "ms_shadow", "ms_accumulate", "ms_total")},
            "shadow_kernel": {"achieved": shadow_bytes(st_count) / (sh_ms * 1e-3) / 1e9 if sh_ms > 0 else 0.0, "unit": "GB/s"},
            "would_panic": st_count["would_panic"],
        }
        if world == 1 and not args.no_cpu:
            from oracle import oracle_ffi
            hc = sd.realize(oracle_ffi.load())
            stc, dt, n_rows, step = cpu_sample(hc, integrator, msaa)
            line["cpu_baseline"] = {"value": stc["n_samples"] / dt / 1e6, "unit": "Msamples/s", "cores": oracle_ffi.load()["get_threads"](), "kind": "port",
                                    "sample": f"rows 0,{step},.. ({n_rows} of {H}) of the same frame at full spp, {dt:.1f} s",
                                    "mrays_per_s": (stc["n_rays_extend"] + stc["n_rays_shadow"]) / dt / 1e6}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
