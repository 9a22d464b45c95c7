#!/usr/bin/env python
"""bench.py -- throughput of the B200 wave_tracer hot path on BASELINE.json's metric (Msamples/sec).

Workload (config.workload): BASELINE.json configs[1], scenes/diffraction_simple/double_slits.xml -D res=1440,spp=1024,pattern=true
(film 1440x360, 5.31e8 samples), geometry/emitters/sensor restated procedurally (wave_tracer_b200/scenes.py), integrator plt_bdpt with
Fraunhofer FSD, max_depth 16, as the reference file selects (double_slits.xml:42-44).  `--integrator plt_path` runs the same geometry
with plt_path forward + UTD (as double_slits_and_reflectors.xml does).  A "step" renders the full film at `--spp-per-step` samples per
element (sample indices [step*S, step*S+S) of the 1024): cost is linear in spp, samples are independent.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N>1; one rank per GPU; NCCL film reduce every step)
  python bench.py --impl reference ...                     the CPU implementation of the same path (oracle port; the reference
                                                           itself cannot be built here) on all host cores, bounded sample per step
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RES, SPP = 1440, 1024
WORKLOAD = "diffraction_simple/double_slits res=1440 spp=1024 pattern=true (film 1440x360, lambda=0.05mm, procedural restatement)"
INTEGRATORS = {"plt_bdpt": "plt_bdpt (MIS, emitter+sensor direct), Fraunhofer FSD, max_depth 16 -- the reference file's integrator",
               "plt_path": "plt_path forward + UTD FSD, max_depth 16, RR off"}
# --workload: the default is BASELINE.json configs[1]; the others are the remaining GPU configs restated procedurally (their meshes are LFS stubs)
WORKLOADS = {
    "double_slits": dict(res=1440, integrator=None, name=WORKLOAD),
    "etoile": dict(res=720, integrator="plt_path", name="sionna_etoile/etoile res=720 spp=1024 wavelength=10GHz (film 720x540; plt_path forward + UTD, max_depth 16, RR off; "
                   "SYNTHETIC geometry: ground plane + 562 extruded boxes (6746 triangles) on a seeded street plan, the file's ITU materials, emitter and sensor)"),
    "cornell": dict(res=1440, integrator="plt_bdpt", name="cornell-box/box res=1440 spp=1024 (film 1440x1440; plt_bdpt max_depth 8; SYNTHETIC: texture-free procedural variant -- 5 walls, "
                    "dielectric sphere, rough-conductor cube, cube area emitter; monochromatic 550 nm)"),
}


def measured_hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region: ONE long-running `nvidia-smi -lms 200` (B200_PROFILING.md's clocks line), started before
    and terminated after.  (Spawning a fresh nvidia-smi every 200 ms -- NVML init enumerates the whole box each time -- stalled this
    process's launches: the polling itself cost 5-40 % of a step, measured as e2e > value in profiles/r01s3_*.)"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    def __init__(self, index):
        self.index, self.proc, self.stop_flag = index, None, False
    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            import atexit
            atexit.register(self._kill)             # never leave the poller behind, whatever ends this process
        except Exception:
            self.proc = None
    def _kill(self):
        try:
            if self.proc is not None and self.proc.poll() is None: self.proc.kill()
        except Exception:
            pass
    def summary(self):
        samples, reasons, maxmhz = [], set(), None
        if self.proc is not None:
            try:
                self.proc.terminate(); out, _ = self.proc.communicate(timeout=10)
            except Exception:
                out = ""
            for line in out.splitlines():
                f = [x.strip() for x in line.split(",")]
                try:
                    samples.append(float(f[0])); maxmhz = float(f[1])
                except Exception:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"): reasons.add(name)
        s = sorted(samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": maxmhz, "reasons": sorted(reasons), "n_samples": len(s)}


def cpu_leg(built, seconds_target, threads=0):
    """Times the oracle (CPU port of the reference path) on a bounded sample: sample index 0.. of every element."""
    import _oracle
    t0 = time.time(); _, _, st = _oracle.render(built, spp=1, sample_range=(0, 1), threads=threads); t1 = time.time() - t0
    n = max(1, min(64, int(seconds_target / max(t1, 1e-3))))
    _, _, st = _oracle.render(built, spp=n, sample_range=(0, n), threads=threads)
    return st["samples"] / st["seconds"] / 1e6, st["threads"], f"samples 0..{n - 1} of every element of the {built.width}x{built.height} film ({st['samples']} samples, {st['seconds']:.1f} s)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="double_slits", choices=list(WORKLOADS))
    ap.add_argument("--integrator", default=None, choices=["plt_bdpt", "plt_path"])
    ap.add_argument("--sampler", default="uniform", choices=["uniform", "sobolld"], help="the scene sampler (sobolld: stand-in table, see wave_tracer_b200/sobol.py)")
    ap.add_argument("--spp-per-step", type=int, default=16)
    ap.add_argument("--pool", type=int, default=0, help="paths (plt_path) / sample slots (plt_bdpt) in flight; 0: 1M / 256k")
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sort", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true", help="do not record per-kernel CUDA events in the timed region (A/B of the instrumentation cost)")
    ap.add_argument("--flags", type=int, default=0, help="extra WTGPU_RENDER_* flags (A/B measurements)")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    from wave_tracer_b200 import scenes
    wl = WORKLOADS[a.workload]
    a.integrator = a.integrator or wl["integrator"] or "plt_bdpt"
    a.res = a.res or wl["res"]
    bdpt = a.integrator == "plt_bdpt"
    if a.pool == 0: a.pool = (1 << 18) if bdpt else (1 << 20)
    if a.workload == "double_slits":
        sc = scenes.double_slits(res=a.res, spp=SPP, integrator=a.integrator, lut=(2048, 1024)); integ_desc = INTEGRATORS[a.integrator]
    elif a.workload == "etoile":
        sc = scenes.etoile_like(res=a.res, spp=SPP); integ_desc = "plt_path forward + UTD FSD, max_depth 16, RR off -- the reference file's integrator"
    else:
        sc = scenes.cornell_like(res=a.res, spp=SPP, integrator=a.integrator, fsd=bdpt, lut=(2048, 1024)); integ_desc = a.integrator + ", max_depth 8"
        if a.spp_per_step == 16: a.spp_per_step = 2
    if a.sampler == "sobolld":
        from wave_tracer_b200.scene import Sobolld
        sc.sampler = Sobolld()
    built = sc.build()
    W, H = built.width, built.height
    config = {"workload": wl["name"] if a.res == wl["res"] else wl["name"].replace(str(wl["res"]), str(a.res)), "integrator": integ_desc, "film": [W, H],
              "spp_per_step": a.spp_per_step, "sampler": "philox4x32-10 counter streams keyed (seed,pixel,sample)" + ("; scene sampler sobolld (stand-in table)" if a.sampler == "sobolld" else ""),
              "l2": ("%d sample slots x 30 KB of subpath vertices/apertures" % a.pool if bdpt else "path-state pool (%d paths x 0.7 KB)" % a.pool) + " exceed the 126 MB L2; no flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return
        vals = []; step_ms = []
        for s in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            v, cores, sample = cpu_leg(built, 8.0)
            if s >= a.warmup: vals.append(v); step_ms.append((time.perf_counter() - t0) * 1e3)
        v = sum(vals) / len(vals)
        print(json.dumps({"impl": "reference", "metric": "Msamples/sec", "value": v, "unit": "Msamples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": sum(step_ms) / len(step_ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config, "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from wave_tracer_b200 import GpuScene, _abi
    from wave_tracer_b200.parallel import render_distributed
    import ctypes as C
    import numpy as np
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    gs = GpuScene(built, local)
    flags = (0 if a.no_kernel_timing else 2) | (1 if a.no_sort else 0) | a.flags      # WTGPU_RENDER_TIME_KERNELS
    S = a.spp_per_step
    # weak scaling: every rank renders S samples per element per step (disjoint sample ranges across ranks)
    def step(i):
        base = (i * world + rank) * S
        dev = torch.device("cuda", local)
        block = torch.zeros((H, W, 1, 2), dtype=torch.float32, device=dev); light = torch.zeros((H, W, 1), dtype=torch.float32, device=dev)
        st = gs.render_into(block.data_ptr(), light.data_ptr(), SPP, 0x5EED, (base, base + S), None, True, a.pool, flags, torch.cuda.current_stream().cuda_stream)
        if world > 1:
            flat = torch.cat([block.reshape(-1), light.reshape(-1)]); dist.reduce(flat, dst=0, op=dist.ReduceOp.SUM)
        return st

    clk = ClockSampler(local)
    if not os.environ.get("WT_BENCH_NO_CLOCKS"): clk.start()     # started before the warm-up so that its NVML start-up is over when the timed region begins
    for i in range(a.warmup):
        step(i)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats = [step(a.warmup + i) for i in range(a.steps)]
    e1.record()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    clocks = clk.summary()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    samples_rank = sum(s["samples"] for s in stats)
    total_samples = samples_rank * world
    value = total_samples / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel: algorithmic bytes from the device counters of the same run (DESIGN.md "Roofline model")
    peak, which = measured_hbm_peak()
    tot = lambda k: sum(s[k] for s in stats)
    trav_ms, shade_ms, conn_ms, its = tot("traverse_ms"), tot("shade_ms"), tot("connect_ms"), tot("iterations")
    if bdpt:
        # k_bd_traverse: walker record read (224 B) + hit record + key + list entry written (144+4+4 B) per walker step, 256 B per node, 48 B per triangle
        b_trav = tot("walker_steps") * (224 + 152) + 256 * tot("traverse_nodes") + 48 * tot("traverse_tris")
        # k_bd_connect<*>: task (4 B) + sample header (48 B) + the distinct 272-B vertex records a strategy reads (2,2,2,2,4 by class) + 12 B of
        # pdf/delta scalars per subpath vertex for MIS (~ 4 vertices) + shadow-ray nodes/triangles + the film/L0 atomic (8 B)
        strat = [sum(s["strategies"][c] for s in stats) for c in range(5)]
        b_conn = sum(n * (4 + 48 + u * 272 + 48 + 8) for n, u in zip(strat, (2, 2, 2, 2, 4))) + 256 * (tot("nodes_visited") - tot("traverse_nodes")) + 48 * (tot("tris_tested") - tot("traverse_tris"))
        # a "launch" of k_bd_connect = the five class launches of one iteration; of k_bd_traverse = k_bd_gtraverse + k_bd_resolve
        kern, kms, kbytes, launches_k = ("k_bd_connect", conn_ms, b_conn, its) if conn_ms >= trav_ms else ("k_bd_traverse", trav_ms, b_trav, its)
        share = {"k_bd_traverse": trav_ms / ms, "k_bd_shade(+fsd_finish)": shade_ms / ms, "k_bd_connect": conn_ms / ms}
    else:
        core_b, hit_b = 240, 160     # PathCore read + HitRec/key write per segment (16-B chunks: 15 / 10)
        kbytes = tot("segments") * (core_b + hit_b + 4) + 256 * tot("traverse_nodes") + 48 * tot("traverse_tris")
        kern, kms, launches_k = "k_traverse", trav_ms, its
        share = {"k_traverse": trav_ms / ms, "k_shade": shade_ms / ms}
    # DRAM traffic of one launch of that kernel from the committed `ncu --set full` capture of this command (profiles/ncu_traffic.json)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(a.workload, {}).get(kern)
        if tj and a.res == wl["res"]: traffic, traffic_src = tj["bytes"], tj["capture"]
    except Exception:
        pass
    achieved = kbytes / max(kms * 1e-3, 1e-12) / 1e9
    roofline = {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": which,
                "traffic": traffic, "traffic_source": traffic_src, "alg_bytes_per_launch": kbytes / max(1, launches_k), "avg_launch_ms": kms / max(1, launches_k), "share_of_step": share}

    # ---- e2e: through the public API with HOST inputs and outputs inside the timed region, at N GPUs: every step uploads the scene tables
    # (H2D), renders this rank's sample range, reduces the films to rank 0 (N>1) and reads the film back to the host (D2H).
    # N=1 is the plain C-ABI call with host film buffers (wtgpu_render does the copies); N>1 is wave_tracer_b200.parallel.
    from wave_tracer_b200 import render
    h2d = sum(C.sizeof(t) * n for t, n in ((_abi.Node, built.desc.n_nodes), (_abi.Leaf, built.desc.n_leaves), (_abi.Tri, built.desc.n_tris), (_abi.TriMeta, built.desc.n_tris),
              (_abi.TriShading, built.desc.n_tris), (_abi.Edge, built.desc.n_edges), (_abi.Shape, built.desc.n_shapes), (_abi.Spectrum, built.desc.n_spectra),
              (_abi.Bsdf, built.desc.n_bsdfs), (_abi.Emitter, built.desc.n_emitters), (_abi.KDist, built.desc.n_emitters))) + 4 * (built.desc.n_kdist_data + 1024) + \
        (8 * (built.desc.fsd_lut_n + built.desc.fsd_lut_m ** 2) if bdpt else 0)      # + the Fraunhofer sampling tables
    d2h = W * H * 3 * 4
    def e2e_step(i):
        base = (i * world + rank) * S
        if world == 1:
            return render(built, spp=SPP, device=local, sample_range=(base, base + S), pool_size=a.pool)[2]["samples"]
        g = GpuScene(built, local)
        dev = torch.device("cuda", local)
        block = torch.zeros((H, W, 1, 2), dtype=torch.float32, device=dev); light = torch.zeros((H, W, 1), dtype=torch.float32, device=dev)
        st = g.render_into(block.data_ptr(), light.data_ptr(), SPP, 0x5EED, (base, base + S), None, True, a.pool, 0, torch.cuda.current_stream().cuda_stream)
        flat = torch.cat([block.reshape(-1), light.reshape(-1)]); dist.reduce(flat, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0: flat.cpu()
        g.close()
        return st["samples"]
    e2e_step(0)     # warm
    # per-step wall time (barrier + device sync on both sides, max over ranks); the reported figure uses the MEDIAN step so that one slow
    # driver call (a cudaFree of the multi-GB pools at scene destruction was seen to take 0.7 s once in a while) does not decide it
    e2e_times, n_step = [], 0
    for i in range(max(3, min(a.steps, 5))):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        n_step = e2e_step(i)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        dt = time.time() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        e2e_times.append(dt)
    e2e_med = sorted(e2e_times)[len(e2e_times) // 2]
    e2e = n_step * world / e2e_med / 1e6

    if rank == 0:
        out = {"metric": "Msamples/sec", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
               "e2e": {"value": e2e, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "step_ms": [round(1e3 * t, 1) for t in e2e_times], "statistic": "median step"},
               "gpu_launches": int(sum(s["kernel_launches"] for s in stats)), "roofline": roofline, "clocks": clocks,
               "phases_ms_per_step": {k: tot(k) / a.steps for k in ("gpu_ms", "generate_ms", "traverse_ms", "sort_ms", "shade_ms", "connect_ms")} | {"iterations": its / a.steps},
               "counters": {k: int(sum(s[k] for s in stats)) for k in ("samples", "segments", "ray_casts", "cone_casts", "shadow_casts", "nodes_visited", "tris_tested", "splats", "capacity_overflows")}}
        if world == 1 and not a.no_cpu_baseline:
            v, cores, sample = cpu_leg(built, 12.0)
            out["cpu_baseline"] = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(out))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
