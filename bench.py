#!/usr/bin/env python
"""bench.py -- throughput of the B200 wave_tracer hot path on BASELINE.json's metric (Msamples/sec).

Workloads (--workload; `config.workload` names the one that ran).  BASELINE.json's north_star quotes the metric on cornell-box and
sionna_etoile, so the DEFAULT is the cornell box:
  cornell       configs[2]  scenes/cornell-box/box.xml -D res=1440,spp=1024: plt_bdpt max_depth 16 (RR, MIS, Fraunhofer FSD), CIE-RGB/D55 film
                1440x1440, CFL spot emitters + 7000 K area source over the visible spectrum, Au/Al/SF5/SF11 materials -- the reference file restated
                element for element (wave_tracer_b200/data/scenes/cornell_box.xml); SYNTHETIC stand-ins for its LFS-stub PLY meshes (~280k triangles)
  etoile        configs[3]  scenes/sionna_etoile/etoile.xml -D res=720,spp=1024,wavelength=10GHz: plt_path forward + UTD, max_depth 16, RR off, film
                720x540; SYNTHETIC city of 562 extruded buildings (~94k triangles) for its 563 LFS-stub PLYs
  sponza        configs[4]  scenes/sponza/sponza_day.xml -D res=1920: plt_path backward max_depth 64, CIE-RGB/D50 film 1920x1440; SYNTHETIC atrium (~277k triangles)
  double_slits  configs[1]  scenes/diffraction_simple/double_slits.xml -D res=1440,spp=1024,pattern=true: plt_bdpt + Fraunhofer FSD, film 1440x360 (10 triangles)
A "step" renders the full film at `--spp-per-step` samples per element (sample indices [step*S, step*S+S) of the scene's spp): cost is linear
in spp, samples are independent.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N>1; one rank per GPU; one NCCL film reduce per step)
  python bench.py --impl reference ...                     the CPU implementation of the same path (oracle port, all host cores; the reference
                                                           itself cannot be built here), bounded sample per step
  --scaling strong                                         N ranks split the step's samples (default: weak -- every rank renders S per element)
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    "cornell": dict(res=1440, spp=1024, spp_per_step=2, pool=1 << 20,
                    name="cornell-box/box res=1440 spp=1024 visible spectrum (film 1440x1440 RGB CIE/D55; plt_bdpt max_depth 16, RR, MIS, Fraunhofer FSD; box.xml restated element "
                         "for element; SYNTHETIC stand-ins for its three LFS-stub PLY meshes and one PNG: config.standins)"),
    "etoile": dict(res=720, spp=1024, spp_per_step=16, pool=1 << 20,
                   name="sionna_etoile/etoile res=720 spp=1024 wavelength=10GHz (film 720x540; plt_path forward + UTD, max_depth 16, RR off; SYNTHETIC geometry: ground plane + "
                        "562 extruded polygonal buildings of 6 storeys on a seeded street plan, the file's ITU materials, emitter and sensor)"),
    "sponza": dict(res=1920, spp=4096, spp_per_step=1, pool=1 << 20,
                   name="sponza/sponza_day res=1920 spp=4096 visible spectrum (film 1920x1440 RGB CIE/D50; plt_path backward max_depth 64, RR; the file's sun, LED bulb and sky panel; "
                        "SYNTHETIC procedural atrium for the LFS-stub OBJ)"),
    "double_slits": dict(res=1440, spp=1024, spp_per_step=16, pool=1 << 18,
                         name="diffraction_simple/double_slits res=1440 spp=1024 pattern=true (film 1440x360, lambda=0.05mm; plt_bdpt max_depth 16, Fraunhofer FSD; procedural restatement)"),
}
PARITY_RES = {"cornell": (96, 4), "etoile": (96, 4), "sponza": (96, 2), "double_slits": (256, 8)}      # (res, spp) of the parity render of the same workload


def make_scene(workload, res, spp, integrator=None, small=False):
    """-> (Scene, integrator description, table_size)."""
    from wave_tracer_b200 import scenes
    if workload == "cornell":
        sc = scenes.cornell_box(res=res, spp=spp, lut=(512, 256) if small else (2048, 1024), dragon_tris=11520 if small else 184320, bunny_tris=5120 if small else 81920)
        return sc, "plt_bdpt (MIS, emitter+sensor direct, RR), Fraunhofer FSD, max_depth 16 -- box.xml:8-13", 1024
    if workload == "etoile":
        return scenes.etoile_like(res=res, spp=spp, detail=6), "plt_path forward + UTD FSD, max_depth 16, RR off -- etoile.xml:24-29", 256
    if workload == "sponza":
        return scenes.sponza_like(res=res, spp=spp, detail=.5 if small else 1.0), "plt_path backward + UTD FSD, max_depth 64, RR -- sponza_day.xml:8-11", 512
    integrator = integrator or "plt_bdpt"
    sc = scenes.double_slits(res=res, spp=spp, integrator=integrator, lut=(512, 256) if small else (2048, 1024))
    return sc, {"plt_bdpt": "plt_bdpt (MIS, emitter+sensor direct), Fraunhofer FSD, max_depth 16 -- double_slits.xml:42-44",
                "plt_path": "plt_path forward + UTD FSD, max_depth 16, RR off"}[integrator], 256


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
        self.index, self.proc = index, None
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
    """Times the oracle (CPU port of the reference path, liboracle.so: -O3 -march=x86-64-v3) on a bounded sample of the workload: a centred tile
    of the film at 1+ samples per element, sized to `seconds_target` seconds from a probe."""
    import _oracle
    W, H = built.width, built.height
    def run(tile, n):
        _, _, st = _oracle.render(built, spp=built.spp, sample_range=(0, n), tile=tile, threads=threads)
        return st
    probe_w = max(16, min(W, 64)); probe_h = max(16, min(H, 64))
    x0, y0 = (W - probe_w) // 2, (H - probe_h) // 2
    st = run((x0, y0, x0 + probe_w, y0 + probe_h), 1)
    rate = st["samples"] / max(st["seconds"], 1e-4)
    want = max(probe_w * probe_h, int(rate * seconds_target))
    n = 1
    tw, th = W, H
    if want < W * H:        # a centred tile with the film's aspect ratio
        f = (want / (W * H)) ** .5; tw, th = max(16, int(W * f)), max(16, int(H * f))
    else:
        n = max(1, min(64, want // (W * H)))
    x0, y0 = (W - tw) // 2, (H - th) // 2
    st = run((x0, y0, x0 + tw, y0 + th), n)
    return st["samples"] / st["seconds"] / 1e6, st["threads"], (f"samples 0..{n - 1} of the centred {tw}x{th} tile of the {W}x{H} film ({st['samples']} samples, {st['seconds']:.1f} s)")


def parity_record(workload, integrator, device):
    """Per-pixel parity of THIS build against the oracle on the same workload at reduced size (reported with every throughput)."""
    import numpy as np
    import _oracle
    from wave_tracer_b200 import render, develop
    res, spp = PARITY_RES[workload]
    sc, _, tsz = make_scene(workload, res, spp, integrator, small=True)
    b = sc.build(table_size=tsz)
    blk, lgt, st = render(b, spp=spp, device=device)
    oblk, olgt, ost = _oracle.render(b, spp=spp)
    g = develop(b, spp, blk, lgt); o = develop(b, spp, oblk, olgt)
    lit = np.abs(o) > 0
    return {"rel_l2": float(np.linalg.norm(g - o) / max(np.linalg.norm(o), 1e-300)), "mean_rel": float((np.abs(g - o)[lit] / np.abs(o[lit])).mean()) if lit.any() else 0.0,
            "vs": "CPU oracle (f64 film), equal seeds", "film": [b.width, b.height, b.channels], "spp": spp, "triangles": int(b.desc.n_tris),
            "counters_equal": bool(st["segments"] == ost["segments"] and st["samples"] == ost["samples"]), "capacity_overflows": int(st["capacity_overflows"]),
            "gate": {"rel_l2": 1e-3, "mean_rel": 2e-4}}


def short_record(workload, device, steps=2, warmup=2):
    """A short device-timed run of another BASELINE workload (its own film size, pool and spp per step) with its parity record: embedded in the
    default line as `other_workloads` so that one bench run states all four north_star / BASELINE configurations."""
    import torch
    from wave_tracer_b200 import GpuScene
    wl = WORKLOADS[workload]
    sc, integ, tsz = make_scene(workload, wl["res"], wl["spp"])
    built = sc.build(table_size=tsz)
    W, H, Cn = built.width, built.height, built.channels
    gs = GpuScene(built, device)
    dev = torch.device("cuda", device)
    S = wl["spp_per_step"]
    def step(i):
        block = torch.zeros((H, W, Cn, 2), dtype=torch.float32, device=dev); light = torch.zeros((H, W, Cn), dtype=torch.float32, device=dev)
        return gs.render_into(block.data_ptr(), light.data_ptr(), wl["spp"], 0x5EED, (i * S, i * S + S), None, True, wl["pool"], 0, torch.cuda.current_stream().cuda_stream)
    for i in range(warmup): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); stats = [step(warmup + i) for i in range(steps)]; e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    gs.close()
    n = sum(s["samples"] for s in stats)
    return {"workload": wl["name"], "integrator": integ, "film": [W, H, Cn], "triangles": int(built.desc.n_tris), "spp_per_step": S, "steps": steps, "warmup": warmup,
            "value": n / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms / steps, "capacity_overflows": int(sum(s["capacity_overflows"] for s in stats)),
            "parity": parity_record(workload, None, device)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cornell", choices=list(WORKLOADS))
    ap.add_argument("--integrator", default=None, choices=["plt_bdpt", "plt_path"], help="double_slits only")
    ap.add_argument("--sampler", default="uniform", choices=["uniform", "sobolld"], help="the scene sampler (sobolld: stand-in table, see wave_tracer_b200/sobol.py)")
    ap.add_argument("--spp-per-step", type=int, default=0)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--pool", type=int, default=0, help="paths (plt_path) / sample slots (plt_bdpt) in flight; 0: the workload's default")
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="N=1: do not append the short records of the other BASELINE workloads (config.other_workloads)")
    ap.add_argument("--no-sort", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true", help="do not record per-kernel CUDA events in the timed region (A/B of the instrumentation cost)")
    ap.add_argument("--flags", type=int, default=0, help="extra WTGPU_RENDER_* flags (A/B measurements)")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[a.workload]
    a.res = a.res or wl["res"]
    SPP = wl["spp"]
    S = a.spp_per_step or wl["spp_per_step"]
    a.pool = a.pool or wl["pool"]
    sc, integ_desc, tsz = make_scene(a.workload, a.res, SPP, a.integrator)
    bdpt = type(sc.integrator).__name__ == "PltBdpt"
    if a.sampler == "sobolld":
        from wave_tracer_b200.scene import Sobolld
        sc.sampler = Sobolld()
    built = sc.build(table_size=tsz)
    W, H, Cn = built.width, built.height, built.channels
    config = {"workload": wl["name"] if a.res == wl["res"] else wl["name"].replace(str(wl["res"]), str(a.res)), "integrator": integ_desc, "film": [W, H, Cn],
              "triangles": int(built.desc.n_tris), "edges": int(built.desc.n_edges), "bvh_nodes": int(built.desc.n_nodes),
              "standins": [f"{k}: {f} -> {w}" for k, f, w in getattr(sc, "standins", [])],
              "spp_per_step": S, "sampler": "philox4x32-10 counter streams keyed (seed,pixel,sample)" + ("; scene sampler sobolld (stand-in table)" if a.sampler == "sobolld" else ""),
              "l2": ("%d sample slots of subpath vertices/apertures (GBs)" % a.pool if bdpt else "path-state pool (%d paths x 0.7 KB)" % a.pool) + " exceed the 126 MB L2; no flush needed"}

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
                          "ms_per_step": sum(step_ms) / len(step_ms), "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config, "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample,
                                                             "build": "oracle/liboracle.so, g++ -O3 -march=x86-64-v3 -ffp-contract=off"},
                          "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from wave_tracer_b200 import GpuScene, _abi
    import ctypes as C
    import numpy as np
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    gs = GpuScene(built, local)
    # The timed region runs the product configuration (sub-pools on their own streams).  WTGPU_RENDER_TIME_KERNELS (per-kernel CUDA events for the
    # roofline) serialises the render into one sub-pool -- measured 20 % slower on cornell -- so it is used in an INSTRUMENTED REPEAT of the first
    # timed steps right after the timed region, never inside it.
    flags = (1 if a.no_sort else 0) | a.flags
    flags_inst = flags | 2
    # weak scaling: every rank renders S samples per element per step (disjoint sample ranges across ranks); strong: the S samples of a step are split
    def sample_range(i):
        if a.scaling == "weak":
            base = (i * world + rank) * S
            return base, base + S
        from wave_tracer_b200.parallel import partition_samples
        b0, b1 = partition_samples(S, rank, world)
        return i * S + b0, i * S + b1
    dev = torch.device("cuda", local)
    def step(i, flags=flags):
        s0, s1 = sample_range(i)
        block = torch.zeros((H, W, Cn, 2), dtype=torch.float32, device=dev); light = torch.zeros((H, W, Cn), dtype=torch.float32, device=dev)
        st = gs.render_into(block.data_ptr(), light.data_ptr(), SPP, 0x5EED, (s0, s1), None, True, a.pool, flags, torch.cuda.current_stream().cuda_stream) if s1 > s0 else None
        if world > 1:
            flat = torch.cat([block.reshape(-1), light.reshape(-1)]); dist.reduce(flat, dst=0, op=dist.ReduceOp.SUM)
        return st

    clk = ClockSampler(local)
    if not os.environ.get("WT_BENCH_NO_CLOCKS"): clk.start()     # started before the warm-up so that its NVML start-up is over when the timed region begins
    for i in range(a.warmup):
        step(i)         # (the first step also settles the list capacities: wtgpu_stats::passes > 1 only here)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats = [st for st in (step(a.warmup + i) for i in range(a.steps)) if st is not None]
    e1.record()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    clocks = clk.summary()
    ms = e0.elapsed_time(e1)
    samples_rank = sum(s["samples"] for s in stats)
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        n = torch.tensor([samples_rank], device="cuda", dtype=torch.float64); dist.all_reduce(n, op=dist.ReduceOp.SUM); total_samples = float(n.item())
    else:
        total_samples = samples_rank
    value = total_samples / (ms * 1e-3) / 1e6
    timed_stats = stats
    # instrumented repeat of the first timed steps (same sample ranges): per-kernel device times and the counters the roofline's bytes come from
    n_inst = 0 if a.no_kernel_timing else min(a.steps, 2)
    if n_inst:
        step(a.warmup, flags_inst)      # untimed: the pools are re-cut for one sub-pool here
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        stats = [st for st in (step(a.warmup + i, flags_inst) for i in range(n_inst)) if st is not None]
        i1.record(); torch.cuda.synchronize()
        ms_inst = i0.elapsed_time(i1)
    else:
        ms_inst = ms

    # ---- roofline of the dominant kernel: algorithmic bytes from the device counters of the same run (DESIGN.md "Roofline model")
    peak, which = measured_hbm_peak()
    tot = lambda k: sum(s[k] for s in stats)
    trav_ms, shade_ms, conn_ms, its = tot("traverse_ms"), tot("shade_ms"), tot("connect_ms"), tot("iterations")
    if bdpt:
        # k_bd_traverse: walker record read (224 B) + traversal record + hit record + key + list entry written (48+48+4+4 B) per walker step, 256 B per node,
        # 48 B per triangle test, 4 B per returned triangle id (written, then read by the resolve kernel)
        b_trav = tot("walker_steps") * (224 + 104) + 256 * tot("traverse_nodes") + 48 * tot("traverse_tris")
        # k_bd_connect<*>: task (8 B) + sample header (48 B) + the distinct 272-B vertex records a strategy reads (2,2,2,2,4 by class) + 12 B of
        # pdf/delta scalars per subpath vertex for MIS (~ 4 vertices) + shadow-ray nodes/triangles + the film/L0 atomic (8 B)
        strat = [sum(s["strategies"][c] for s in stats) for c in range(5)]
        b_conn = sum(n * (8 + 48 + u * 272 + 48 + 8) for n, u in zip(strat, (2, 2, 2, 2, 4))) + 256 * (tot("nodes_visited") - tot("traverse_nodes")) + 48 * (tot("tris_tested") - tot("traverse_tris"))
        # a "launch" of k_bd_connect = the five class launches of one iteration; of k_bd_traverse = k_bd_gtraverse + k_bd_resolve
        kern, kms, kbytes, launches_k = ("k_bd_connect", conn_ms, b_conn, its) if conn_ms >= trav_ms else ("k_bd_traverse", trav_ms, b_trav, its)
        share = {"k_bd_traverse": trav_ms / ms_inst, "k_bd_shade(+fsd_finish)": shade_ms / ms_inst, "k_bd_connect": conn_ms / ms_inst}
    else:
        core_b, hit_b = 240, 48 + 48     # PathCore read + traversal record / HitRec write per segment
        kbytes = tot("segments") * (core_b + hit_b + 4) + 256 * tot("traverse_nodes") + 48 * tot("traverse_tris")
        kern, kms, launches_k = "k_traverse", trav_ms, its
        share = {"k_traverse": trav_ms / ms_inst, "k_shade": shade_ms / ms_inst}
    # DRAM traffic of one launch of that kernel from the committed `ncu --set full` capture of this command (profiles/ncu_traffic.json)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(a.workload, {}).get(kern)
        if tj and a.res == wl["res"] and tj.get("spp_per_step") == S: traffic, traffic_src = tj["bytes"], tj["capture"]
    except Exception:
        pass
    achieved = kbytes / max(kms * 1e-3, 1e-12) / 1e9
    roofline = {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": which,
                "traffic": traffic, "traffic_source": traffic_src, "alg_bytes_per_launch": kbytes / max(1, launches_k), "avg_launch_ms": kms / max(1, launches_k), "share_of_step": share,
                "measured_in": "instrumented repeat of the first %d timed steps (WTGPU_RENDER_TIME_KERNELS: one sub-pool, per-kernel CUDA events on its stream), %.1f ms per step" % (n_inst, ms_inst / max(1, n_inst))}

    # ---- e2e: through the public API with HOST inputs and outputs inside the timed region, at N GPUs: every step uploads the scene tables
    # (H2D), renders this rank's sample range, reduces the films to rank 0 (N>1) and reads the film back to the host (D2H).
    # N=1 is the plain C-ABI call with host film buffers (wtgpu_render does the copies); N>1 is wave_tracer_b200.parallel's scheme.
    from wave_tracer_b200 import render
    h2d = sum(C.sizeof(t) * n for t, n in ((_abi.Node, built.desc.n_nodes), (_abi.Leaf, built.desc.n_leaves), (_abi.Tri, built.desc.n_tris), (_abi.TriMeta, built.desc.n_tris),
              (_abi.TriShading, built.desc.n_tris), (_abi.Edge, built.desc.n_edges), (_abi.Shape, built.desc.n_shapes), (_abi.Spectrum, built.desc.n_spectra),
              (_abi.Bsdf, built.desc.n_bsdfs), (_abi.Emitter, built.desc.n_emitters), (_abi.KDist, built.desc.n_emitters))) + 4 * (built.desc.n_kdist_data + built.desc.n_spectrum_data + 1024) + \
        (8 * (built.desc.fsd_lut_n + built.desc.fsd_lut_m ** 2) if bdpt else 0)      # + the Fraunhofer sampling tables
    d2h = W * H * Cn * 3 * 4
    caps = gs.capacities()
    def e2e_step(i):
        s0, s1 = sample_range(i)
        g = GpuScene(built, local); g.set_capacities(caps)      # a fresh handle per step (scene upload inside the timed region) that knows the scene's list lengths
        if world == 1:
            n = render(built, spp=SPP, device=local, sample_range=(s0, s1), pool_size=a.pool, gpu_scene=g)[2]["samples"]
            g.close(); return n
        block = torch.zeros((H, W, Cn, 2), dtype=torch.float32, device=dev); light = torch.zeros((H, W, Cn), dtype=torch.float32, device=dev)
        st = g.render_into(block.data_ptr(), light.data_ptr(), SPP, 0x5EED, (s0, s1), None, True, a.pool, 0, torch.cuda.current_stream().cuda_stream) if s1 > s0 else {"samples": 0}
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
            n = torch.tensor([float(n_step)], device="cuda", dtype=torch.float64); dist.all_reduce(n, op=dist.ReduceOp.SUM); n_step = float(n.item())
        e2e_times.append(dt)
    e2e_med = sorted(e2e_times)[len(e2e_times) // 2]
    e2e = n_step / e2e_med / 1e6

    if rank == 0:
        out = {"metric": "Msamples/sec", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
               "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
               "e2e": {"value": e2e, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "step_ms": [round(1e3 * t, 1) for t in e2e_times], "statistic": "median step"},
               "gpu_launches": int(sum(s["kernel_launches"] for s in timed_stats)), "roofline": roofline, "clocks": clocks,
               "phases_ms_per_step": {k: tot(k) / max(1, len(stats)) for k in ("gpu_ms", "generate_ms", "traverse_ms", "sort_ms", "shade_ms", "connect_ms")} | {"iterations": its / max(1, len(stats)), "of": "the instrumented repeat"},
               "counters": {k: int(sum(s[k] for s in timed_stats)) for k in ("samples", "segments", "ray_casts", "cone_casts", "shadow_casts", "nodes_visited", "tris_tested", "splats", "capacity_overflows", "stack_drops")},
               "capacities": dict(zip(("cone_tris", "edges", "fraunhofer_segments", "apertures_per_subpath", "vertices_per_subpath"), gs.capacities())) | {"passes_in_timed_steps": max(s["passes"] for s in timed_stats), "pool_used": timed_stats[-1]["pool_used"], "subpools": timed_stats[-1].get("subpools")}}
        if world == 1 and not a.no_parity:
            out["parity"] = parity_record(a.workload, a.integrator, local)
        if world == 1 and not a.no_other_workloads and a.res == wl["res"]:
            gs.close()
            out["other_workloads"] = [short_record(w, local) for w in WORKLOADS if w != a.workload]
        if world == 1 and not a.no_cpu_baseline:
            v, cores, sample = cpu_leg(built, 12.0)
            out["cpu_baseline"] = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample, "build": "oracle/liboracle.so, g++ -O3 -march=x86-64-v3 -ffp-contract=off"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
