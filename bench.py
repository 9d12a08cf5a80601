#!/usr/bin/env python3
"""bench.py -- headline benchmark of the XPBD tet-FEM substep path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE config 4): synthetic Kuhn-split beam, 407 x 64 x 64 cells = 10,002,432 tets,
1,723,800 vertices, Neo-Hookean Jacobi, 1 iteration per substep, dt = 1/1200.
One bench STEP = one 60 Hz frame = 20 substeps (the reference's GPU default numSubsteps,
src/main.js:26), issued as ONE tetsim_step call (one CUDA-graph launch).
Metric: tet-constraint projections per second = tets x iterations x substeps / time, in Mtet/s
(one projection = one solveElem = 2 scalar XPBD constraints).

  value      device-resident throughput (state stays in HBM), CUDA events, max over ranks
  e2e        same metric through the public SoftBody API with HOST buffers: per step the caller's
             pos/prevPos/vel are uploaded from pinned host memory, the frame is solved, positions
             are read back -- copies inside the timed region
  roofline   clustered Jacobi tile kernel: algorithmic bytes (56 B/tet + 32 B/vertex) / mean launch
             time (CUDA events, tetsim_time_kernel) vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own algorithm (src/Softbody.js sequential Gauss-Seidel, restated in C:
             oracle/softbody_oracle.c, 1 thread -- the sweep is inherently sequential) on a bounded
             sample of the same mesh, on this box's host cores
--impl reference runs only that CPU arm (the reference is JavaScript; no JS engine exists in this
image, so the C restatement is the reference arm -- kind "port").
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

FRAME_DT = 1.0 * (1.0 / 60.0)   # timeScale * timeStep, src/main.js:79
METRIC = "tet_constraint_projections_per_s"
UNIT = "Mtet/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", default="407,64,64", help="beam cells x,y,z (default = 10,002,432 tets)")
    ap.add_argument("--jitter", type=float, default=0.0, help="displace the beam's interior vertices by +-jitter*h (numpy default_rng(1234)); "
                    "0 = the regular Kuhn grid of BASELINE config 4")
    ap.add_argument("--substeps", type=int, default=20, help="substeps per step (frame)")
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--cluster-size", type=int, default=0, help="tets per tile (default 512 for the NH tile kernel -- measured fastest, profiles/r2_tile_experiments.txt -- and 128 for the polar tile kernel)")
    ap.add_argument("--no-reorder", action="store_true")
    ap.add_argument("--atomic", action="store_true", help="deterministic=0: REDG flush instead of per-tile partials")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the same 10M-tet mesh split over N GPUs (BASELINE config 4); weak: beam length x N")
    ap.add_argument("--cpu-substeps", type=int, default=6, help="substeps of the CPU baseline sample (rank 0, N=1)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "allreduce", "halo", "peer"],
                    help="multi-GPU boundary exchange: ncclAllReduce over all ranks, grouped ncclSend/ncclRecv with the neighbour ranks, or the "
                         "fused peer-memory exchange (the tile kernel stores its boundary partials straight into the sharers' buffers over "
                         "NVLink); auto = peer (measured fastest, profiles/r2_peer_experiments.txt), falling back to halo without peer access")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank-vs-1-rank check at N > 1")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config rates (BASELINE configs 1, 2, 3, 5) at N = 1")
    ap.add_argument("--workload", default="beam", choices=["beam", "polar"],
                    help="beam: BASELINE config 4 (NH Jacobi tile kernel, the headline); polar: the SoftBodyGPU polar-decomposition "
                         "Jacobi (BASELINE config 2's algorithm) on the same 10M-tet beam, tiled kernels, roofline 148 B/tet + 32 B/vertex")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained leg")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < 3.0:   # at least one sample
            time.sleep(0.05)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(verts, tets, substeps, dt, polar=False):
    """The reference's CPU algorithm (sequential Gauss-Seidel, src/Softbody.js:195-240) via the C restatement; for the polar
    workload the C restatement of the WebGL passes (src/SoftbodyGPU.js:59-376)."""
    import oracle
    ref = oracle.PolarOracle(verts, tets) if polar else oracle.SoftBodyOracle(verts, tets)
    t0 = time.perf_counter()
    for _ in range(substeps):
        ref.simulate(dt)
    sec = time.perf_counter() - t0
    return (tets.size // 4) * substeps / sec / 1e6, sec


def cpu_dragon_substeps_per_s(substeps=300):
    """BASELINE config 1, the metric's "reference CPU substeps/s": the Dragon mesh, Neo-Hookean Gauss-Seidel in the
    reference's order at dt = 1/600 (10 substeps per frame), C restatement of src/Softbody.js, 1 thread."""
    import oracle
    from tetsim_b200 import mesh
    m = mesh.load_dragon()
    ref = oracle.SoftBodyOracle(m["tet_verts"], m["tet_ids"])
    for _ in range(10):
        ref.simulate(FRAME_DT / 10)
    t0 = time.perf_counter()
    for _ in range(substeps):
        ref.simulate(FRAME_DT / 10)
    return substeps / (time.perf_counter() - t0)


def vec_rel(x, ref):
    x, r = np.asarray(x, np.float64).reshape(-1, 3), np.asarray(ref, np.float64).reshape(-1, 3)
    return float(np.max(np.linalg.norm(x - r, axis=1) / np.maximum(np.linalg.norm(r, axis=1), 1e-30)))


def jacobi_vs_gs(stream, substeps=100, iters_list=(1, 2, 4, 8, 16)):
    """How far NH-Jacobi(iters) is from the reference's answer: Dragon, 100 substeps at dt = 1/600 (BASELINE config 1),
    vector-relative position error against the reference-order Gauss-Seidel run by this library in BITEXACT arithmetic
    (bit-identical to src/Softbody.js, tests/test_parity_gpu.py), and the tet-projection rate divided by iters
    ("equal-substep" rate).  The reference never runs Jacobi Neo-Hookean (README.md:25)."""
    import torch
    import tetsim_b200 as ts
    from tetsim_b200 import mesh
    m = mesh.load_dragon()
    dt = FRAME_DT / 10
    pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=10)
    gs = ts.SoftBody(m["tet_verts"], m["tet_ids"], None, pp, solver="gs_exact", arithmetic="bitexact", stream=stream.cuda_stream)
    for _ in range(substeps // 10):
        gs.step(pp)
    ref = gs.pos.copy()
    gs.close()
    y0 = m["tet_verts"].reshape(-1, 3)[:, 1]
    out = {"reference": "NH Gauss-Seidel in tet order, BITEXACT (== src/Softbody.js), Dragon, %d substeps at dt=1/600" % substeps,
           "free_fall_drop_m": float(np.mean(y0) - np.mean(ref.reshape(-1, 3)[:, 1])), "by_iters": {}}
    for it in iters_list:
        jb = ts.SoftBody(m["tet_verts"], m["tet_ids"], None, pp, solver="jacobi", arithmetic="fast", iters=it, stream=stream.cuda_stream)
        for _ in range(substeps // 10):
            jb.step(pp)
        x = jb.pos.copy()
        # shape error with the rigid translation removed: how different the deformed shape is, in units of the body size
        a, b = x.reshape(-1, 3).astype(np.float64), ref.reshape(-1, 3).astype(np.float64)
        shape = float(np.max(np.linalg.norm((a - a.mean(0)) - (b - b.mean(0)), axis=1)) / np.max(np.ptp(b, axis=0)))
        out["by_iters"][str(it)] = {"vec_rel_err": vec_rel(x, ref), "shape_err_rel_body_size": shape}
        jb.close()
    return out


def config_rates(stream, device_index, frames=20):
    """Driver-visible rates of the BASELINE configs the headline does not cover (1, 2, 3(i), 3(ii), 5), one B200, each with
    its own clock sample: substeps/s and tet-projections/s through tetsim_step (one CUDA-graph launch per frame)."""
    import torch
    import tetsim_b200 as ts
    from tetsim_b200 import mesh
    m = mesh.load_dragon()
    out = {}

    def rate(key, what, make, pp, nframes):
        body = make()
        for _ in range(3):
            body.step(pp)
        body.synchronize()
        smp = ClockSampler(device_index)
        smp.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(nframes):
            body.step(pp)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        info = body.info()
        sub = nframes * pp["numSubsteps"]
        out[key] = {"what": what, "substeps_per_s": sub / ms * 1e3, "Mtet_per_s": info["numTets"] * info["iters"] * sub / ms / 1e3,
                    "tets": info["numTets"], "launches_per_substep": info["launchesPerSubstep"], "ms_timed": ms,
                    "finite": bool(np.isfinite(body.pos).all()), "clocks": smp.stop()}
        body.close()

    p10 = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=10)
    p20 = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=20)
    sk = dict(stream=stream.cuda_stream)
    V, T = m["tet_verts"], m["tet_ids"]
    rate("C1_C3i_gs_exact_bitexact", "Dragon, NH Gauss-Seidel in the reference order, the reference's arithmetic (bit-identical to src/Softbody.js)",
         lambda: ts.SoftBody(V, T, None, p10, solver="gs_exact", arithmetic="bitexact", **sk), p10, frames)
    rate("C1_C3i_gs_exact_fast", "Dragon, NH Gauss-Seidel in the reference order, f32",
         lambda: ts.SoftBody(V, T, None, p10, solver="gs_exact", arithmetic="fast", **sk), p10, frames)
    rate("C3ii_gs_color_fast", "Dragon, NH Gauss-Seidel by greedy graph colouring (32 colours), f32",
         lambda: ts.SoftBody(V, T, None, p10, solver="gs_color", arithmetic="fast", **sk), p10, frames)
    rate("C2_polar_fast", "Dragon, polar-decomposition Jacobi (SoftBodyGPU), 20 substeps/frame, f32",
         lambda: ts.SoftBodyGPU(V, T, None, dict(p20), **sk), p20, frames)
    rate("C2_polar_bitexact", "Dragon, polar-decomposition Jacobi (SoftBodyGPU), 20 substeps/frame, the shader's arithmetic",
         lambda: ts.SoftBodyGPU(V, T, None, dict(p20), arithmetic="bitexact", **sk), p20, frames)
    wb = list(mesh.wide_bounds(64.0))
    for n in (8, 28):
        v, t = mesh.tile_bodies(V, T, n, n, y_shift=-0.40)
        pw = dict(p10, worldBounds=wb)
        rate("C5_%dx_gs_exact_fast" % (n * n), "%d tiled Dragons + ground, NH Gauss-Seidel in the reference order per body, f32" % (n * n),
             lambda: ts.SoftBody(v, t, None, pw, solver="gs_exact", arithmetic="fast", **sk), pw, max(3, frames // 2))
        pj = dict(p20, worldBounds=wb)
        rate("C5_%dx_jacobi_tiles" % (n * n), "%d tiled Dragons + ground, NH Jacobi tile kernel (unstructured mesh), 20 substeps/frame" % (n * n),
             lambda: ts.SoftBody(v, t, None, pj, solver="jacobi", arithmetic="fast", cluster_size=512, **sk), pj, frames)
    # BASELINE config 4 on an UNSTRUCTURED-ish mesh: the headline beam is a regular Kuhn grid (the best case for tile locality);
    # the same 10,002,432 tets with every interior vertex displaced by +-0.2 h
    vj, tj = mesh.make_beam((407, 64, 64), jitter=0.2)
    pj = dict(p20, worldBounds=wb)
    bj = ts.SoftBody(vj, tj, None, pj, solver="jacobi", arithmetic="fast", cluster_size=512, **sk)
    k_ms, k_bytes = bj.time_kernel(30)
    peak, _ = load_peaks()
    rate("C4_beam_jitter02", "10,002,432-tet beam, interior vertices jittered +-0.2 h, NH Jacobi tile kernel, 20 substeps/frame", lambda: bj, pj, 20)
    out["C4_beam_jitter02"]["tile_kernel_ms"] = k_ms
    out["C4_beam_jitter02"]["tile_kernel_frac_of_hbm_peak"] = k_bytes / (k_ms * 1e-3) / 1e9 / peak
    out["jacobi_vs_gs"] = jacobi_vs_gs(stream)
    return out


def c5_sharded(stream, rank, world, local_rank, dist, n=28, frames=10):
    """BASELINE config 5 on N GPUs: n*n tiled Dragons + ground, Neo-Hookean Gauss-Seidel in the reference order per body,
    bodies sharded across the ranks (tetsim_b200.mesh.shard_bodies: independent bodies, NO exchange).  Rate = all bodies'
    tets x substeps / max-over-ranks time.  At N > 1 rank 0 also runs the whole scene on its own GPU and the merged
    shard results must be BIT-identical to it (GS per body is deterministic and bodies do not interact)."""
    import torch
    import tetsim_b200 as ts
    from tetsim_b200 import mesh
    m = mesh.load_dragon()
    v, t = mesh.tile_bodies(m["tet_verts"], m["tet_ids"], n, n, y_shift=-0.40)
    pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=10, worldBounds=list(mesh.wide_bounds(64.0)))
    sv, st, vid, _ = mesh.shard_bodies(v, t, rank, world)
    body = ts.SoftBody(sv, st, None, pp, solver="gs_exact", arithmetic="fast", device=local_rank, stream=stream.cuda_stream)
    for _ in range(2):
        body.step(pp)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    smp = ClockSampler(local_rank)
    smp.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(frames):
        body.step(pp)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = smp.stop()
    sub = frames * pp["numSubsteps"]
    M = t.size // 4
    out = {"what": "%d tiled Dragons + ground, NH Gauss-Seidel in the reference order per body (f32), bodies sharded over %d GPU(s), no exchange" % (n * n, world),
           "bodies": n * n, "tets": M, "substeps_per_s": sub / ms * 1e3, "Mtet_per_s": M * sub / ms / 1e3, "ms_timed": ms,
           "bodies_rank0": int(body.info()["numComponents"]), "clocks": clocks}
    if world > 1:
        full = torch.full((v.size,), float("nan"), dtype=torch.float32)
        full.view(-1, 3)[torch.from_numpy(vid.astype(np.int64))] = torch.from_numpy(body.pos.copy()).view(-1, 3)
        g = full.cuda()
        parts = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts, g)
        if rank == 0:
            P = torch.stack(parts).cpu().numpy()
            owned = np.isfinite(P)
            merged = np.where(owned, P, 0.0).sum(axis=0).astype(np.float32)
            single = ts.SoftBody(v, t, None, pp, solver="gs_exact", arithmetic="fast", device=local_rank, stream=stream.cuda_stream)
            for _ in range(2 + frames):
                single.step(pp)
            ref = single.pos.copy()
            single.close()
            out["parity"] = {"each_vertex_on_exactly_one_rank": bool((owned.sum(axis=0) == 1).all()),
                             "bit_identical_to_1_gpu": bool(np.array_equal(merged.view(np.uint32), ref.view(np.uint32))),
                             "substeps": (2 + frames) * pp["numSubsteps"]}
        dist.barrier()
    body.close()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.exchange == "auto":
        args.exchange = "peer"
    polar = args.workload == "polar"
    if not args.cluster_size:
        args.cluster_size = 128 if polar else 512
    if polar:
        args.no_configs = True
    cells = tuple(int(c) for c in args.cells.split(","))
    if args.scaling == "weak":
        cells = (cells[0] * max(world, 1), cells[1], cells[2])
    dt = FRAME_DT / args.substeps
    from tetsim_b200 import mesh

    mesh_name = "beam %dx%dx%d cells Kuhn-split" % cells + (", interior vertices jittered +-%.2f h" % args.jitter if args.jitter else "")
    workload = "%s, NH Jacobi iters=%d, dt=1/%d, %d substeps/step" % (mesh_name, args.iters, round(1.0 / dt), args.substeps)
    if polar:
        workload = "%s, polar-decomposition shape-matching Jacobi (SoftBodyGPU, src/SoftbodyGPU.js:59-376), dt=1/%d, %d substeps/step" % (
            mesh_name, round(1.0 / dt), args.substeps)

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample: one substep per bench step, of the full mesh when the whole run then fits ~2 minutes of host time,
        # else of a beam shortened along x (same cross-section, same tet shapes; the rate is per tet)
        rcells = cells
        budget_tets = 2.5e6 * 120.0 / max(args.steps + args.warmup, 1)      # at the slowest host rate seen (~2.5 M tet/s)
        full_tets = 6.0 * cells[0] * cells[1] * cells[2]
        if budget_tets < full_tets:
            rcells = (max(16, int(cells[0] * budget_tets / full_tets)), cells[1], cells[2])
        verts, tets = mesh.make_beam(rcells, jitter=args.jitter)
        M = tets.size // 4
        import oracle
        ref = oracle.PolarOracle(verts, tets) if polar else oracle.SoftBodyOracle(verts, tets)
        per_step = 1
        for _ in range(args.warmup):
            ref.simulate(dt)
        t0 = time.perf_counter()
        for _ in range(args.steps * per_step):
            ref.simulate(dt)
        sec = time.perf_counter() - t0
        val = M * args.steps * per_step / sec / 1e6
        sample = "%d substep(s) of %s (%d tets) per step; sequential Gauss-Seidel, C restatement of src/Softbody.js (no JS engine in image)" % (
            per_step, "the full mesh" if rcells == cells else "a %dx%dx%d-cell section of the %dx%dx%d beam" % (rcells + cells), M)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64-expr/f32-store",
                "data": "synthetic",
                "config": {"workload": ("%s, polar-decomposition Jacobi (the 7 passes of src/SoftbodyGPU.js restated in C, f32), dt=1/%d, 1 substep/step"
                                        if polar else
                                        "%s, NH sequential Gauss-Seidel in tet order (the reference's own algorithm, src/Softbody.js:206-208), "
                                        "dt=1/%d, 1 substep/step") % (mesh_name, round(1.0 / dt)),
                           "algorithm": ("polar-decomposition Jacobi, same algorithm as the GPU arm" if polar else
                                         "sequential Gauss-Seidel (reference); the GPU arm runs Jacobi on the same mesh -- see configs/jacobi_vs_gs in the GPU arm's line"),
                           "tets": M, "verts": verts.size // 3},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                                 "host_cores": os.cpu_count(), "dragon_substeps_per_s": cpu_dragon_substeps_per_s()},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import tetsim_b200 as ts
    from tetsim_b200 import _capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; tetsim_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert args.gpus == world, "--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world)

    verts, tets = mesh.make_beam(cells, jitter=args.jitter)
    N, M = verts.size // 3, tets.size // 4
    def make_nccl_id():
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            import ctypes
            raw = ctypes.create_string_buffer(128)
            _capi.check(_capi.lib().tetsim_nccl_unique_id(raw))
            buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=args.substeps, worldBounds=list(mesh.wide_bounds(64.0)))
    stream = torch.cuda.Stream(device=local_rank)   # a real (non-default) stream: the library enqueues on it, events time it
    torch.cuda.set_stream(stream)

    def make_body(nccl_id):
        if polar:
            assert world == 1, "the polar workload is single-GPU (multi-GPU tet partitioning exists for the NH Jacobi solver only)"
            return ts.SoftBodyGPU(verts, tets, None, pp, arithmetic="fast", cluster_size=args.cluster_size, reorder=not args.no_reorder,
                                  device=local_rank, stream=stream.cuda_stream)
        return ts.SoftBody(verts, tets, None, pp, solver="jacobi", arithmetic="fast", iters=args.iters,
                           cluster_size=args.cluster_size, reorder=not args.no_reorder, deterministic=not args.atomic,
                           device=local_rank, stream=stream.cuda_stream, rank=rank, world_size=world,
                           nccl_unique_id=nccl_id, exchange=args.exchange)

    body = make_body(make_nccl_id() if world > 1 and args.exchange != "peer" else None)
    if world > 1 and args.exchange == "peer":   # hand-shake of the peer-memory exchange buffers
        mine = torch.frombuffer(bytearray(body.ipc_handle()), dtype=torch.uint8).cuda()
        blobs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(blobs, mine)
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        try:
            body.set_peers([bytes(b.cpu().numpy().tobytes()) for b in blobs])
        except ts.TetSimError as e:   # e.g. no peer access between two devices: every rank falls back together
            sys.stderr.write("rank %d: peer-memory exchange unavailable (%s); falling back to the NCCL neighbour exchange\n" % (rank, e))
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            body.close()
            args.exchange = "halo"
            nccl_id = make_nccl_id()
            body = make_body(nccl_id)
    info = body.info()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- N > 1: the partitioned run against the same mesh on ONE GPU, outside every timed region ----
    parity = None
    if world > 1 and not args.no_parity:
        frames = 2                                        # 40 substeps from the initial state
        for _ in range(frames):
            body.step(pp)
        pos = torch.from_numpy(body.pos.copy()).cuda()    # caller order, NaN where not resident on this rank
        res = torch.from_numpy(body.resident.astype(np.uint8)).cuda()
        allpos = [torch.empty_like(pos) for _ in range(world)]
        allres = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(allpos, pos)
        dist.all_gather(allres, res)
        if rank == 0:
            P = torch.stack(allpos).cpu().numpy().reshape(world, N, 3)
            R = torch.stack(allres).cpu().numpy().astype(bool)
            merged = np.zeros((N, 3), np.float32)
            for r in range(world):
                merged[R[r]] = P[r][R[r]]
            shared = R.sum(axis=0) > 1
            identical = all(np.array_equal(P[r][R[r] & shared].view(np.uint32), merged[R[r] & shared].view(np.uint32)) for r in range(world))
            single = ts.SoftBody(verts, tets, None, pp, solver="jacobi", arithmetic="fast", iters=args.iters, cluster_size=args.cluster_size,
                                 reorder=not args.no_reorder, deterministic=not args.atomic, device=local_rank, stream=stream.cuda_stream)
            for _ in range(frames):
                single.step(pp)
            ref = single.pos.reshape(N, 3).astype(np.float64)
            single.close()
            err = float(np.max(np.linalg.norm(merged - ref, axis=1) / np.linalg.norm(ref, axis=1)))
            parity = {"err": err, "replicas_bit_identical": bool(identical), "every_vertex_resident": bool(R.any(axis=0).all()),
                      "shared_vertices": int(shared.sum()), "substeps": frames * args.substeps,
                      "against": "the same mesh, kernels and options on 1 GPU (vector-relative position error; the only difference "
                                 "is the summation order of the rank-shared vertices' partial sums)"}
        del allpos, allres, pos, res
        barrier()

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        body.step(pp)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = body.info()["kernelLaunches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        body.step(pp)
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = body.info()["kernelLaunches"] - launches0
    ms_step = ms_total / args.steps
    proj_per_step = M * args.iters * args.substeps
    value = proj_per_step / (ms_step * 1e-3) / 1e6

    # ---- sustained: the same loop for >= 2 s of device time (the K timed steps above are a burst of a fraction of a second;
    # this kernel is issue-bound, so its rate follows the SM clock, and a B200 under a long load settles below its boost) ----
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(2200.0 / ms_step) + 1)
        smp2 = ClockSampler(local_rank)
        smp2.start()
        e0.record(stream)
        for _ in range(n_sus):
            body.step(pp)
        e1.record(stream)
        barrier()
        ms_sus = max_over_ranks(e0.elapsed_time(e1)) / n_sus
        sustained = {"value": proj_per_step / (ms_sus * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_sus, "steps": n_sus,
                     "seconds_timed": ms_sus * n_sus / 1e3, "clocks": smp2.stop()}

    # ---- dominant kernel alone (roofline) ----
    try:
        k_ms, k_bytes = body.time_kernel(10 if polar else 50)
    except ts.TetSimError:      # a handle without a tile kernel (TETSIM_POLAR_CSR=1 comparison runs)
        k_ms, k_bytes = float("nan"), 0
    peak, peak_src = load_peaks()
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    kernel_key = "k_polar_tiles<%d>" % args.cluster_size if polar else "k_jacobi_tilesN<%d,2,2,%d>" % (args.cluster_size, 4 if args.cluster_size == 512 else 0)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if world == 1 and os.path.exists(tp) and tuple(cells) == (407, 64, 64):
        try:   # one ncu --set full capture of this kernel on this mesh (a number taken under ncu is never a bench value)
            ent = json.load(open(tp))["kernels"].get(kernel_key)
            if ent:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:
            traffic = None
    roofline = {"kernel": ("k_polar_tiles<%d> (tiled polar shape matching, tetsim_b200/csrc/kernels_fast.cu)" if polar else
                           "k_jacobi_tilesN<%d, 2 tets/thread> (persistent tile kernel, tetsim_b200/csrc/kernels_fast.cu)") % args.cluster_size, "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k_bytes, "ms_per_launch": k_ms,
                "share_of_step": k_ms * args.iters * args.substeps / ms_step,
                "tets_per_s_kernel_alone": info["localTets"] / (k_ms * 1e-3)}

    # ---- end to end through the C ABI with HOST buffers ----
    # Per step: upload this step's inputs (positions + velocities of the vertices resident on this rank, from pinned host
    # memory), solve the frame, download the resulting positions.  prevPos is NOT an input of simulate(): it is
    # overwritten at src/Softbody.js:200 before anything reads it.  Two figures: `serial` waits for each frame's
    # positions before the next upload starts; `value` (the headline) streams frames double-buffered -- the download of
    # frame k (tetsim_get_positions_resident_async) overlaps the upload and solve of frame k+1, each step still moving
    # its own inputs and outputs inside the timed region.
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        ids = body.resident_ids
        nloc = ids.size
        own = ids >= 0
        pos_r = np.nan_to_num(body.pos_resident)
        vel_c = np.nan_to_num(body.vel).reshape(-1, 3)
        vel_r = np.zeros((nloc, 3), np.float32)
        vel_r[own] = vel_c[ids[own]]
        h_pos = torch.from_numpy(pos_r).pin_memory()
        h_vel = torch.from_numpy(vel_r.reshape(-1)).pin_memory()
        h_out = [torch.empty(3 * nloc, dtype=torch.float32).pin_memory() for _ in range(2)]
        lib, hnd = _capi.lib(), body._h
        prm = ts.softbody._params_struct(pp)
        pp_, pv_ = C.c_void_p(h_pos.data_ptr()), C.c_void_p(h_vel.data_ptr())
        po_ = [C.c_void_p(t.data_ptr()) for t in h_out]

        def upload_and_step():
            _capi.check(lib.tetsim_set_state_resident(hnd, pp_, None, pv_))
            _capi.check(lib.tetsim_step(hnd, FRAME_DT, args.substeps, C.byref(prm)))

        def serial_step():
            upload_and_step()
            _capi.check(lib.tetsim_get_positions_resident(hnd, po_[0]))

        def streamed(n):
            upload_and_step()
            _capi.check(lib.tetsim_get_positions_resident_async(hnd, po_[0]))
            for k in range(1, n):
                upload_and_step()
                _capi.check(lib.tetsim_wait_positions(hnd))          # frame k-1 is in host memory
                _capi.check(lib.tetsim_get_positions_resident_async(hnd, po_[k & 1]))
            _capi.check(lib.tetsim_wait_positions(hnd))

        n_e2e = max(3, min(args.steps, 20))
        for _ in range(2):
            serial_step()
        barrier()
        e0.record(stream)
        for _ in range(n_e2e):
            serial_step()
        e1.record(stream)
        barrier()
        ms_serial = max_over_ranks(e0.elapsed_time(e1)) / n_e2e
        streamed(3)
        barrier()
        e0.record(stream)
        streamed(n_e2e)
        e1.record(stream)
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / n_e2e
        assert np.isfinite(h_out[(n_e2e - 1) & 1].numpy()[np.repeat(own, 3)]).all()
        tot = torch.tensor([float(nloc)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        nres = int(tot.item())
        e2e = {"value": proj_per_step / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": 2 * 3 * nres * 4, "d2h_bytes_per_step": 3 * nres * 4, "steps": n_e2e,
               "mode": "frames streamed double-buffered: the download of frame k overlaps the upload + solve of frame k+1",
               "serial": {"value": proj_per_step / (ms_serial * 1e-3) / 1e6, "ms_per_step": ms_serial,
                          "mode": "each frame's positions are in host memory before the next upload starts"},
               "api": "tetsim_set_state_resident(pos, vel from pinned host; all ranks' resident vertices) + tetsim_step + "
                      "tetsim_get_positions_resident[_async] (to pinned host)"}

    clocks = sampler.stop()   # sampled across the timed steps, the kernel-alone loop and the end-to-end steps

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec = cpu_reference_run(verts, tets, args.cpu_substeps, dt, polar)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
               "dragon_substeps_per_s": cpu_dragon_substeps_per_s(),
               "sample": "%d substeps of the full %d-tet mesh (%.1f s), sequential Gauss-Seidel C restatement of "
                         "src/Softbody.js, 1 thread (the sweep is inherently sequential; no JS engine in image)"
                         % (args.cpu_substeps, M, sec)}

    configs = None
    if not args.no_configs:
        body.synchronize()
        c5 = c5_sharded(stream, rank, world, local_rank, dist if world > 1 else None)   # every rank takes part
        if rank == 0:
            configs = config_rates(stream, local_rank) if world == 1 else {}
            configs["C5_784x_gs_exact_fast_sharded"] = c5

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "tets": M, "verts": N, "iters": args.iters, "substeps_per_step": args.substeps,
                       "cluster_size": info["clusterSize"], "clusters_rank0": info["numClusters"],
                       "boundary_verts": info["boundaryVerts"], "boundary_tiles_rank0": info["boundaryTiles"], "deterministic": not args.atomic,
                       "parallelism": ("tet-partition x%d (RCB), %s of boundary dx per iteration, overlapped with interior tiles" % (world, {"allreduce": "ncclAllReduce", "halo": "neighbour ncclSend/ncclRecv", "peer": "fused peer-memory exchange: the tile kernel stores its boundary partials into the sharers' buffers (cudaIpc over NVLink, no NCCL on the data path)"}[args.exchange])) if world > 1 else "single GPU",
                       "exchange": args.exchange if world > 1 else None,
                       "nvlink_bytes_per_iteration": (info["boundaryVerts"] * 2.4 * 32 * 1 if world > 1 and args.exchange == "peer" else None),
                       "nvlink_bytes_note": ("peer exchange: every tile partial of a rank-shared vertex (about 2.4 per vertex and rank on this mesh) is one 32-byte "
                                             "tagged entry stored into each OTHER sharer's buffer; figure = boundary_verts x 2.4 x 32 B x (sharers - 1 = 1 for planar cuts), per direction and iteration, all cuts together") if world > 1 else None,
                       "l2": "working set per substep (%.0f MB) exceeds the 126 MB L2; no flush needed" % ((56.0 * M + 144.0 * N) / 1e6)},
            "scalar_constraints_per_s_M": 2 * value,
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
            "sustained": sustained,
        }
        if parity is not None:
            line["parity"] = parity
        if configs is not None:
            line["configs"] = configs
        print(json.dumps(line))
    barrier()   # every rank idle before any rank frees its (possibly peer-mapped) buffers
    body.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
