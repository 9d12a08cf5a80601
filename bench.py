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
    ap.add_argument("--substeps", type=int, default=20, help="substeps per step (frame)")
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--cluster-size", type=int, default=512, help="tets per tile (512 measured fastest on one GPU, profiles/r1_tile_sweep.txt)")
    ap.add_argument("--no-reorder", action="store_true")
    ap.add_argument("--atomic", action="store_true", help="deterministic=0: REDG flush instead of per-tile partials")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the same 10M-tet mesh split over N GPUs (BASELINE config 4); weak: beam length x N")
    ap.add_argument("--cpu-substeps", type=int, default=6, help="substeps of the CPU baseline sample (rank 0, N=1)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "allreduce", "halo", "peer"],
                    help="multi-GPU boundary exchange: ncclAllReduce over all ranks, or grouped ncclSend/ncclRecv with the neighbour ranks; "
                         "auto = all-reduce at 2 GPUs, neighbour exchange beyond (measured 22 %% faster at 8 GPUs, profiles/r1_scaling.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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


def cpu_reference_run(verts, tets, substeps, dt):
    """The reference's CPU algorithm (sequential Gauss-Seidel, src/Softbody.js:195-240) via the C restatement."""
    import oracle
    ref = oracle.SoftBodyOracle(verts, tets)
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


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.exchange == "auto":
        args.exchange = "halo" if world > 2 else "allreduce"
    cells = tuple(int(c) for c in args.cells.split(","))
    if args.scaling == "weak":
        cells = (cells[0] * max(world, 1), cells[1], cells[2])
    dt = FRAME_DT / args.substeps
    from tetsim_b200 import mesh

    workload = "beam %dx%dx%d cells Kuhn-split, NH Jacobi iters=%d, dt=1/%d, %d substeps/step" % (
        cells[0], cells[1], cells[2], args.iters, round(1.0 / dt), args.substeps)

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        verts, tets = mesh.make_beam(cells)
        M = tets.size // 4
        import oracle
        ref = oracle.SoftBodyOracle(verts, tets)
        per_step = 1  # bounded sample: 1 substep of the full mesh per bench step
        for _ in range(args.warmup):
            ref.simulate(dt)
        t0 = time.perf_counter()
        for _ in range(args.steps * per_step):
            ref.simulate(dt)
        sec = time.perf_counter() - t0
        val = M * args.steps * per_step / sec / 1e6
        sample = "%d substep(s) of the full %d-tet mesh per step; sequential Gauss-Seidel, C restatement of src/Softbody.js (no JS engine in image)" % (per_step, M)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64-expr/f32-store",
                "data": "synthetic", "config": {"workload": workload, "tets": M, "verts": verts.size // 3},
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

    verts, tets = mesh.make_beam(cells)
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

    # ---- dominant kernel alone (roofline) ----
    k_ms, k_bytes = body.time_kernel(50)
    peak, peak_src = load_peaks()
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_jacobi_tiles_dram_bytes_per_launch") if world == 1 else None
        except Exception:
            traffic = None
    roofline = {"kernel": "k_jacobi_tilesN<%d, 2 tets/thread> (persistent tile kernel, tetsim_b200/csrc/kernels_fast.cu)" % args.cluster_size, "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k_bytes, "ms_per_launch": k_ms,
                "share_of_step": k_ms * args.iters * args.substeps / ms_step,
                "tets_per_s_kernel_alone": info["localTets"] / (k_ms * 1e-3)}

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        res = body.resident
        h_pos = torch.from_numpy(np.nan_to_num(body.pos.copy())).pin_memory()
        h_prev = torch.from_numpy(np.nan_to_num(body.prevPos.copy())).pin_memory()
        h_vel = torch.from_numpy(np.nan_to_num(body.vel.copy())).pin_memory()
        h_out = torch.empty(3 * N, dtype=torch.float32).pin_memory()
        lib, hnd = _capi.lib(), body._h
        import ctypes as C
        prm = ts.softbody._params_struct(pp)

        def e2e_step():
            _capi.check(lib.tetsim_set_state(hnd, C.c_void_p(h_pos.data_ptr()), C.c_void_p(h_prev.data_ptr()),
                                             C.c_void_p(h_vel.data_ptr())))
            _capi.check(lib.tetsim_step(hnd, FRAME_DT, args.substeps, C.byref(prm)))
            _capi.check(lib.tetsim_get_positions(hnd, C.c_void_p(h_out.data_ptr())))

        for _ in range(2):
            e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        e0.record(stream)
        for _ in range(n_e2e):
            e2e_step()
        e1.record(stream)
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / n_e2e
        e2e = {"value": proj_per_step / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": 3 * 3 * N * 4, "d2h_bytes_per_step": 3 * N * 4, "steps": n_e2e,
               "api": "tetsim_set_state(pos,prev,vel from pinned host) + tetsim_step + tetsim_get_positions(to pinned host)"}
        del res

    clocks = sampler.stop()   # sampled across the timed steps, the kernel-alone loop and the end-to-end steps

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec = cpu_reference_run(verts, tets, args.cpu_substeps, dt)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
               "dragon_substeps_per_s": cpu_dragon_substeps_per_s(),
               "sample": "%d substeps of the full %d-tet mesh (%.1f s), sequential Gauss-Seidel C restatement of "
                         "src/Softbody.js, 1 thread (the sweep is inherently sequential; no JS engine in image)"
                         % (args.cpu_substeps, M, sec)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "tets": M, "verts": N, "iters": args.iters, "substeps_per_step": args.substeps,
                       "cluster_size": info["clusterSize"], "clusters_rank0": info["numClusters"],
                       "boundary_verts": info["boundaryVerts"], "boundary_tiles_rank0": info["boundaryTiles"], "deterministic": not args.atomic,
                       "parallelism": ("tet-partition x%d (RCB), %s of boundary dx per iteration, overlapped with interior tiles" % (world, {"allreduce": "ncclAllReduce", "halo": "neighbour ncclSend/ncclRecv", "peer": "peer-memory stores (cudaIpc over NVLink, no NCCL)"}[args.exchange])) if world > 1 else "single GPU",
                       "l2": "working set per substep (%.0f MB) exceeds the 126 MB L2; no flush needed" % ((56.0 * M + 144.0 * N) / 1e6)},
            "scalar_constraints_per_s_M": 2 * value,
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    barrier()   # every rank idle before any rank frees its (possibly peer-mapped) buffers
    body.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
