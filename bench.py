#!/usr/bin/env python
"""bench.py -- tet node-updates/s of the grid-characteristic step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one the metric is quoted on): elastic cube of
216^3 = 10 077 696 nodes / 59 630 250 tets (create_cube.c semantics, 6 tets per hex), split into
8 z-slab zones with one halo layer, material la=7e9 mu=1e9 rho=7.8e6, plane P-wave initial state,
tau = 0.002, srand(12345).  One *step* = TetrMeshSet::do_next_step = 3 axis stages + plastic
stage on every local node, with the halo exchange before each axis stage.  Strong scaling: the 8 zones
are dealt to the N ranks (8/N zones per GPU); the cross-rank halo goes over NVLink (direct peer
stores into the neighbour's halo slots, issued beside the inner nodes that do not need them; NCCL
send/recv as the fallback).  The two end slabs carry a whole 216^2 face of border nodes each, and a
border node costs ~20 inner nodes, so by default the slabs are cut for equal work (--layers auto:
12,32,32,32,32,32,32,12 layers) instead of equal thickness (--layers equal); config.layers says which.
One JSON line is printed by rank 0 (see the task contract for the keys).  Extra keys: roofline.* next to the
contract's figure (kernel_bytes_per_node_stage / frac_kernel_bytes: what the cached inner kernel itself has to
move; traffic / traffic_frac: DRAM bytes ncu measured for these kernel sources), other_configs (1 GPU: the
short lines of BASELINE configs[1], [3], [4]), host_issue_ms_per_step, and state_digest = a checksum of the
values of the nodes >= 80 cells away from the cube's surface after the timed and the per-kernel steps (before
the e2e phase): it does not depend on how the zones are dealt to ranks (the border bases do: every rank owns a
rand() stream, as every MPI rank of the reference does), so runs at 1/2/4/8 GPUs with the same --steps and
--warmup must print the same one.  --workload contact prints the configs[4] line alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gcm-necro_b200"))

LA, MU, RHO, YIELD = 7e9, 1e9, 7.8e6, 1e16
B_STAGE = 292          # algorithmic bytes per node per axis stage (SURVEY 8d / DESIGN.md)
B_PLASTIC = 52
B_STEP = 3 * B_STAGE + B_PLASTIC   # 928 B per node-update
ORACLE = os.path.join(ROOT, "oracle", "_ref", "gcm3d_ref")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                sm.append(float(c[1])); mx.append(float(c[2]))
                for k, nm in enumerate(names):
                    if c[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def zone_owner(n_zones, world):
    return [z * world // n_zones for z in range(n_zones)]


def slab_layers(spec, nz_layers, n_zones, border_weight):
    """Per-zone owned layers.  auto: an end slab's extra face of border nodes costs about
    `border_weight` layers of inner nodes, so end = mid - border_weight (sum preserved)."""
    if spec == "equal" or n_zones < 3:
        if nz_layers % n_zones:
            raise SystemExit("--size must be divisible by --zones for equal slabs")
        return [nz_layers // n_zones] * n_zones
    if spec != "auto":
        v = [int(x) for x in spec.split(",")]
        if len(v) != n_zones or sum(v) != nz_layers:
            raise SystemExit("--layers must list %d values summing to %d" % (n_zones, nz_layers))
        return v
    mid = int(round((nz_layers + 2.0 * border_weight) / n_zones))
    end = (nz_layers - (n_zones - 2) * mid) // 2
    v = [end] + [mid] * (n_zones - 2) + [nz_layers - (n_zones - 2) * mid - end]
    if min(v) < 2:
        return slab_layers("equal", nz_layers, n_zones, border_weight)
    return v


def kernel_source_hash():
    """sha256 over the kernel sources: stamps measurements that were taken outside this run (ncu)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "gcm-necro_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cuh", ".cu", ".cpp", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def run_reference(args, rank):
    """The reference's own CPU implementation (oracle/_ref = its sources compiled here), timed on
    the box's host cores on a bounded sample of the same workload."""
    if rank != 0:
        return
    from gcm_b200 import meshgen
    size, nz = args.ref_size, args.ref_zones
    if not os.path.exists(ORACLE):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/gcm3d_ref not built (run __graft_entry__.build() where /root/reference exists)"}))
        return
    zones = meshgen.make_box_zones(size, size, size, nz)
    with tempfile.TemporaryDirectory() as d:
        meshes, states = [], []
        for k, z in enumerate(zones):
            mp, sp = os.path.join(d, "z%d.gcmb" % k), os.path.join(d, "z%d.state" % k)
            meshgen.write_gcmb(mp, z)
            meshgen.pwave_state(z.coords, LA, MU, RHO, z_c=size / 2.0).tofile(sp)
            meshes.append(mp); states.append(sp)
        k, w = max(1, min(args.steps, args.ref_steps)), min(args.warmup, 1)
        cmd = [ORACLE, "--mesh"] + meshes + ["--state"] + states + ["--steps", str(k + w), "--warmup", str(w), "--seed", "12345"]
        r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(json.dumps({"impl": "reference", "unavailable": "oracle run failed: " + r.stderr[-200:].replace("\n", " ")}))
        return
    j = json.loads(r.stdout.strip().splitlines()[-1])
    n_nodes = j["local_nodes"]
    v = n_nodes * j["timed_steps"] / j["seconds"]
    sample = ("%d^3 = %d-node cube in %d zone(s) on one rank, %d timed step(s) of the reference's own do_next_step (same mesh family / "
              "material / initial state as the 216^3 workload, which needs ~8 GB and ~4 min per step on this path; "
              "node-updates/s is per node, so the figure carries over)" % (size, n_nodes, nz, j["timed_steps"]))
    line = {"impl": "reference", "metric": "tet node-updates/sec", "value": v, "unit": "node-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * j["seconds"] / j["timed_steps"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "elastic cube 216^3 nodes in 8 zones (BASELINE configs[2]); reference timed on a bounded sample", "sample": sample,
                       "same_mesh_size": False},
            "cpu_baseline": {"value": v, "unit": "node-updates/s", "cores": 1, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "node-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline(args):
    """Bounded CPU sample for the b200 arm's line (rank 0, N=1)."""
    from gcm_b200 import meshgen
    if not os.path.exists(ORACLE):
        return {"value": None, "unit": "node-updates/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    size = args.cpu_size
    z = meshgen.make_cube(size)
    with tempfile.TemporaryDirectory() as d:
        mp, sp = os.path.join(d, "m.gcmb"), os.path.join(d, "m.state")
        meshgen.write_gcmb(mp, z)
        meshgen.pwave_state(z.coords, LA, MU, RHO).tofile(sp)
        r = subprocess.run([ORACLE, "--mesh", mp, "--state", sp, "--steps", "3", "--warmup", "1", "--seed", "12345"],
                           capture_output=True, text=True)
    if r.returncode != 0:
        return {"value": None, "unit": "node-updates/s", "cores": 1, "kind": "reference", "sample": "oracle run failed"}
    j = json.loads(r.stdout.strip().splitlines()[-1])
    return {"value": j["local_nodes"] * j["timed_steps"] / j["seconds"], "unit": "node-updates/s", "cores": 1,
            "kind": "reference",
            "sample": "%d^3 = %d-node single-zone cube, %d timed steps of the reference's do_next_step (its sources, g++ -O3, single thread as the reference is)" % (size, j["local_nodes"], j["timed_steps"])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=216, help="cube side in nodes (216 -> 10 077 696 nodes)")
    ap.add_argument("--zones", type=int, default=8)
    ap.add_argument("--ref-size", type=int, default=100, help="cube side of the bounded sample of --impl reference (100 -> 10^6 nodes, BASELINE configs[1] size)")
    ap.add_argument("--ref-zones", type=int, default=1)
    ap.add_argument("--ref-steps", type=int, default=2, help="timed steps of the reference arm (about 11 s each at 10^6 nodes)")
    ap.add_argument("--layers", default="auto", help="auto | equal | comma list of owned z-layers per zone")
    ap.add_argument("--border-weight", type=float, default=20.0, help="cost of a face of border nodes in layers of inner nodes (--layers auto); measured: a border node costs ~20 inner nodes per step (12/32-layer slabs: 0.301 ms per step at 8 GPUs against 0.333 with 6/34)")
    ap.add_argument("--cpu-size", type=int, default=64, help="cube side of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="cube", choices=["cube", "contact"], help="contact: only the two-body configs[4] line")
    ap.add_argument("--contact-side", type=int, default=126)
    ap.add_argument("--contact-calc", default="adhesion", choices=["adhesion", "friction"])
    ap.add_argument("--contact-every-step", action="store_true", help="run the collision detector every step instead of static")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the short configs[1] / configs[3] runs appended at 1 GPU")
    ap.add_argument("--yield-limit", type=float, default=YIELD, help="1e7 = BASELINE configs[3] (plastic corrector fires)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import gcm_b200
    from gcm_b200 import meshgen
    from gcm_b200.distributed import DataBus

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload == "contact":
        # BASELINE configs[4] alone (1 GPU): a short line of its own, for profiling the contact kernels
        if rank == 0:
            print(json.dumps({"metric": "tet node-updates/sec", "config": {"workload": "two %d^3 bodies in contact (BASELINE configs[4])" % args.contact_side},
                              **run_contact(args, local_rank, args.contact_side, args.contact_calc, not args.contact_every_step)}))
        return
    line = run_workload(args, rank, world, local_rank, args.size, args.zones, args.yield_limit, full=True)
    if rank == 0 and world == 1 and not args.no_extra_configs:
        # the other single-GPU configurations BASELINE.json names, as extra keys of the one line (short runs, device-timed)
        extra = {}
        extra["configs[1] single cube 100^3, 1 zone"] = run_workload(args, rank, world, local_rank, 100, 1, YIELD, full=False)
        extra["configs[3] plastic corrector, yield 1e7, 216^3 in 8 zones"] = run_workload(args, rank, world, local_rank, args.size, args.zones, 1e7, full=False)
        for calc, static in (("adhesion", True), ("friction", True), ("adhesion", False)):
            extra["configs[4] two 126^3 cubes 0.01 apart, Bruteforce detector %s, %s contact"
                  % ("static" if static else "every step", calc)] = run_contact(args, local_rank, 126, calc, static)
        line["other_configs"] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_contact(args, local_rank, side, calc, static):
    """BASELINE configs[4]: two bodies of side^3 nodes, the second translated by side - 1 + 0.01 along z, Bruteforce
    collision detector (threshold 1), Adhesion or Friction contact on ~2 x (side-2)^2 nodes.  State: a compression
    pulse across the interface, plus opposite shear velocities for the friction run (v_tau != 0)."""
    import torch
    import gcm_b200
    from gcm_b200 import meshgen
    t0 = time.time()
    za, zb = meshgen.make_cube(side), meshgen.make_cube(side)
    zb.zone = 1
    ms = gcm_b200.TetrMeshSet(n_zones=2, device=local_rank)
    ms.srand(12345)
    ms.get_mesh_by_zone_num(0).load(za, la=LA, mu=MU, rho=RHO, yield_limit=YIELD)
    ms.get_mesh_by_zone_num(1).load(zb, la=LA, mu=MU, rho=RHO, yield_limit=YIELD)
    ms.get_mesh_by_zone_num(1).translate(0.0, 0.0, side - 1 + 0.01)
    ms.set_collision_detector("brute", static=static, threshold=1.0)
    ms.set_contact_calculator(calc)
    ms.pre_process_meshes()
    sa = meshgen.pwave_state(za.coords, LA, MU, RHO, z_c=side - 1.0)
    sb = meshgen.pwave_state(zb.coords, LA, MU, RHO, z_c=0.0)
    if calc == "friction":
        sa[:, 0] += 0.3 * sa[:, 2]
        sb[:, 0] -= 0.3 * sb[:, 2]
    ms.get_mesh_by_zone_num(0).set_values(sa)
    ms.get_mesh_by_zone_num(1).set_values(sb)
    n_total = 2 * side ** 3
    t_setup = time.time() - t0
    stream = torch.cuda.ExternalStream(ms.stream())
    for _ in range(args.warmup):
        ms.do_next_step()
    ms.synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ms.launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        ms.do_next_step()
    e1.record(stream)
    ms.synchronize()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    launches = ms.launch_count() - l0
    ms.set_concurrency(False)
    ms.profile(True)
    for _ in range(min(args.steps, 5)):
        ms.do_next_step()
    ms.synchronize()
    prof = {k: ms.profile_read(i) for i, k in enumerate(["inner_stage", "border_stage", "plastic", "halo_copy", "rng_bases", "inner_generic"])}
    ms.profile(False)
    peak, _ = measured_peak_gbs()
    value = n_total * args.steps / (total_ms * 1e-3)
    out = {"nodes": n_total, "contact_nodes": ms.virt_count(), "ms_per_step": total_ms / args.steps, "value": value, "unit": "node-updates/s",
           "step_frac": value * B_STEP / 1e9 / peak, "steps": args.steps, "gpu_launches": int(launches), "setup_s": t_setup,
           "kernel_ms_per_5_steps_serialised": {k: v[0] for k, v in prof.items()}, "coverage": ms.kernel_stats()}
    ms.close()
    return out


def run_workload(args, rank, world, local_rank, size, nz, yield_limit, full):
    """One workload on the current process group.  full: the whole JSON line (roofline pass, e2e, cpu baseline,
    digest); else a short summary (ms/step, node-updates/s, step-level roofline fraction)."""
    import torch
    import torch.distributed as dist
    import gcm_b200
    from gcm_b200 import meshgen
    from gcm_b200.distributed import DataBus

    owner = zone_owner(nz, world)
    mine = [z for z in range(nz) if owner[z] == rank]
    t0 = time.time()
    layers = slab_layers(args.layers, size, nz, args.border_weight)
    zones = meshgen.make_box_zones(size, size, size, nz, only=set(mine), layers=layers)
    ms = gcm_b200.TetrMeshSet(n_zones=nz, zone_rank=owner, my_rank=rank, device=local_rank)
    ms.srand(12345)
    for z in mine:
        ms.get_mesh_by_zone_num(z).load(zones[z], la=LA, mu=MU, rho=RHO, yield_limit=yield_limit)
    bus = None
    if world > 1:
        bus = DataBus(ms)
        bus.create_custom_types()
        bus.sync_nodes_startup()
    ms.pre_process_meshes()
    n_local = 0
    host = {}
    deep = {}
    for z in mine:
        m = ms.get_mesh_by_zone_num(z)
        st = meshgen.pwave_state(zones[z].coords, LA, MU, RHO, z_c=size / 2.0)
        m.set_values(st)
        n_local += m.counts()["local_nodes"]
        host[z] = torch.from_numpy(st.copy()).pin_memory()
        c = zones[z].coords
        margin = min(80, size // 3)
        sel = np.all((c >= margin) & (c <= size - 1 - margin), axis=1) & (zones[z].placement == 1)
        gid = (np.rint(c[sel, 2]).astype(np.int64) * size + np.rint(c[sel, 1]).astype(np.int64)) * size + np.rint(c[sel, 0]).astype(np.int64)
        deep[z] = (np.nonzero(sel)[0], gid)
    zones = None
    t_setup = time.time() - t0
    stream = torch.cuda.ExternalStream(ms.stream())

    def step():
        if bus is None:
            ms.do_next_step()
        else:
            bus.do_next_step()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_issue = [0.0]

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t_host = time.perf_counter()
        for _ in range(k):
            fn()
        e1.record(stream)
        host_issue[0] = (time.perf_counter() - t_host) * 1e3 / k    # host time to ISSUE a step (no waiting inside)
        ms.synchronize()
        barrier()
        ms_ = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms_, op=dist.ReduceOp.MAX)
        return float(ms_.item())

    for _ in range(args.warmup):
        step()
    ms.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ms.launch_count()
    total_ms = timed(step, args.steps)
    launches = ms.launch_count() - l0
    host_issue_ms = host_issue[0]
    clocks = sampler.stop() if rank == 0 else None

    n_total = torch.tensor([n_local], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(n_total)
    n_total = int(n_total.item())
    value = n_total * args.steps / (total_ms * 1e-3)
    if not full:
        peak, _ = measured_peak_gbs()
        out = {"nodes": n_total, "zones": nz, "ms_per_step": total_ms / args.steps, "value": value, "unit": "node-updates/s",
               "step_frac": value / world * B_STEP / 1e9 / peak, "steps": args.steps, "yield_limit": yield_limit,
               "gpu_launches": int(launches), "coverage": ms.kernel_stats()}
        ms.close()
        return out

    # ---- per-kernel timing for the roofline (same steps, events around each launch) ----------
    # The timed loop above runs the border kernels on a second stream beside the inner kernel; for the
    # per-kernel durations the streams are serialised, so that each kernel is timed alone (the events sit
    # on the stream the kernel is launched on) and not stretched by whatever shares the SMs with it.
    concurrent = os.environ.get("GCM_B200_CONCURRENT", "1") != "0"
    ms.set_concurrency(False)
    ms.profile(True)
    for _ in range(min(args.steps, 5)):
        step()
    ms.synchronize()
    ms.set_concurrency(concurrent)
    prof = {k: ms.profile_read(i) for i, k in enumerate(["inner_stage", "border_stage", "plastic", "halo_copy", "rng_bases", "inner_generic"])}
    ms.profile(False)
    peak, peak_src = measured_peak_gbs()
    in_ms, in_launches, in_units = prof["inner_stage"]
    achieved = (B_STAGE * in_units) / (in_ms * 1e-3) / 1e9 if in_ms > 0 else 0.0
    traffic = None
    try:
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the inner kernel on this workload,
        # from the committed ncu --set full capture (profiles/README.md says which)
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
        # only when it was captured from exactly these kernel sources on this workload: a stale figure is worse than none
        if args.size == t.get("size") and args.zones == t.get("zones") and world == 1 and t.get("kernel_sha") == kernel_source_hash():
            traffic = t["bytes_per_launch"]
    except Exception:
        pass
    avg_launch_s = in_ms * 1e-3 / max(in_launches, 1)
    kb = 100       # what the cached inner kernel itself has to move per node-stage: record in 48 + record out 48 + pattern id 4
    roofline = {"bound": "hbm", "kernel": "inner-node axis stage", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                "note": "achieved/frac use SURVEY 8d's 292 algorithmic bytes per node-stage (CSR + tet table read every stage); the "
                        "round-2 kernel caches owner and weights of inner nodes, so it beats that figure -- see kernel_bytes / traffic_frac",
                "kernel_bytes_per_node_stage": kb,
                "frac_kernel_bytes": (kb * in_units) / (in_ms * 1e-3) / 1e9 / peak if in_ms > 0 else None,
                "traffic_frac": (traffic / avg_launch_s / 1e9 / peak) if (traffic and avg_launch_s > 0) else None,
                "algorithmic_bytes_per_node_stage": B_STAGE, "avg_launch_ms": in_ms / max(in_launches, 1),
                "nodes_per_launch": in_units / max(in_launches, 1),
                "step_frac": value / world * B_STEP / 1e9 / peak,
                "kernel_timing": "CUDA events around each launch, in a pass with the border stream serialised behind the inner kernel",
                "kernel_ms_share": {k: v[0] for k, v in prof.items()},
                "kernel_sha": kernel_source_hash(), "coverage": ms.kernel_stats()}

    # ---- state digest: nodes >= 80 cells from the surface (rank-independent, see the module docstring) ----
    ms.synchronize()
    dig = np.zeros(2, np.int64)
    for z in mine:
        idx, gid = deep[z]
        if len(idx):
            v = ms.get_mesh_by_zone_num(z).get_values()[idx].view(np.uint32).astype(np.int64)
            wts = (gid % 1000003 + 1)[:, None] * (np.arange(9, dtype=np.int64) + 1)[None, :]
            dig[0] += int((v * wts % 2147483647).sum() % 2147483647)
            dig[1] += len(idx)
    digt = torch.from_numpy(dig).cuda()
    if world > 1:
        dist.all_reduce(digt)
    state_digest = "%d:%d" % (int(digt[1].item()), int(digt[0].item()) % 2147483647)
    steps_before_digest = int(round(ms.get_current_time() / 0.002))

    # ---- end to end through the C ABI with HOST buffers ---------------------------------------
    e2e = None
    if not args.no_e2e:
        host_out = {z: torch.empty_like(host[z]).pin_memory() for z in mine}

        def e2e_step():
            # one batch: this step's input state from pinned host memory, the step, its result back to pinned host memory.
            # The copies run on the library's own streams, so the upload of the next batch and the download of the last
            # one overlap (full duplex); nothing waits on the host inside the loop
            for z in mine:
                ms.get_mesh_by_zone_num(z).upload_values_async(host[z].data_ptr())
            step()
            for z in mine:
                ms.get_mesh_by_zone_num(z).download_values_async(host_out[z].data_ptr())

        def e2e_loop():
            for _ in range(k2):
                e2e_step()
            ms.join_io()          # the closing event on the set's stream sits behind the last download
        k2 = max(1, min(args.steps, 10))
        e2e_step()
        ms.synchronize()
        e_ms = timed(e2e_loop, 1)
        bytes_io = sum(int(host[z].numel()) * 4 for z in mine)
        e2e = {"value": n_total * k2 / (e_ms * 1e-3), "unit": "node-updates/s", "steps": k2,
               "h2d_bytes_per_step": bytes_io, "d2h_bytes_per_step": bytes_io,
               "note": "per rank and step: pinned host values[9N] -> device (H2D + transpose), do_next_step, device -> pinned host (transpose + D2H); "
                       "the three stages are pipelined over the library's copy streams, one host synchronisation after the last step"}
        chk = ms.get_mesh_by_zone_num(mine[0]).get_values()
        if not np.array_equal(chk, host_out[mine[0]].numpy().reshape(chk.shape)):
            raise SystemExit("e2e: the downloaded state differs from the device state")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args)

    line = None
    if rank == 0:
        line = {"metric": "tet node-updates/sec", "value": value, "unit": "node-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "elastic cube %d^3 = %d nodes, 6 tets/hex, %d z-slab zones (BASELINE configs[2]%s)"
                                       % (size, n_total, nz, "; yield 1e7 = configs[3]" if yield_limit < 1e15 else ""),
                           "zones_per_gpu": nz // world if nz % world == 0 else [owner.count(r) for r in range(world)],
                           "layers": layers, "layers_rule": args.layers,
                           "tau": 0.002, "material": [LA, MU, RHO, yield_limit], "seed": 12345,
                           "l2": "inputs larger than L2 (state %.0f MB + mesh > 126 MB per GPU)" % (n_total / world * 72 / 1e6),
                           "setup_s": t_setup},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "host_issue_ms_per_step": host_issue_ms, "clocks": clocks,
                "state_digest": state_digest, "steps_total_before_digest": steps_before_digest}
    ms.close()
    return line


if __name__ == "__main__":
    main()
