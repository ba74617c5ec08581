"""Worker of the multi-process tests (launched by torch.distributed.run).

    mode host : gloo, host-only sets (device -1): DataBus.create_custom_types + start-up sync_nodes +
                pre_process_meshes across 2 ranks; every rank checks its zones' flags/faces against
                the golden fixture (the reference ran the same zones on one rank).
    mode gpu  : nccl, one GPU per rank: full steps with the per-stage NCCL halo exchange; values
                compared bit for bit with the golden fixture.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "gcm-necro_b200"))
import util_parity as up  # noqa: E402
import gcm_b200  # noqa: E402
from gcm_b200.distributed import DataBus  # noqa: E402


def main():
    mode, name = sys.argv[1], sys.argv[2]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "gpu":
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    else:
        dist.init_process_group("gloo")
    if name == "mp24":
        # 24^3 cube in 4 z-slab zones, 2 per rank; the golden reference is a single-process run of the
        # product on this rank's GPU, compared after THREE steps (12 stage exchanges, both record layers in
        # use) on nodes >= 10 cells away from the outer border (a step spreads information by 3 cells; the
        # border bases depend on the per-rank rand() stream, see tests/test_multiprocess.py)
        from gcm_b200 import meshgen
        zs = meshgen.make_box_zones(24, 24, 24, 4, jitter=0.1, seed=2)
        case = dict(name=name, zones=zs, states=[meshgen.generic_state(z.coords, up.LA, up.MU, up.RHO, seed=6) for z in zs],
                    material=(up.LA, up.MU, up.RHO, 1e16), steps=3, seed=5, borders=[], tap_step=-1)
        one = up.run_product(case, device=lr)
        gold = []
        for z, r in zip(zs, one):
            c = np.rint(z.coords)
            deep = np.all((c >= 10) & (c <= 13), axis=1) & (z.placement == 1)
            fl = r["flags"].copy()
            fl[~deep] &= ~2          # compare() / local_mask() only look at Local nodes
            gold.append(dict(values=r["values"], flags=r["flags"], faces=r["faces"], cmp_flags=fl))
    else:
        case = up.make_case(name)
        gold = up.load_golden(name)
    nz = len(case["zones"])
    owner = [z * world // nz for z in range(nz)]
    mine = [z for z in range(nz) if owner[z] == rank]
    la, mu, rho, yl = case["material"]
    ms = gcm_b200.TetrMeshSet(n_zones=nz, zone_rank=owner, my_rank=rank, device=lr if mode == "gpu" else -1)
    ms.srand(case["seed"])
    for z in mine:
        ms.get_mesh_by_zone_num(z).load(case["zones"][z], la=la, mu=mu, rho=rho, yield_limit=yl)
    bus = DataBus(ms)
    bus.create_custom_types()
    bus.sync_nodes_startup()
    ms.pre_process_meshes()
    bad = []
    for z in mine:
        m = ms.get_mesh_by_zone_num(z)
        if not np.array_equal(m.get_flags() & 15, gold[z]["flags"]):
            bad.append("zone %d flags" % z)
        if not np.array_equal(m.get_faces(), gold[z]["faces"]):
            bad.append("zone %d faces" % z)
    if mode == "gpu":
        for z in mine:
            ms.get_mesh_by_zone_num(z).set_values(case["states"][z])
        for _ in range(case["steps"]):
            bus.do_next_step()
        ms.synchronize()
        for z in mine:
            v = ms.get_mesh_by_zone_num(z).get_values()
            loc = up.local_mask(gold[z].get("cmp_flags", gold[z]["flags"]))
            if not np.array_equal(v[loc], gold[z]["values"][loc]):
                bad.append("zone %d values (%d differ)" % (z, int((v[loc] != gold[z]["values"][loc]).sum())))
    allbad = [None] * world
    dist.all_gather_object(allbad, bad)
    ms.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        flat = [b for bb in allbad for b in bb]
        print("MP_RESULT", "OK" if not flat else flat)
    sys.exit(0 if not any(allbad) else 1)


if __name__ == "__main__":
    main()
