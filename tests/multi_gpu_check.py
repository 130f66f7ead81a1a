"""Multi-GPU parity check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank solves the same graph (a) alone (world=1) and (b) as its shard of the N-rank solve, and
checks: identical iters_run, S_vec within 1e-12, objective history within 1e-11,
rotations within 1e-6 deg; CEMP's SVec bit-identical (per-edge arithmetic does not depend on the shard) and
CEMP+GCW rotations within 1e-6 deg.  With ``--full`` the same comparison also runs at cfg-4 size (n=10000, p=0.1,
1.5e8 slots, inputs from the device generator).  Prints MULTI_GPU_OK on rank 0 when all ranks agree.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import desc_b200                                   # noqa: E402
from desc_b200 import dist as ddist                # noqa: E402
from oracle import desc_oracle as O                # noqa: E402  (generator + metric only)


def solve(Ind, R, rank, world, nccl_id, local, iters, lr, n_sample):
    with desc_b200.Solver(Ind, R, device=local, rank=rank, world=world, nccl_id=nccl_id) as s:
        info = s.build_incidence(n_sample=n_sample, seed=5)
        s.cycle_inconsistency()
        S, hist, k = s.pgd(iters, desc_b200.ConstantStepSize(lr))
        Rot = s.gcw()
        # SURVEY 8(f) #3 on the same sharded incidence: CEMP (edge-sharded reweighting + all-gather) and CEMP+GCW
        cemp = s.cemp(4, [1.0, 4.0, 16.0])
        Rc = s.cemp_gcw()
        # stages that run replicated on every rank: LAA refinement (DESC.m:265-312), CEMP+MST, MPLS loop
        Rl, sc_l = s.refine(S_vec=S, R_init=Rot)
        Rm0 = s.mst_init()
        MP = dict(stop_threshold=1e-3, max_iter=20, reweighting=[16.0], thresholding=[0.95, 0.9, 0.85, 0.8],
                  cycle_info_ratio=1.0 / (np.arange(1, 21) + 1))
        Rm, sc_m = s.mpls_refine(MP)
        return dict(info=info, S=S, hist=hist, k=k, R=Rot, S0=s.S0(), w=s.w(), cemp=cemp, Rc=Rc, Rl=Rl, sc_l=sc_l,
                    Rm0=Rm0, Rm=Rm, sc_m=sc_m)


def main():
    rank, world, local = ddist.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for case, (n, p, q, sigma, iters, lr, ns) in enumerate([(300, 0.4, 0.2, 0.1, 40, 0.01, 0), (120, 0.5, 0.1, 0.0, 120, 1.0, 11),
                                                             (500, 0.2, 0.3, 0.05, 25, 0.05, -1)]):
        mo = O.uniform_topology(n, p, q, sigma, rng=100 + case)
        nccl_id = ddist.exchange_nccl_id(desc_b200.nccl_unique_id, device="cuda")
        one = solve(mo["Ind"], mo["RijMat"], 0, 1, None, local, iters, lr, ns)
        many = solve(mo["Ind"], mo["RijMat"], rank, world, nccl_id, local, iters, lr, ns)
        dS = float(np.max(np.abs(one["S"] - many["S"])))
        dh = float(np.max(np.abs(one["hist"][:, 1] - many["hist"][:, 1]) / np.maximum(np.abs(one["hist"][:, 1]), 1e-9)))
        ang = float(O.aligned_angle_deg(one["R"], many["R"]).mean())
        # local slots of the sharded run are a contiguous piece of the single-rank arrays
        a = many["info"]
        rowptr, _ = (None, None)
        dC = float(np.max(np.abs(one["cemp"] - many["cemp"])))
        angC = float(O.aligned_angle_deg(one["Rc"], many["Rc"]).mean())
        good = (one["k"] == many["k"] and dS <= 1e-12 and dh <= 1e-11 and ang <= 1e-6 and
                a["m_cycle"] == one["info"]["m_cycle"] and dC == 0.0 and angC <= 1e-6)
        angL = float(O.aligned_angle_deg(one["Rl"], many["Rl"]).mean())
        angM = float(O.aligned_angle_deg(one["Rm"], many["Rm"]).mean())
        good = (good and one["sc_l"].size == many["sc_l"].size and one["sc_m"].size == many["sc_m"].size and
                angL <= 1e-6 and angM <= 1e-6 and np.array_equal(one["Rm0"], many["Rm0"]) and
                float(np.max(np.abs(one["sc_m"] - many["sc_m"]), initial=0.0)) <= 1e-9)
        # S0 / w of the shard against the matching slice of the single-rank result
        with desc_b200.Solver(mo["Ind"], mo["RijMat"], device=local) as s1:
            s1.build_incidence(n_sample=ns, seed=5)
            rp, _ = s1.incidence()
        sl = slice(int(rp[a["edge_begin"]]), int(rp[a["edge_end"]]))
        good = good and np.array_equal(one["S0"][sl], many["S0"]) and float(np.max(np.abs(one["w"][sl] - many["w"]), initial=0.0)) <= 1e-12
        print("rank %d case %d: iters %d/%d dS=%.2e dobj=%.2e dR=%.2e deg dCEMP=%.1e dR_cemp=%.2e dR_laa=%.1e dR_mpls=%.1e shard=[%d,%d) slots=%d %s" % (
            rank, case, one["k"], many["k"], dS, dh, ang, dC, angC, angL, angM, a["edge_begin"], a["edge_end"], a["local_slots"],
            "ok" if good else "MISMATCH"), flush=True)
        ok = ok and good
    if "--full" in sys.argv:
        # cfg-4 size (n=10000, p=0.1: 5.0e6 edges, 1.5e8 slots) through the same comparison: N-rank == 1-rank
        with desc_b200.Uniform_Topology(10000, 0.1, 0.2, 0.1, "uniform", seed=0, device=local, on_device=True) as mo:
            nccl_id = ddist.exchange_nccl_id(desc_b200.nccl_unique_id, device="cuda")
            res = []
            for (rk, wd, nid) in ((0, 1, None), (rank, world, nccl_id)):
                with desc_b200.Solver(mo.Ind, mo.RijMat, n=mo.n, device=local, rank=rk, world=wd, nccl_id=nid) as s:
                    info = s.build_incidence(n_sample=0, seed=1)
                    s.cycle_inconsistency()
                    S, hist, k = s.pgd(12, desc_b200.ConstantStepSize(0.01))
                    res.append(dict(info=info, S=S, hist=hist, k=k, R=s.gcw()))
            one, many = res
            dS = float(np.max(np.abs(one["S"] - many["S"]) / np.maximum(np.abs(one["S"]), 1e-12)))
            dh = float(np.max(np.abs(one["hist"][:, 1] - many["hist"][:, 1]) / np.maximum(np.abs(one["hist"][:, 1]), 1e-9)))
            ang = float(O.aligned_angle_deg(one["R"], many["R"]).mean())
            good = one["k"] == many["k"] and dS <= 1e-10 and dh <= 1e-11 and ang <= 1e-6
            print("rank %d cfg4: iters %d/%d rel dS=%.2e dobj=%.2e dR=%.2e deg slots=%d %s" % (
                rank, one["k"], many["k"], dS, dh, ang, many["info"]["local_slots"], "ok" if good else "MISMATCH"), flush=True)
            ok = ok and good
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if int(t.item()) == 1 else "MULTI_GPU_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
