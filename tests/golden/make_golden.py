"""Generates the golden fixtures in this directory.  Run from the repo root:

    python tests/golden/make_golden.py

The reference (ColeWyeth/DESC) is MATLAB-only and ships no golden vectors, and no MATLAB/Octave
exists in the build image, so these vectors are produced by ``oracle/desc_literal.py`` -- the
loop-for-loop restatement of Algorithms/DESC.m:14-263 and Utils/GCW.m with the reference's dense
data structures -- on graphs drawn by the restated generators (Models/Uniform_Topology.m,
Models/Nonuniform_Topology.m).  PARITY UNPINNED against an execution of the real reference; the
``.mat`` twin of every fixture lets a MATLAB user replay them through the real code (for the
``nosample`` fixture, where datasample is never called, the real reference is deterministic).

Each ``.npz`` holds the inputs (Ind, RijMat, R_orig, ErrVec, params) and the literal outputs
(incidence arrays, S0_long, wijk, S_vec, hist, iters_run, R_est of GCW) and, for the whole ``DESC()``
call, the refinement stage DESC.m:265-312 started from them (R_laa, laa_scores; oracle.laa_refine, a
statement-by-statement restatement of the loop and of Utils/Weighted_LAA.m).

``cemp_*.npz`` hold a with-replacement draw ``CoIndMat`` standing in for CEMP.m:63 (MATLAB's
``datasample`` stream cannot be restated, so the draw is part of the fixture), the literal CEMP.m / CEMP_GCW.m
outputs on it (``S0Mat``, ``SVec`` after every reweighting, ``R_est``), the alignment metric
(Utils/Rotation_Alignment.m) of ``R_est`` against ``R_orig``, and the MPLS.m:152-256 stages on the same draw
(``R_mst``, ``mst_edges``, ``R_mpls``, ``mpls_scores`` with the demo's MPLS_parameters).  A MATLAB user can replay them by replacing the
``datasample`` call with the stored ``CoIndMat`` column.

    python tests/golden/make_golden.py            # everything
    python tests/golden/make_golden.py cemp       # only the fixtures whose name starts with "cemp"
    python tests/golden/make_golden.py gen        # only the generator fixtures (counter-based draws)
"""
import os
import sys

import numpy as np
import scipy.io

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import desc_oracle as O          # noqa: E402
from oracle.desc_literal import desc_literal, cemp_literal, cemp_draw  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, generator, args, n_sample override (None = reference rule), sampler seed, rule, iters
    ("uniform_n60_sigma0", "uniform", dict(n=60, p=0.5, q=0.2, sigma=0.0, model="uniform"), 10, 3,
     ("const", 0.01), 100),
    ("uniform_n70_noise", "uniform", dict(n=70, p=0.5, q=0.3, sigma=0.1, model="uniform"), 12, 5,
     ("const", 0.01), 60),
    ("uniform_n50_nosample", "uniform", dict(n=50, p=0.45, q=0.2, sigma=0.05, model="self-consistent"), None, 0,
     ("const", 0.05), 40),
    ("nonuniform_n64_adv", "nonuniform", dict(n=64, p=0.5, p_node_crpt=0.3, p_edge_crpt=0.5, sigma_in=0.05,
                                               sigma_out=0.1, crpt_type="adv"), 9, 11, ("const", 1.0), 30),
    ("uniform_n48_piecewise", "uniform", dict(n=48, p=0.6, q=0.25, sigma=0.02, model="uniform"), 8, 2,
     ("piecewise", 0.05, 10), 45),
    ("uniform_n48_adam", "uniform", dict(n=48, p=0.6, q=0.25, sigma=0.02, model="uniform"), 8, 2,
     ("adam", 0.002, 0.9, 0.999, 20), 35),
]


CEMP_CASES = [
    # name, generator args, CEMP_parameters, draw seed
    ("cemp_uniform_n40", dict(n=40, p=0.5, q=0.25, sigma=0.05, model="uniform"),
     dict(max_iter=6, reweighting=[1.0, 2.0, 4.0, 8.0, 16.0, 32.0], nsample=12), 5),
    ("cemp_uniform_n36_shortbeta", dict(n=36, p=0.3, q=0.2, sigma=0.0, model="uniform"),
     dict(max_iter=5, reweighting=[1.0, 3.0], nsample=7), 6),
]


def make_cemp(prefix):
    for idx, (name, a, P, dseed) in enumerate(CEMP_CASES):
        if not name.startswith(prefix):
            continue
        mo = O.uniform_topology(a["n"], a["p"], a["q"], a["sigma"], a["model"], rng=np.random.default_rng(2000 + idx))
        Co, ptr, apex = cemp_draw(mo["Ind"], P["nsample"], dseed)
        SVec, ex = cemp_literal(mo["Ind"], mo["RijMat"], P, Co, return_gcw=True)
        _, R_align, mean_err, med_err = O.rotation_alignment(ex["R_est"], mo["R_orig"])
        # MPLS.m:152-256 on the same draw and the literal SVec (oracle.mst_init / mpls_refine)
        MP = dict(stop_threshold=1e-3, max_iter=100, reweighting=[P["reweighting"][-1]],
                  thresholding=[0.95, 0.9, 0.85, 0.8], cycle_info_ratio=1.0 / (np.arange(1, 101) + 1))   # compare_algorithms.m:36-40
        inc = O.cemp_incidence(mo["Ind"], cycles=(ptr, apex))
        S0 = O.cycle_inconsistency(inc, mo["RijMat"])
        R_mst, tree = O.mst_init(mo["Ind"], mo["RijMat"], SVec)
        R_mpls, mi = O.mpls_refine(mo["Ind"], mo["RijMat"], inc, S0, SVec, R_mst, MP, return_info=True)
        out = dict(R_mst=R_mst, mst_edges=tree + 1, R_mpls=R_mpls, mpls_scores=mi["scores"],
                   mpls_stop_threshold=np.float64(MP["stop_threshold"]), mpls_max_iter=np.int64(MP["max_iter"]),
                   mpls_reweighting=np.array(MP["reweighting"]), mpls_thresholding=np.array(MP["thresholding"]),
                   mpls_cycle_info_ratio=np.array(MP["cycle_info_ratio"]))
        out.update(Ind=mo["Ind"], RijMat=mo["RijMat"], R_orig=mo["R_orig"], ErrVec=mo["ErrVec"],
                   max_iter=np.int64(P["max_iter"]), reweighting=np.array(P["reweighting"], dtype=np.float64),
                   nsample=np.int64(P["nsample"]), CoIndMat=Co, cyc_ptr=ptr, cyc_apex=apex, S0Mat=ex["S0Mat"],
                   SVec=SVec, SVec_hist=np.array(ex["hist"]), R_est=ex["R_est"], R_align=R_align,
                   mean_error=np.float64(mean_err), median_error=np.float64(med_err))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        scipy.io.savemat(os.path.join(HERE, name + ".mat"), out, do_compression=True)
        print("%-26s n=%d m=%d slots=%d mean|SVec-ErrVec| %.4g  CEMP+GCW error %.4g / %.4g deg" % (
            name, a["n"], mo["Ind"].shape[0], apex.size, np.mean(np.abs(SVec - mo["ErrVec"])), mean_err, med_err))


def make_generators(prefix):
    """``gen_*.npz``: outputs of the counter-based generators (oracle/desc_models_ctr.py == csrc/gen.cu) for fixed
    seeds -- pins the draw specification (streams, mixing constants, Box-Muller pairing) against accidental change."""
    from oracle import desc_models_ctr as M
    cases = [("gen_uniform_n24", lambda: M.uniform_topology(24, 0.5, 0.3, 0.1, "uniform", seed=5)),
             ("gen_selfconsistent_n20", lambda: M.uniform_topology(20, 0.6, 0.4, 0.05, "self-consistent", seed=6)),
             ("gen_ring_n40_w6", lambda: M.uniform_topology(40, 0.5, 0.2, 0.1, "uniform", seed=7, ring=6)),
             ("gen_nonuniform_n24_adv", lambda: M.nonuniform_topology(24, 0.5, 0.4, 0.5, 0.05, 0.1, "adv", seed=8))]
    for name, fn in cases:
        if not name.startswith(prefix):
            continue
        mo = fn()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), Ind=mo["Ind"], RijMat=mo["RijMat"], R_orig=mo["R_orig"],
                            ErrVec=mo["ErrVec"], corrupted=mo["corrupted"])
        print("%-26s m=%d corrupted=%d" % (name, mo["Ind"].shape[0], int(mo["corrupted"].sum())))


def make_rule(spec):
    if spec[0] == "const":
        return O.ConstantStepSize(spec[1])
    if spec[0] == "piecewise":
        return O.PiecewiseStepSize(spec[1], spec[2])
    return O.HybridGradient(spec[1], spec[2], spec[3], spec[4])


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else ""
    make_cemp(prefix)
    make_generators(prefix)
    for idx, (name, gen, args, ns, seed, rule_spec, iters) in enumerate(CASES):
        if not name.startswith(prefix):
            continue
        rng = np.random.default_rng(1000 + idx)
        a = dict(args)
        n = a.pop("n")
        if gen == "uniform":
            mo = O.uniform_topology(n, a["p"], a["q"], a["sigma"], a["model"], rng=rng)
        else:
            mo = O.nonuniform_topology(n, a["p"], a["p_node_crpt"], a["p_edge_crpt"], a["sigma_in"],
                                       a["sigma_out"], a["crpt_type"], rng=rng)
        params = dict(iters=iters, Gradient=make_rule(rule_spec))
        R_est, S_vec, ex = desc_literal(mo["Ind"], mo["RijMat"], params, seed=seed, n_sample=ns)
        out = dict(
            Ind=mo["Ind"], RijMat=mo["RijMat"], R_orig=mo["R_orig"], ErrVec=mo["ErrVec"],
            n_sample_arg=np.int64(-1 if ns is None else ns), sampler_seed=np.int64(seed),
            rule=np.array(rule_spec[1:], dtype=np.float64), rule_kind=np.str_(rule_spec[0]), iters=np.int64(iters),
            n_sample=np.int64(ex["n_sample"]), cum_ind=ex["cum_ind"], CoDeg_pos_ind=ex["CoDeg_pos_ind"],
            Ind_jk=ex["Ind_jk"], Ind_ki=ex["Ind_ki"], IJK=ex["IJK"], IKJ=ex["IKJ"], JKI=ex["JKI"],
            S0_long=ex["S0_long"], wijk=ex["wijk"], S_vec=S_vec, hist=ex["hist"], iters_run=np.int64(ex["iters_run"]),
            R_est=R_est)
        R_laa, li = O.laa_refine(mo["Ind"], mo["RijMat"], S_vec, R_est, return_info=True)
        out["R_laa"] = R_laa
        out["laa_scores"] = li["scores"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        mat = {k: v for k, v in out.items() if k not in ("rule_kind",)}
        mat["rule_kind"] = rule_spec[0]
        scipy.io.savemat(os.path.join(HERE, name + ".mat"), mat, do_compression=True)
        print("%-26s n=%d m=%d m_cycle=%d n_sample=%d iters_run=%d obj %.6g -> %.6g" % (
            name, n, mo["Ind"].shape[0], ex["S0_long"].size, ex["n_sample"], ex["iters_run"], ex["hist"][0, 1],
            ex["hist"][-1, 1]))


if __name__ == "__main__":
    main()
