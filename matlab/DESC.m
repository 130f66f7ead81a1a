function [R_est, R_init, S_vec] = DESC(Ind, RijMat, params)
% Drop-in for Algorithms/DESC.m:14 -- what Demo/compare_algorithms.m:72 calls.
% Stages 1-4 of the reference (DESC.m:14-263: incidence, d_ijk, PGD, GCW) run on the GPU.
% Stage 5, the Lie-algebraic refinement (DESC.m:265-312), is delegated to desc_b200_refine, which
% drives the reference's own host utilities (Utils/Weighted_LAA.m etc.) until the device version
% (SURVEY 8(f) "next #1") lands.
    out    = desc_b200_run(Ind, RijMat, params, true);
    R_init = out.R_est;
    S_vec  = out.S_vec;
    R_est  = desc_b200_refine(Ind, RijMat, R_init, S_vec);
end
