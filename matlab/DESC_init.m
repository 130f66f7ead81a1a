function [R_est, S_vec] = DESC_init(Ind, RijMat, params)
% Drop-in for Algorithms/DESC_init.m:14 -- same signature, runs on the GPU through desc_b200_mex.
    out = desc_b200_run(Ind, RijMat, params, true);
    R_est = out.R_est;
    S_vec = out.S_vec;
end
