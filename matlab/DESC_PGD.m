function [S_vec] = DESC_PGD(Ind, RijMat, params)
% Drop-in for Algorithms/DESC_PGD.m:14 -- same signature, runs on the GPU through desc_b200_mex.
    out = desc_b200_run(Ind, RijMat, params, false);
    S_vec = out.S_vec;
end
