function R_est = Spectral(Ind, RijMat)
% Drop-in for Algorithms/Spectral.m:15 on the GPU.
    R_est = desc_b200_mex('spectral', double(Ind), double(RijMat));
end
