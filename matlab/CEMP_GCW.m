function R_est = CEMP_GCW(Ind, RijMat, CEMP_parameters)
% Drop-in for Algorithms/CEMP_GCW.m:25 on the GPU: CEMP (CEMP_GCW.m:25-125) followed by the weighted
% spectral recovery with weights 1./(SVec+1e-8) (CEMP_GCW.m:127-159).
    seed = 0;
    if isfield(CEMP_parameters, 'seed'), seed = CEMP_parameters.seed; end
    out = desc_b200_mex('cemp', double(Ind), double(RijMat), CEMP_parameters.max_iter, ...
                        double(CEMP_parameters.reweighting), CEMP_parameters.nsample, seed, true);
    for iter = 1:CEMP_parameters.max_iter
        fprintf('Reweighting Iteration %d Completed!\n', iter);                          % CEMP_GCW.m:124
    end
    disp('Completed!');
    R_est = out.R_est;
end
