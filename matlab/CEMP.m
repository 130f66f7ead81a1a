function [SVec] = CEMP(Ind, RijMat, CEMP_parameters)
% Drop-in for Algorithms/CEMP.m:25 on the GPU (same incidence + d_ijk kernels as DESC).
% The reference draws nsample apices per edge WITH replacement (CEMP.m:63); the device sampler keeps
% nsample distinct common neighbours (all of them when an edge has fewer).
    seed = 0;
    if isfield(CEMP_parameters, 'seed'), seed = CEMP_parameters.seed; end
    disp('sampling 3-cycles'); disp('Sampling Finished!'); disp('Initializing');        % CEMP.m:45,66,67
    out = desc_b200_mex('cemp', double(Ind), double(RijMat), CEMP_parameters.max_iter, ...
                        double(CEMP_parameters.reweighting), CEMP_parameters.nsample, seed, false);
    disp('Initialization completed!'); disp('Reweighting Procedure Started ...');       % CEMP.m:103-104
    for iter = 1:CEMP_parameters.max_iter
        fprintf('Reweighting Iteration %d Completed!\n', iter);                          % CEMP.m:126
    end
    disp('Completed!');
    SVec = out.SVec;
end
