function [R_est, R_init] = MPLS(Ind, RijMat, CEMP_parameters, MPLS_parameters)
% Drop-in for Algorithms/MPLS.m:28 on the GPU: CEMP (:66-150), minimum spanning tree + rotations along it
% (:152-195, R_init = CEMP+MST) and the MPLS reweighting loop (:198-256).
    seed = 0;
    if isfield(CEMP_parameters, 'seed'), seed = CEMP_parameters.seed; end
    out = desc_b200_mex('mpls', double(Ind), double(RijMat), CEMP_parameters.max_iter, ...
                        double(CEMP_parameters.reweighting), CEMP_parameters.nsample, seed, ...
                        MPLS_parameters.stop_threshold, MPLS_parameters.max_iter, double(MPLS_parameters.reweighting), ...
                        double(MPLS_parameters.thresholding), double(MPLS_parameters.cycle_info_ratio));
    disp('Rotation Initialized!'); disp('Start MPLS reweighting ...');        % MPLS.m:216-217
    for it = 1:numel(out.scores)
        fprintf('Iter %d: ||\x394R||= %f\n', it, out.scores(it));             % MPLS.m:248
    end
    disp('DONE!');
    R_est = out.R_est; R_init = out.R_init;
end
