function out = desc_b200_run(Ind, RijMat, params, want_R)
% DESC_B200_RUN  shared body of the drop-in DESC_PGD / DESC_init / DESC shims (DESC.m:14-263 on the GPU).
    n_sample = 0; seed = 0;
    if isfield(params, 'n_sample'), n_sample = params.n_sample; end   % 0 = reference rule DESC.m:43
    if isfield(params, 'seed'),     seed = params.seed;         end   % sampler seed (replaces datasample's RNG)
    rule = desc_b200_rule(params.Gradient);
    plots = isfield(params, 'make_plots') && params.make_plots;
    disp('compute R cycle'); disp('S0Mat');                           % DESC.m:132,145
    if plots   % DESC.m:235-239: per-iteration S_vec error, GCW and alignment error, computed on the device
        out = desc_b200_mex('solve', double(Ind), double(RijMat), params.iters, rule, n_sample, seed, want_R, ...
                            double(params.ErrVec), double(params.R_orig));
    else
        out = desc_b200_mex('solve', double(Ind), double(RijMat), params.iters, rule, n_sample, seed, want_R);
    end
    disp('Initialization completed!');                                % DESC.m:160
    disp('Reweighting Procedure Started ...');                        % DESC.m:162
    for it = 1:out.iters_run                                          % DESC.m:241
        fprintf('iter %d: average change in S_vec %f, objective value: %f\n', it, out.hist(it,1), out.hist(it,2));
    end
    if isprop(params.Gradient, 't'), params.Gradient.t = out.t; end   % handle object: call counter advances
    if plots                                                          % DESC.m:315-344
        figure; tiledlayout(2,2);
        nexttile; plot(out.diag(:,1)); title('Convergence of Corruption Estimate Vector (S_vec, sampled)');
        xlabel('Iteration number'); ylabel('Average distance to true corruption');
        nexttile; plot(out.hist(:,2)); title('Convergence of Objective Function (sampled)');
        xlabel('Iteration number'); ylabel('Value of Objective Function');
        nexttile; plot(out.diag(:,2)); title('Convergence of Rotation Estimate, Mean (sampled)');
        xlabel('Iteration number'); ylabel('Mean Error in R estimate (degrees)'); ylim([0 inf]);
        nexttile; plot(out.diag(:,3)); title('Convergence of Rotation Estimate, Median (sampled)');
        xlabel('Iteration number'); ylabel('Median Error in R estimate (degrees)'); ylim([0 inf]);
    end
end
