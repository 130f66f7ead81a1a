function out = desc_b200_run(Ind, RijMat, params, want_R)
% DESC_B200_RUN  shared body of the drop-in DESC_PGD / DESC_init / DESC shims (DESC.m:14-263 on the GPU).
    if isfield(params, 'make_plots') && params.make_plots
        error('DESC:b200', ['params.make_plots=true (per-iteration GCW diagnostics, DESC.m:235-239,315-344) ' ...
              'is not part of the device hot path; set make_plots=false']);
    end
    n_sample = 0; seed = 0;
    if isfield(params, 'n_sample'), n_sample = params.n_sample; end   % 0 = reference rule DESC.m:43
    if isfield(params, 'seed'),     seed = params.seed;         end   % sampler seed (replaces datasample's RNG)
    rule = desc_b200_rule(params.Gradient);
    disp('compute R cycle'); disp('S0Mat');                           % DESC.m:132,145
    out = desc_b200_mex('solve', double(Ind), double(RijMat), params.iters, rule, n_sample, seed, want_R);
    disp('Initialization completed!');                                % DESC.m:160
    disp('Reweighting Procedure Started ...');                        % DESC.m:162
    for it = 1:out.iters_run                                          % DESC.m:241
        fprintf('iter %d: average change in S_vec %f, objective value: %f\n', it, out.hist(it,1), out.hist(it,2));
    end
    if isprop(params.Gradient, 't'), params.Gradient.t = out.t; end   % handle object: call counter advances
end
