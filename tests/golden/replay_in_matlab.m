% replay_in_matlab.m -- pins the oracle against an EXECUTION of the unmodified reference.
%
% The build image has neither MATLAB nor Octave, so the golden fixtures in this directory come from the repo's
% own restatement (tests/golden/make_golden.py).  This script closes that gap for whoever has MATLAB (R2019b or
% newer: the reference uses max(...,[],'all')) with the Statistics Toolbox (datasample): it runs the reference's
% own Algorithms/DESC_init.m, untouched, on the fixture `uniform_n50_nosample.mat` and prints the comparisons the
% parity tests make.  On that fixture no edge has more common neighbours than n_sample (max co-degree 16 < 30), so
% datasample (DESC.m:84 / DESC_init.m) is never reached and the reference is fully deterministic: every number below
% must agree to rounding.
%
% Usage (from the root of a checkout of ColeWyeth/DESC, with this repo at <repo>):
%     addpath Utils Models Algorithms
%     run('<repo>/tests/golden/replay_in_matlab.m')
%
% Expected output: every line ends in "OK".  Tolerances are the north_star's: S_vec and the objective history within
% 1e-10 relative, identical iteration count, rotations within 1e-6 degrees (mean) after gauge alignment with the
% reference's own Utils/Rotation_Alignment.m.  Please report the MATLAB release and the printed numbers in
% INTEGRATION.md ("MATLAB replay") if you run it.
here = fileparts(mfilename('fullpath'));
g = load(fullfile(here, 'uniform_n50_nosample.mat'));

params.iters = double(g.iters);
params.learning_rate = g.rule(1);
params.Gradient = ConstantStepSize(g.rule(1));   % Utils/ConstantStepSize.m
params.make_plots = false;
params.ErrVec = g.ErrVec;
params.R_orig = g.R_orig;

% capture the per-iteration objective the reference prints (DESC_init.m: 'iter %d: average change ... objective value')
txt = evalc('[R_est, S_vec] = DESC_init(g.Ind, g.RijMat, params);');
tok = regexp(txt, 'iter (\d+): average change in S_vec ([\d\.eE+-]+), objective value: ([\d\.eE+-]+)', 'tokens');
iters_run = numel(tok);
obj_printed = cellfun(@(t) str2double(t{3}), tok);

relerr = @(a, b, fl) max(abs(a(:) - b(:)) ./ max(abs(b(:)), fl));
chk = @(name, v, tol) fprintf('%-52s %.3e  (tol %.0e)  %s\n', name, v, tol, char((v <= tol) * 'OK ' + (v > tol) * 'BAD'));

fprintf('MATLAB %s\n', version);
fprintf('%-52s %d vs %d  %s\n', 'iterations run (DESC_init.m early stop)', iters_run, double(g.iters_run), ...
        char((iters_run == double(g.iters_run)) * 'OK ' + (iters_run ~= double(g.iters_run)) * 'BAD'));
chk('S_vec: max relative error (floor 1e-12)', relerr(S_vec, g.S_vec, 1e-12), 1e-10);
% the reference prints the objective with %f (6 decimals): compare at that resolution
chk('objective history as printed (absolute, %f)', max(abs(obj_printed(:) - g.hist(1:iters_run, 2))), 1e-6);
[~, ~, mean_err, ~] = Rotation_Alignment(R_est, g.R_est);       % Utils/Rotation_Alignment.m
chk('R_est vs fixture: mean angle after alignment (deg)', mean_err, 1e-6);
[~, ~, e_ref, ~] = Rotation_Alignment(R_est, g.R_orig);
[~, ~, e_fix, ~] = Rotation_Alignment(g.R_est, g.R_orig);
chk('error vs ground truth: |reference - fixture| (deg)', abs(e_ref - e_fix), 1e-6);
mask_ref = S_vec > quantile(S_vec, 0.8);                        % the down-weighting mask of DESC.m:276-282
mask_fix = g.S_vec > quantile(g.S_vec, 0.8);
fprintf('%-52s %d differing edges  %s\n', 'classification mask S_vec > quantile(S_vec, 0.8)', nnz(mask_ref ~= mask_fix), ...
        char((nnz(mask_ref ~= mask_fix) == 0) * 'OK ' + (nnz(mask_ref ~= mask_fix) > 0) * 'BAD'));
