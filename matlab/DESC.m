function [R_est, R_init, S_vec] = DESC(Ind, RijMat, params)
% Drop-in for Algorithms/DESC.m:14 -- what Demo/compare_algorithms.m:72 calls.
% Stages 1-4 of the reference (DESC.m:14-263: incidence, d_ijk, PGD, GCW) and stage 5, the weighted
% Lie-algebraic refinement (DESC.m:265-312), all run on the GPU; this file only prints the
% reference's progress lines.
    out    = desc_b200_run(Ind, RijMat, params, 2);   % 2: GCW + refinement on one handle, one upload of the inputs
    R_init = out.R_est;
    S_vec  = out.S_vec;
    disp('Rotation Initialized!'); disp('Start DESC refinement ...');      % DESC.m:283-284
    for it = 1:numel(out.scores)
        fprintf('Iter %d: ||\x394R||= %f\n', it, out.scores(it));          % DESC.m:305
    end
    R_est = out.R_refined;
    disp('DONE!');                                                          % DESC.m:313
end
