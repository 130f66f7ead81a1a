function R_ref = desc_b200_refine(Ind, RijMat, R0, s)
% DESC_B200_REFINE  host-side stand-in for the refinement stage of the reference (DESC.m:265-312):
% iteratively re-weighted Lie-algebraic averaging started from the GPU result (R0, s).
% Uses the reference's Utils on the MATLAB path: Build_Amatrix, R2Q, Weighted_LAA, q2R.
    cfg = struct('tol', 1e-3, 'max_it', 100, 'q_floor', 0.8, 'q_step', 0.05, ...
                 'w_hi', 1e4, 'w_lo', 1e-4, 'expo', 0.75);            % constants of DESC.m:272-280
    edges = Ind.';
    A     = Build_Amatrix(edges);
    q_abs = R2Q(R0);
    q_rel = R2Q(permute(RijMat, [2 1 3]));                             % LAA estimates R' (DESC.m:265)
    s     = s(:);
    level = 1;
    wts   = reweight(s, level, cfg);
    disp('Rotation Initialized!'); disp('Start DESC refinement ...');
    it = 1; delta = inf;
    while delta > cfg.tol && it < cfg.max_it
        [q_abs, W, B, delta] = Weighted_LAA(edges, q_abs, q_rel, A, wts);
        resid = sqrt(sum((A*W(2:end,2:4) - B).^2, 2)) / pi;            % normalised edge residuals
        mix   = 1/(it+1);
        blend = (1-mix)*resid + mix*s;                                 % DESC.m:293
        level = max(cfg.q_floor, level - cfg.q_step);
        wts   = reweight(blend, level, cfg);
        fprintf('Iter %d: ||\x394R||= %f\n', it, delta);
        it = it + 1;
    end
    R_ref = zeros(3, 3, size(q_abs,1));
    for v = 1:size(q_abs,1)
        R_ref(:,:,v) = q2R(q_abs(v,:));
    end
    disp('DONE!');
end

function w = reweight(x, level, cfg)
% weights x^-0.75 clipped from above; edges beyond the `level` quantile get the floor weight
    w = min(1 ./ (x.^cfg.expo), cfg.w_hi);
    w(x > quantile(x, level)) = cfg.w_lo;
end
