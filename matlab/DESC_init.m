function [R_est, S_vec] = DESC_init(Ind, RijMat, params)
% Drop-in for Algorithms/DESC_init.m:14 -- same signature, runs on the GPU through desc_b200_mex.
    out = desc_b200_run(Ind, RijMat, params, true);
    R_est = out.R_est;
    S_vec = out.S_vec;
    if isfield(params, 'make_plots') && params.make_plots            % DESC_init.m:261-262
        dlmwrite('linear_convergence_rotation_error.csv', out.diag(:,2)', 'delimiter', ',', '-append');
        dlmwrite('linear_convergence_svec_error.csv', out.diag(:,1)', 'delimiter', ',', '-append');
    end
end
