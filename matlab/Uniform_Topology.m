function [model_out] = Uniform_Topology(n, p, q, sigma, model, seed)
% Drop-in for Models/Uniform_Topology.m:24 on the GPU.  MATLAB's global RNG stream is replaced by
% counter-based draws of `seed` (optional 6th argument, default 0); any model other than 'uniform' means
% self-consistent corruption, as in the reference (:76,83).
    if ~exist('model','var'), model = 'uniform'; end
    if ~exist('seed','var'),  seed = 0;          end
    kind = double(~strcmp(model, 'uniform'));
    model_out = desc_b200_mex('generate', kind, 0, n, 0, p, q, sigma, 0, 0, 0, seed);
end
