function [model_out] = Nonuniform_Topology(n, p, p_node_crpt, p_edge_crpt, sigma_in, sigma_out, crpt_type, seed)
% Drop-in for Models/Nonuniform_Topology.m:26 on the GPU (counter-based draws of `seed`, default 0).
    if ~exist('crpt_type','var'), crpt_type = 'uniform'; end
    if ~exist('seed','var'),      seed = 0;              end
    switch crpt_type
        case 'uniform',         kind = 2;
        case 'self-consistent', kind = 3;
        case 'adv',             kind = 4;
        otherwise, error('DESC:b200', 'crpt_type must be uniform, self-consistent or adv');
    end
    model_out = desc_b200_mex('generate', kind, 0, n, 0, p, 0, sigma_in, sigma_out, p_node_crpt, p_edge_crpt, seed);
end
