function rule = desc_b200_rule(g)
% DESC_B200_RULE  params.Gradient (handle object of the reference) -> plain struct for the MEX gateway.
%   Utils/ConstantStepSize.m, Utils/PiecewiseStepSize.m, Utils/HybridGradient.m
    rule = struct('kind',0,'strategy',0,'lr',0,'decay_interval',1,'beta_1',0,'beta_2',0,'t',0);
    switch class(g)
        case 'ConstantStepSize'
            rule.kind = 0; rule.lr = g.learning_rate;
        case 'PiecewiseStepSize'
            rule.kind = 1; rule.lr = g.learning_rate; rule.decay_interval = g.decay_interval; rule.t = g.t;
        case 'HybridGradient'
            rule.kind = 2; rule.lr = g.lr; rule.beta_1 = g.beta_1; rule.beta_2 = g.beta_2;
            rule.decay_interval = g.decay_interval; rule.t = g.t; rule.strategy = g.strategy;
            if g.t ~= 0 && g.strategy == 0
                error('DESC:b200', ['HybridGradient with t>0: the Adam moments m_t/v_t live on the device ' ...
                      'inside one solve; pass a fresh object (t==0)']);
            end
        otherwise
            error('DESC:b200', 'params.Gradient of class %s is not supported on the device', class(g));
    end
end
