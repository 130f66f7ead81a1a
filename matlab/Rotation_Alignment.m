function [R_out, R_align, mean_error, median_error] = Rotation_Alignment(R_est, R_gt)
% Drop-in for Utils/Rotation_Alignment.m:13 on the GPU (errors in degrees).
    out = desc_b200_mex('align', double(R_est), double(R_gt));
    R_out = out.R_out; R_align = out.R_align;
    mean_error = out.mean_error; median_error = out.median_error;
end
