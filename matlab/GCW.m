function R_est = GCW(Ind, AdjMat, RijMat, S_vec) %#ok<INUSL>
% Drop-in for Utils/GCW.m:1.  AdjMat is only a mask in the reference (GCW.m:20) and is implied by Ind.
    R_est = desc_b200_mex('gcw', double(Ind), double(RijMat), double(S_vec(:)));
end
