function sinr = precodedSINR(H, sigma, W)
%PRECODEDSINR Drop-in for communication.phyLayer.precodedSINR (+communication/+phyLayer/precodedSINR.m:11).
% H may carry a third dimension: the REs of a batch that share W are evaluated in one launch (the reference calls the
% function once per RE and TPMI, pmiSelect.m:44-53).
    sinr = isac_precoded_sinr_mex(double(H), double(sigma), double(W));
end
