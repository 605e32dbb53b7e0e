function [aziEst, eleEst] = digitalBF(numDets, radarEstParams, Ra)
%DIGITALBF Drop-in for sensing.estimation.doaEstimation.digitalBF (+sensing/+estimation/+doaEstimation/digitalBF.m:1).
% UPA arrays: the reference's peak picker (tools.find2DPeaks) does not exist, so aziEst / eleEst come back empty there.
    cfg = sensing.estimation.isacDoaConfig(radarEstParams);
    [L, aziEst] = isac_doa_mex(cfg, 2, numDets, double(Ra)); %#ok<ASGLU>
    eleEst = NaN(size(aziEst));                        % ULA: no elevation estimate (music.m:104)
end
