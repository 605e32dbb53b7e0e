function [PMISet, info] = dlPMISelect(carrier, csirs, reportConfig, nLayers, H, varargin)
%DLPMISELECT Drop-in for communication.phyLayer.dlPMISelect (+communication/+phyLayer/dlPMISelect.m:1).
% isac_dl_pmi_mex marshals onto isac_pmi_plan_create / isac_dl_pmi_select_dev / isac_dl_pmi_collect / isac_dl_pmi_get_info.
    if nargin == 6, nVar = varargin{1}; else, nVar = 1e-10; end
    [cfg, rc] = communication.phyLayer.isacCsiConfig(carrier, csirs, reportConfig, nLayers, H, nVar);
    [i1, i2, sinrPerRE, sinrPerSubband, W, reK, reL] = isac_dl_pmi_mex(cfg, nLayers, single(H), double(nVar));
    PMISet.i1 = i1(:).'; PMISet.i2 = i2(:).';
    % scatter the compact [nRE x nLayers x ...] array into the reference's K x L x ... NaN grid (dlPMISelect.m:384,421)
    sz = size(sinrPerRE);
    info.SINRPerRE = NaN([rc.NSizeBWP*12, carrier.SymbolsPerSlot, sz(2:end)]);
    for e = 1:numel(reK), info.SINRPerRE(reK(e), reL(e), :) = reshape(sinrPerRE(e, :), 1, 1, []); end
    info.SINRPerSubband = sinrPerSubband;
    info.W = W;
end
