function [PMISet, info] = dlPMISelect(carrier, csirs, reportConfig, nLayers, H, varargin)
%DLPMISELECT Drop-in for communication.phyLayer.dlPMISelect (+communication/+phyLayer/dlPMISelect.m:1).
% isac_dl_pmi_mex marshals onto isac_pmi_plan_create / isac_dl_pmi_select_dev / isac_dl_pmi_collect / isac_dl_pmi_get_info.
    if nargin == 6, nVar = varargin{1}; else, nVar = 1e-10; end
    [cfg, rc] = communication.phyLayer.isacCsiConfig(carrier, csirs, reportConfig, nLayers, H, nVar);
    [i1, i2, sinrPerRE, sinrPerSubband, W, reK, reL, mp] = isac_dl_pmi_mex(cfg, nLayers, single(H), double(nVar));
    PMISet.i1 = i1(:).'; PMISet.i2 = i2(:).';
    if cfg.nPanels >= 2
        % Type1MultiPanel: un-flatten the index set [i20 i21 i22 | i11 i12 i13 i141 i142 i143] (dlPMISelect.m:455-457, :489)
        sz = size(sinrPerRE); full = [mp(1:3) sz(4) sz(5) mp(4:7)];
        sinrPerRE = reshape(sinrPerRE, [sz(1:2) full]);
        sinrPerSubband = reshape(sinrPerSubband, [size(sinrPerSubband, 1) sz(2) full]);
        W = reshape(W, [size(W, 1) size(W, 2) full]);
        PMISet.i1 = NaN(1, 6); PMISet.i2 = NaN(3, numel(i2));
        if ~any(isnan(i1))
            [a, b, c, d] = ind2sub(mp(4:7), i1(3)); PMISet.i1 = [i1(1) i1(2) a b c d];
        end
        for sb = find(~isnan(i2(:).'))
            [a, b, c] = ind2sub(mp(1:3), i2(sb)); PMISet.i2(:, sb) = [a; b; c];
        end
    end
    % scatter the compact [nRE x nLayers x ...] array into the reference's K x L x ... NaN grid (dlPMISelect.m:384,421)
    sz = size(sinrPerRE);
    info.SINRPerRE = NaN([rc.NSizeBWP*12, carrier.SymbolsPerSlot, sz(2:end)]);
    for e = 1:numel(reK), info.SINRPerRE(reK(e), reL(e), :) = reshape(sinrPerRE(e, :), 1, 1, []); end
    info.SINRPerSubband = sinrPerSubband;
    info.W = W;
end
