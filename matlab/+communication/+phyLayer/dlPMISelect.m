function [PMISet, info] = dlPMISelect(carrier, csirs, reportConfig, nLayers, H, varargin)
%DLPMISELECT Drop-in for communication.phyLayer.dlPMISelect (+communication/+phyLayer/dlPMISelect.m:1).
% Argument validation (nr5g:dlPMISelect:* identifiers) stays in MATLAB exactly as in the reference's validateInputs
% (:511-851, unchanged copy expected on the path as communication.phyLayer.validateDLPMIInputs); the RE list it
% produces and the validated configuration are handed to isac_dl_pmi_mex, which marshals onto
% isac_pmi_plan_create / isac_dl_pmi_select_dev / isac_dl_pmi_collect / isac_dl_pmi_get_info.
    if nargin == 6, nVar = varargin{1}; else, nVar = 1e-10; end
    [rc, csirsIndSubs, nVar] = communication.phyLayer.validateDLPMIInputs(carrier, csirs, reportConfig, nLayers, H, nVar);
    bwpStart = rc.NStartBWP - carrier.NStartGrid;
    k = csirsIndSubs(:,1); l = csirsIndSubs(:,2);
    keep = (k >= bwpStart*12 + 1) & (k <= (bwpStart + rc.NSizeBWP)*12);
    cfg = struct('nPorts', csirs.NumCSIRSPorts(1), 'N1', rc.PanelDimensions(1), 'N2', rc.PanelDimensions(2), ...
                 'O1', rc.OverSamplingFactors(1), 'O2', rc.OverSamplingFactors(2), 'codebookMode', rc.CodebookMode, ...
                 'nSizeBWP', rc.NSizeBWP, 'nStartBWP', rc.NStartBWP, 'subbandSize', max([rc.SubbandSize 0]), ...
                 'pmiSubband', strcmpi(rc.PMIMode, 'Subband'), 'cqiSubband', 0, 'K', carrier.NSizeGrid*12, ...
                 'L', carrier.SymbolsPerSlot, 'subsetRestriction', uint8(rc.CodebookSubsetRestriction), ...
                 'i2Restriction', uint8(rc.i2Restriction), 'reK', int32(k(keep) - bwpStart*12), 'reL', int32(l(keep)));
    [i1, i2, sinrPerRE, sinrPerSubband, W, reK, reL] = isac_dl_pmi_mex(cfg, nLayers, single(H), double(nVar));
    PMISet.i1 = i1(:).'; PMISet.i2 = i2(:).';
    % scatter the compact [nRE x nLayers x ...] array into the reference's K x L x ... NaN grid (dlPMISelect.m:384,421)
    sz = size(sinrPerRE);
    info.SINRPerRE = NaN([rc.NSizeBWP*12, carrier.SymbolsPerSlot, sz(2:end)]);
    for e = 1:numel(reK), info.SINRPerRE(reK(e), reL(e), :) = reshape(sinrPerRE(e, :), 1, 1, []); end
    info.SINRPerSubband = sinrPerSubband;
    info.W = W;
end
