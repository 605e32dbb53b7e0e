function [cfg, rc] = isacCsiConfig(carrier, csirs, reportConfig, nLayers, H, nVar)
%ISACCSICONFIG Configuration struct of the CSI gateways (isac_dl_pmi_mex, isac_csi_report_mex) from the validated
% reportConfig.  Argument validation (nr5g:dlPMISelect:* identifiers) stays in MATLAB exactly as in the reference's
% validateInputs (dlPMISelect.m:511-851; an unchanged copy is expected on the path as
% communication.phyLayer.validateDLPMIInputs); the RE list it produces is kept to the BWP as dlPMISelect.m:352-356 does.
    [rc, csirsIndSubs] = communication.phyLayer.validateDLPMIInputs(carrier, csirs, reportConfig, nLayers, H, nVar);
    bwpStart = rc.NStartBWP - carrier.NStartGrid;
    k = csirsIndSubs(:,1); l = csirsIndSubs(:,2);
    keep = (k >= bwpStart*12 + 1) & (k <= (bwpStart + rc.NSizeBWP)*12);
    rir = ones(1, 8);
    if isfield(rc, 'RIRestriction') && ~isempty(rc.RIRestriction), rir(1:numel(rc.RIRestriction)) = rc.RIRestriction; end
    cqiSubband = isfield(rc, 'CQIMode') && strcmpi(rc.CQIMode, 'Subband');
    nPanels = 0; pd = rc.PanelDimensions;
    if isfield(rc, 'CodebookType') && strcmpi(rc.CodebookType, 'Type1MultiPanel')   % PanelDimensions = [Ng N1 N2] (:629-644)
        nPanels = pd(1); pd = pd(2:3);
    end
    cfg = struct('nPanels', nPanels, 'nPorts', csirs.NumCSIRSPorts(1), 'N1', pd(1), 'N2', pd(2), ...
                 'O1', rc.OverSamplingFactors(1), 'O2', rc.OverSamplingFactors(2), 'codebookMode', rc.CodebookMode, ...
                 'nSizeBWP', rc.NSizeBWP, 'nStartBWP', rc.NStartBWP, 'subbandSize', max([rc.SubbandSize 0]), ...
                 'pmiSubband', double(strcmpi(rc.PMIMode, 'Subband')), 'cqiSubband', double(cqiSubband), ...
                 'K', carrier.NSizeGrid*12, 'L', carrier.SymbolsPerSlot, ...
                 'subsetRestriction', uint8(rc.CodebookSubsetRestriction(:)), 'i2Restriction', uint8(rc.i2Restriction(:)), ...
                 'riRestriction', uint8(rir(:)), 'reK', int32(k(keep) - bwpStart*12), 'reL', int32(l(keep)));
end
