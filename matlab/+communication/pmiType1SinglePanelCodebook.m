function W = pmiType1SinglePanelCodebook(reportConfig, nLayers)
%PMITYPE1SINGLEPANELCODEBOOK Drop-in for communication.pmiType1SinglePanelCodebook (+communication/pmiType1SinglePanelCodebook.m:1;
% consumer schedulerEntity.m:736-777).  Reproduces the gNB-side copy including its two deviations from the UE-side copy of
% dlPMISelect.m:853-1349 (ranks 3-4 with >= 16 ports: the i13 index is dropped; rank 2, mode 2, N2 > 1: floor(i2/4)).
% reportConfig: PanelDimensions [N1 N2], OverSamplingFactors [O1 O2], CodebookMode, CodebookSubsetRestriction, i2Restriction.
    pd = reportConfig.PanelDimensions; os = reportConfig.OverSamplingFactors;
    nPorts = 2*prod(pd);
    csr = []; if isfield(reportConfig, 'CodebookSubsetRestriction'), csr = reportConfig.CodebookSubsetRestriction; end
    i2r = []; if isfield(reportConfig, 'i2Restriction'), i2r = reportConfig.i2Restriction; end
    mode = 1; if isfield(reportConfig, 'CodebookMode'), mode = reportConfig.CodebookMode; end
    cfg = struct('nPanels', 0, 'nPorts', nPorts, 'N1', pd(1), 'N2', pd(2), 'O1', os(1), 'O2', os(2), 'codebookMode', mode, ...
                 'nSizeBWP', 1, 'nStartBWP', 0, 'subbandSize', 0, 'pmiSubband', 0, 'cqiSubband', 0, 'K', 12, 'L', 14, ...
                 'subsetRestriction', uint8(csr(:)), 'i2Restriction', uint8(i2r(:)), 'riRestriction', uint8(ones(8, 1)), ...
                 'reK', int32(zeros(0, 1)), 'reL', int32(zeros(0, 1)));
    W = isac_codebook_mex(cfg, nLayers, 1);
end
