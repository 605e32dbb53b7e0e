function [rc, csirsIndSubs, nVar] = validateDLPMIInputs(carrier, csirs, reportConfig, nLayers, H, nVar)
%VALIDATEDLPMIINPUTS Argument checks and CSI-RS RE list shared by the dlPMISelect / riSelect / cqiSelect drop-ins.
%   Own implementation of what the reference keeps as a LOCAL function of dlPMISelect (validateInputs,
%   +communication/+phyLayer/dlPMISelect.m:511-851) and therefore cannot be called from a shim.  Same checks, same
%   'nr5g:dlPMISelect:*' / 'nr5g:hDLPMISelect:*' error identifiers, same defaults; additionally the two report fields only
%   riSelect (RIRestriction, riSelect.m:440-456) and cqiSelect (CQIMode, cqiSelect.m:865-870) validate are passed through
%   so one validated struct serves all three gateways.
%     rc            validated report configuration (NStartBWP, NSizeBWP, CodebookType, CodebookMode, PanelDimensions,
%                   OverSamplingFactors, PMIMode, CQIMode, PRGSize, SubbandSize, CodebookSubsetRestriction, i2Restriction,
%                   RIRestriction)
%     csirsIndSubs  [nRE x 3] subscripts (k, l, port) of the NZP CSI-RS REs of port 1, lowest RE of every CDM group
%     nVar          noise variance clipped at 1e-10 (dlPMISelect.m:846-850)
    fcn = 'hDLPMISelect';
    validateattributes(carrier, {'nrCarrierConfig'}, {'scalar'}, fcn, 'CARRIER');
    validateattributes(csirs, {'nrCSIRSConfig'}, {'scalar'}, fcn, 'CSIRS');
    nPorts = csirs.NumCSIRSPorts;
    if any(nPorts ~= nPorts(1))
        error('nr5g:dlPMISelect:InvalidCSIRSPorts', ...
            'All the CSI-RS resources must be configured to have the same number of CSI-RS ports.');
    end
    nPorts = nPorts(1);
    cdm = cellstr(csirs.CDMType);
    if ~all(strcmpi(cdm, cdm{1}))
        error('nr5g:dlPMISelect:InvalidCSIRSCDMTypes', ...
            'All the CSI-RS resources must be configured to have the same CDM lengths.');
    end

    % ---- bandwidth part ------------------------------------------------------------------------------------------
    rc = struct();
    rc.NStartBWP = bwpField(reportConfig, 'NStartBWP', carrier.NStartGrid, {'scalar','integer','nonnegative','<=',2473}, ...
                            'the start of BWP', fcn);
    rc.NSizeBWP = bwpField(reportConfig, 'NSizeBWP', carrier.NSizeGrid, {'scalar','integer','positive','<=',275}, ...
                           'the size of BWP', fcn);
    if rc.NStartBWP < carrier.NStartGrid
        error('nr5g:dlPMISelect:InvalidNStartBWP', ...
            'The starting resource block of BWP (%d) must be greater than or equal to the starting resource block of carrier (%d).', ...
            rc.NStartBWP, carrier.NStartGrid);
    end
    if rc.NSizeBWP + rc.NStartBWP > carrier.NStartGrid + carrier.NSizeGrid
        error('nr5g:dlPMISelect:InvalidBWPLimits', ...
            ['The sum of starting resource block of BWP (%d) and the size of BWP (%d) must be less than or equal to the sum of ' ...
             'starting resource block of carrier (%d) and size of the carrier (%d).'], ...
            rc.NStartBWP, rc.NSizeBWP, carrier.NStartGrid, carrier.NSizeGrid);
    end

    % ---- codebook type / mode / panel ------------------------------------------------------------------------------
    rc.CodebookType = 'Type1SinglePanel';
    if isfield(reportConfig, 'CodebookType')
        rc.CodebookType = validatestring(reportConfig.CodebookType, {'Type1SinglePanel','Type1MultiPanel'}, fcn, 'CodebookType field');
    end
    single = strcmp(rc.CodebookType, 'Type1SinglePanel');
    rc.CodebookMode = 1;
    if isfield(reportConfig, 'CodebookMode')
        validateattributes(reportConfig.CodebookMode, {'numeric'}, {'scalar','integer','positive','<=',2}, fcn, 'CodebookMode field');
        rc.CodebookMode = reportConfig.CodebookMode;
    end
    N1 = 1; N2 = 1; O = [1 1];
    if single
        if nPorts > 2
            pd = panelField(reportConfig, 2, 'PanelDimensions field for type 1 single panel codebooks', fcn);
            N1 = pd(1); N2 = pd(2);
            if 2*N1*N2 ~= nPorts
                error('nr5g:dlPMISelect:InvalidPanelDimensions', ...
                    ['For the configured number of CSI-RS ports (%d), the given panel configuration [%d %d] is not valid. Note that, ' ...
                     'two times the product of panel dimensions (%d) must be equal to the number of CSI-RS ports (%d).'], ...
                    nPorts, N1, N2, 2*N1*N2, nPorts);
            end
            % TS 38.214 Table 5.2.2.2.1-2: (N1,N2) -> (O1,O2)
            tab = [2 1 4 1; 2 2 4 4; 4 1 4 1; 3 2 4 4; 6 1 4 1; 4 2 4 4; 8 1 4 1; 4 3 4 4; 6 2 4 4; 12 1 4 1; 4 4 4 4; 8 2 4 4; 16 1 4 1];
            row = find(tab(:,1) == N1 & tab(:,2) == N2, 1);
            if isempty(row)
                error('nr5g:dlPMISelect:InvalidPanelConfiguration', ...
                    ['The given panel configuration [%d %d] is not valid for the given CSI-RS configuration. For a number of CSI-RS ' ...
                     'ports, the panel configuration should be one of the possibilities from TS 38.214 Table 5.2.2.2.1-2.'], N1, N2);
            end
            O = tab(row, 3:4);
        end
        rc.PanelDimensions = [N1 N2];
    else
        if ~any(nPorts == [8 16 32])
            error('nr5g:dlPMISelect:InvalidNumCSIRSPortsForMultiPanel', ...
                'For multipanel codebook type, the number of CSI-RS ports must be 8, 16, or 32.');
        end
        pd = panelField(reportConfig, 3, 'PanelDimensions field for type 1 multipanel codebooks', fcn);
        Ng = pd(1); N1 = pd(2); N2 = pd(3);
        if 2*Ng*N1*N2 ~= nPorts
            error('nr5g:dlPMISelect:InvalidMultiPanelDimensions', ...
                ['For the configured number of CSI-RS ports (%d), the given panel configuration [%d %d %d] is not valid. Note that, ' ...
                 'two times the product of panel dimensions (%d) must be equal to the number of CSI-RS ports (%d).'], ...
                nPorts, Ng, N1, N2, 2*Ng*N1*N2, nPorts);
        end
        % TS 38.214 Table 5.2.2.2.2-1: (Ng,N1,N2) -> (O1,O2)
        tab = [2 2 1 4 1; 2 2 2 4 4; 2 4 1 4 1; 4 2 1 4 1; 2 8 1 4 1; 2 4 2 4 4; 4 4 1 4 1; 4 2 2 4 4];
        row = find(tab(:,1) == Ng & tab(:,2) == N1 & tab(:,3) == N2, 1);
        if isempty(row)
            error('nr5g:dlPMISelect:InvalidMultiPanelConfiguration', ...
                ['The given panel configuration [%d %d %d] is not valid for the given CSI-RS configuration. For a number of CSI-RS ' ...
                 'ports, the panel configuration should be one of the possibilities from TS 38.214 Table 5.2.2.2.2-1.'], Ng, N1, N2);
        end
        if rc.CodebookMode == 2 && Ng ~= 2
            error('nr5g:dlPMISelect:InvalidNumPanelsforGivenCodebookMode', ...
                'For codebook mode 2, number of panels Ng (%d) must be 2. Choose appropriate PanelDimensions.', Ng);
        end
        O = tab(row, 4:5);
        rc.PanelDimensions = [Ng N1 N2];
    end
    rc.OverSamplingFactors = O;

    % ---- reporting modes, PRG / subband size -----------------------------------------------------------------------
    rc.PMIMode = modeField(reportConfig, 'PMIMode', fcn);
    rc.CQIMode = modeField(reportConfig, 'CQIMode', fcn);           % cqiSelect.m:865-870
    rc.PRGSize = [];
    if isfield(reportConfig, 'PRGSize') && single
        prg = reportConfig.PRGSize;
        if ~(isnumeric(prg) && isempty(prg))
            validateattributes(prg, {'double','single'}, {'real','scalar'}, fcn, 'PRGSize field');
        end
        if ~(isempty(prg) || any(prg == [2 4]))
            error('nr5g:hDLPMISelect:InvalidPRGSize', 'PRGSize dlPMISelect (%s) must be [], 2, or 4.', num2str(prg));
        end
        rc.PRGSize = prg;
    end
    rc.SubbandSize = [];
    subband = (strcmpi(rc.PMIMode, 'Subband') || strcmpi(rc.CQIMode, 'Subband')) && isempty(rc.PRGSize);
    if subband && rc.NSizeBWP >= 24
        if ~isfield(reportConfig, 'SubbandSize')
            error('nr5g:dlPMISelect:SubbandSizeMissing', ...
                'For the subband mode, SubbandSize field is mandatory when the size of BWP is more than 24 PRBs.');
        end
        validateattributes(reportConfig.SubbandSize, {'double','single'}, {'real','scalar'}, fcn, 'SubbandSize field');
        rc.SubbandSize = reportConfig.SubbandSize;
        % TS 38.214 Table 5.2.1.4-2: BWP size range -> the two configurable subband sizes
        ranges = [24 72 4 8; 73 144 8 16; 145 275 16 32];
        allowed = ranges(rc.NSizeBWP >= ranges(:,1) & rc.NSizeBWP <= ranges(:,2), 3:4);
        if ~any(rc.SubbandSize == allowed)
            error('nr5g:hDLPMISelect:InvalidSubbandSize', ...
                'For the configured BWP size (%d), subband size (%d) must be %d or %d.', rc.NSizeBWP, rc.SubbandSize, allowed(1), allowed(2));
        end
    end

    % ---- restrictions ----------------------------------------------------------------------------------------------
    if nPorts > 2, nCsr = N1*O(1)*N2*O(2); elseif nPorts == 2, nCsr = 6; else, nCsr = 1; end
    rc.CodebookSubsetRestriction = bitField(reportConfig, 'CodebookSubsetRestriction', nCsr, nPorts >= 2, fcn, 'CodebookSubsetRestriction field');
    rc.i2Restriction = bitField(reportConfig, 'i2Restriction', 16, nPorts > 2 && single, fcn, 'i2Restriction field');
    if single                                                       % riSelect.m:440-456
        rc.RIRestriction = bitField(reportConfig, 'RIRestriction', 8, true, fcn, 'RIRestriction field in type 1 single panel codebook type');
    else
        rc.RIRestriction = bitField(reportConfig, 'RIRestriction', 4, true, fcn, 'RIRestriction field in type 1 multi panel codebook type');
    end

    % ---- nLayers, H ------------------------------------------------------------------------------------------------
    maxNu = 8; if ~single, maxNu = 4; end
    validateattributes(nLayers, {'numeric'}, {'scalar','integer','positive','<=',maxNu}, fcn, ...
        sprintf('NLAYERS(%d) when codebook type is "%s"', nLayers, rc.CodebookType));
    validateattributes(H, {'double','single'}, {}, fcn, 'H');
    validateattributes(ndims(H), {'double'}, {'>=',2,'<=',4}, fcn, 'number of dimensions of H');

    % ---- NZP CSI-RS REs: port 1, lowest RE of every CDM group (dlPMISelect.m:796-828) ------------------------------
    types = cellstr(csirs.CSIRSType);
    nZP = sum(strcmpi(types, 'zp'));
    perRes = nrCSIRSIndices(carrier, csirs, 'IndexStyle', 'subscript', 'OutputResourceFormat', 'cell');
    perRes = perRes(nZP+1:end);
    switch lower(cdm{1})                         % OFDM symbols a CDM group spans; 0 = no CDM, keep every RE
        case 'nocdm',   div = 0;
        case 'fd-cdm2', div = 1;
        case 'cdm4',    div = 2;
        otherwise,      div = 4;                 % 'CDM8'
    end
    csirsIndSubs = zeros(0, 3);
    for r = 1:numel(perRes)
        sub = double(perRes{r});
        sub = sub(sub(:,3) == 1, :);
        if div > 0
            sub = sub(1:size(sub,1)/div, :);     % the REs of the first symbol of each group
            sub = sub(1:2:end, :);               % lowest subcarrier of each frequency-domain pair
        end
        csirsIndSubs = [csirsIndSubs; sub]; %#ok<AGROW>
    end
    if ~isempty(csirsIndSubs)
        validateattributes(H, {class(H)}, {'size', [carrier.NSizeGrid*12 carrier.SymbolsPerSlot NaN nPorts]}, fcn, 'H');
        maxRank = min(size(H, 3), nPorts);
        if nLayers > maxRank
            error('nr5g:hDLPMISelect:InvalidNumLayers', ...
                'The given antenna configuration (%dx%d) supports only up to (%d) layers.', nPorts, size(H, 3), maxRank);
        end
    end

    % ---- nVar ------------------------------------------------------------------------------------------------------
    validateattributes(nVar, {'double','single'}, {'scalar','real','nonnegative','finite'}, fcn, 'NVAR');
    nVar = max(double(nVar), 1e-10);
end

function v = bwpField(cfg, name, default, attrs, what, fcn)
    if ~isfield(cfg, name)
        error(['nr5g:dlPMISelect:' name 'Missing'], '%s field is mandatory.', name);
    end
    v = cfg.(name);
    if isnumeric(v) && isempty(v)
        v = default;                              % [] = the whole carrier
    else
        validateattributes(v, {'double','single'}, attrs, fcn, what);
    end
    v = double(v);
end

function pd = panelField(cfg, n, what, fcn)
    if ~isfield(cfg, 'PanelDimensions')
        error('nr5g:dlPMISelect:PanelDimensionsMissing', 'PanelDimensions field is mandatory.');
    end
    validateattributes(cfg.PanelDimensions, {'double','single'}, {'vector','numel',n}, fcn, what);
    pd = double(cfg.PanelDimensions(:).');
end

function m = modeField(cfg, name, fcn)
    m = 'Wideband';
    if isfield(cfg, name)
        m = validatestring(cfg.(name), {'Wideband','Subband'}, fcn, [name ' field']);
    end
end

function bits = bitField(cfg, name, n, applicable, fcn, what)
    bits = ones(1, n);
    if applicable && isfield(cfg, name) && ~isempty(cfg.(name))
        validateattributes(cfg.(name), {'numeric'}, {'vector','binary','numel',n}, fcn, what);
        bits = double(cfg.(name)(:).');
    end
end
