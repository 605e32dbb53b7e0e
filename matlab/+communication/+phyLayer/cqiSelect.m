function [CQI, PMISet, CQIInfo, PMIInfo] = cqiSelect(carrier, csirs, reportConfig, nLayers, H, varargin)
%CQISELECT Drop-in for communication.phyLayer.cqiSelect (+communication/+phyLayer/cqiSelect.m:1; call site uePhy.m:907).
% CQI and PMISet come from one fused device evaluation (isac_cqi_select_dev).  The information outputs are filled like the
% reference's (cqiSelect.m:636-695): CQIInfo.SINRPerSubbandPerCW / SubbandCQI from the same evaluation; PMIInfo and
% CQIInfo.SINRPerRBPerCW need SINRPerRE, which only communication.phyLayer.dlPMISelect keeps -- it is called (as the
% reference does at cqiSelect.m:507) only when a caller asks for a third or fourth output (uePhy.m:907 asks for two).
    nVar = 1e-10; SINRTable = [];
    if nargin >= 6, nVar = varargin{1}; end
    if nargin >= 7, SINRTable = varargin{2}; end
    if isempty(SINRTable), t = communication.setupSINRtoCQIMappingTable(); SINRTable = t.downlinkSINR90pc; end
    [cfg, rc] = communication.phyLayer.isacCsiConfig(carrier, csirs, reportConfig, nLayers, H, nVar);
    [~, i1, i2, CQI, sbcw, mp] = isac_csi_report_mex(cfg, single(H), double(nVar), double(SINRTable(:)), 0, 2, nLayers);
    PMISet = communication.phyLayer.isacPMISet(cfg, i1, i2, mp);
    nCW = ceil(nLayers/4);
    CQI = CQI(:, 1:nCW);                          % one codeword up to four layers (cqiSelect.m:576-632)
    if nargout < 3, return; end
    sbcw = sbcw(:, 1:nCW);
    subbandCQI = NaN(size(sbcw));                 % getCQI (cqiSelect.m:697-722): last table entry <= SINR in dB, 0 if none
    for q = find(~isnan(sbcw(:))).'
        subbandCQI(q) = sum(SINRTable(:) <= 10*log10(sbcw(q)));
    end
    if strcmpi(rc.CQIMode, 'Wideband')            % cqiSelect.m:686-690
        sbcw = sbcw(1, :); subbandCQI = subbandCQI(1, :);
    end
    CQIInfo.SINRPerSubbandPerCW = sbcw;
    CQIInfo.SubbandCQI = subbandCQI;
    [~, PMIInfo] = communication.phyLayer.dlPMISelect(carrier, csirs, reportConfig, nLayers, H, nVar);
    CQIInfo.SINRPerRBPerCW = sinrPerRB(PMIInfo.SINRPerRE, PMISet, cfg, rc, carrier, nCW);
end

function out = sinrPerRB(SINRPerRE, PMISet, cfg, rc, carrier, nCW)
% SINR per RB and codeword at the reported PMI (getSINRperRB, cqiSelect.m:724-766): pick each subband's (i2, i1) slice,
% add the layers of a codeword (nrLayerDemap: floor(nu/2) layers in codeword 1 when nu > 4), average the REs of an RB.
    nRB = rc.NSizeBWP; L = carrier.SymbolsPerSlot; nu = size(SINRPerRE, 3);
    out = NaN(nRB, L, nCW);
    if any(isnan(PMISet.i1)) || cfg.nPanels >= 2, return; end   % multi-panel: SINRPerRBPerCW is left NaN (11-D slices of :734-736)
    if cfg.pmiSubband && cfg.subbandSize > 0 && nRB >= 24   % PMI subband sizes (getDownlinkPMISubbandInfo, dlPMISelect.m:1836-1887)
        first = cfg.subbandSize - mod(rc.NStartBWP, cfg.subbandSize);
        sizes = first;
        while sum(sizes) < nRB, sizes(end+1) = min(cfg.subbandSize, nRB - sum(sizes)); end %#ok<AGROW>
    else
        sizes = nRB;
    end
    sel = NaN(nRB*12, L, nu);
    rb0 = 0;
    for sb = 1:numel(sizes)
        rows = rb0*12 + 1 : (rb0 + sizes(sb))*12;
        if ~isnan(PMISet.i2(sb))
            sel(rows, :, :) = SINRPerRE(rows, :, :, PMISet.i2(sb), PMISet.i1(1), PMISet.i1(2), PMISet.i1(3));
        end
        rb0 = rb0 + sizes(sb);
    end
    if nCW == 1, groups = {1:nu}; else, groups = {1:floor(nu/2), floor(nu/2)+1:nu}; end
    perCW = NaN(nRB*12, L, nCW);
    for c = 1:nCW
        s = sum(sel(:, :, groups{c}), 3);                  % NaN where the RE carries no CSI-RS
        perCW(:, :, c) = s;
    end
    for rb = 1:nRB
        blk = perCW((rb-1)*12 + (1:12), :, :);
        out(rb, :, :) = mean(blk, 1, 'omitnan');           % all-NaN RB -> NaN, as the reference's mean(...,'omitnan')
    end
end
