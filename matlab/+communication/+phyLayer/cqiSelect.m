function [CQI, PMISet, CQIInfo, PMIInfo] = cqiSelect(carrier, csirs, reportConfig, nLayers, H, varargin)
%CQISELECT Drop-in for communication.phyLayer.cqiSelect (+communication/+phyLayer/cqiSelect.m:1; call site uePhy.m:907).
% CQIInfo / PMIInfo carry the fields the callers of the reference read (uePhy.m:907-932 reads none of them); ask
% communication.phyLayer.dlPMISelect for the full SINR arrays.
    nVar = 1e-10; SINRTable = [];
    if nargin >= 6, nVar = varargin{1}; end
    if nargin >= 7, SINRTable = varargin{2}; end
    if isempty(SINRTable), t = communication.setupSINRtoCQIMappingTable(); SINRTable = t.downlinkSINR90pc; end
    cfg = communication.phyLayer.isacCsiConfig(carrier, csirs, reportConfig, nLayers, H, nVar);
    [~, i1, i2, CQI] = isac_csi_report_mex(cfg, single(H), double(nVar), double(SINRTable(:)), 0, 2, nLayers);
    PMISet.i1 = i1(:).'; PMISet.i2 = i2(:).';
    if nLayers <= 4, CQI = CQI(:, 1); end          % one codeword (cqiSelect.m:576-632)
    CQIInfo = struct('SINRPerSubbandPerCW', [], 'TransportBLER', []);
    PMIInfo = struct('SINRPerRE', [], 'SINRPerSubband', [], 'W', []);
end
