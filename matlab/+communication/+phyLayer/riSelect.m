function [RI, PMISet] = riSelect(carrier, csirs, reportConfig, H, varargin)
%RISELECT Drop-in for communication.phyLayer.riSelect (+communication/+phyLayer/riSelect.m:1; call site uePhy.m:900).
% All valid ranks are scored in one fused launch (isac_ri_select_dev) instead of one dlPMISelect call per rank (:254-285).
    if nargin == 5, nVar = varargin{1}; else, nVar = 1e-10; end
    cfg = communication.phyLayer.isacCsiConfig(carrier, csirs, reportConfig, 1, H, nVar);
    [RI, i1, i2, ~, ~, mp] = isac_csi_report_mex(cfg, single(H), double(nVar), 0, 0, 1);
    PMISet = communication.phyLayer.isacPMISet(cfg, i1, i2, mp);
end
