function [pmi, sinr, subbandIndices] = pmiSelect(nlayers, hest, noiseest, bandSize)
%PMISELECT Drop-in for communication.phyLayer.pmiSelect (+communication/+phyLayer/pmiSelect.m:28; call site gNBPhy.m:1035).
    if noiseest == 0                                   % pmiSelect.m:40, :59-64
        pmi = NaN; sinr = NaN; subbandIndices = NaN; return
    end
    [pmi, sinr, subbandIndices] = isac_ul_pmi_mex(nlayers, single(hest), double(noiseest), double(bandSize));
    pmi = pmi(:);                                      % column like [~,pmi] = max(sinrBands,[],2) (:56)
end
