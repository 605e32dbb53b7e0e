function [antsym, antind] = prgPrecode(siz, nstartgrid, portsym, portind, F)
%PRGPRECODE Drop-in for communication.phyLayer.prgPrecode (+communication/+phyLayer/prgPrecode.m:53; gNBPhy.m:822,826).
% One launch over all REs; the per-PRG zero grids of the reference (:128) are never formed.
    if ismatrix(F), F = reshape(F, size(F, 1), size(F, 2), 1); end      % wideband precoder (:86-92)
    proto = portsym;
    [antsym, antind] = isac_prg_precode_mex(double(siz(:).'), double(nstartgrid), single(complex(portsym)), int32(portind), ...
                                            single(complex(F)));
    antsym = cast(antsym, 'like', proto);
    antind = uint32(antind);
end
