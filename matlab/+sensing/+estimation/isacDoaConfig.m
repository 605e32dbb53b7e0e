function cfg = isacDoaConfig(radarEstParams, cfg)
%ISACDOACONFIG Scan-grid / array fields of the DoA gateways (music.m:11-15, radarParams.m:120-124).
    if nargin < 2, cfg = struct; end
    cfg.aGran = radarEstParams.azimuthScanGranularity; cfg.aMax = radarEstParams.azimuthScanScale;
    cfg.eGran = radarEstParams.elevationScanGranularity; cfg.eMax = radarEstParams.elevationScanScale;
    cfg.isUpa = 0; cfg.nAnts = 0; cfg.nX = 0; cfg.nY = 0;
    ant = radarEstParams.antennaType;
    if isa(ant, 'parameters.baseStation.antenna.upa')
        cfg.isUpa = 1; cfg.nX = ant.nV; cfg.nY = ant.nH;
    else
        cfg.nAnts = ant.numElements;
    end
end
