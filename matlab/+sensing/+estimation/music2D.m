function estResults = music2D(rdrEstParams, bsParams, rxGrid, txGrid)
%MUSIC2D Drop-in for sensing.estimation.music2D (+sensing/+estimation/music2D.m:1).
    cfg = struct('scsHz', bsParams.scs*1e3, 'fc', rdrEstParams.fc, 'Tsri', rdrEstParams.Tsri, ...
                 'rMax', rdrEstParams.cfarEstZone(1,2), 'vZone', rdrEstParams.cfarEstZone(2,2));   % music2D.m:35-43
    cfg = sensing.estimation.isacDoaConfig(rdrEstParams, cfg);
    estResults = isac_music2d_mex(cfg, single(rxGrid), single(txGrid));
end
