function estResults = fft2D(radarEstParams, cfar, rxGrid, txGrid)
%FFT2D Drop-in for sensing.estimation.fft2D of the reference (+sensing/+estimation/fft2D.m:1).
% Put <repo>/matlab ahead of the reference on the MATLAB path; the arithmetic runs in libisac_b200.so.
    cfg = struct('nIFFT', radarEstParams.nIFFT, 'nFFT', radarEstParams.nFFT, ...
                 'rRes', radarEstParams.rRes, 'vRes', radarEstParams.vRes, ...
                 'cutRows', double([min(cfar.CUTIdx(1,:)) max(cfar.CUTIdx(1,:))]), ...
                 'cutCols', double([min(cfar.CUTIdx(2,:)) max(cfar.CUTIdx(2,:))]), ...
                 'Pfa', radarEstParams.Pfa);
    det = cfar.cfarDetector2D;                      % phased.CFARDetector2D built by sensing.detection.cfar2D (cfar2D.m:27-33)
    cfg.guardBand = double(det.GuardBandSize(:).') .* [1 1];
    cfg.trainBand = double(det.TrainingBandSize(:).') .* [1 1];
    if isprop(det, 'ProbabilityFalseAlarm'), cfg.Pfa = det.ProbabilityFalseAlarm; end
    cfg = sensing.estimation.isacDoaConfig(radarEstParams, cfg);
    estResults = isac_fft2d_mex(cfg, single(rxGrid), single(txGrid));   % double -> single at the boundary
end
