function rxWaveform = basicRadarChannel(txWaveform, radarParams, targetLoSConditions, varargin)
%BASICRADARCHANNEL Drop-in for sensing.channelModels.basicRadarChannel (+sensing/+channelModels/basicRadarChannel.m:1).
% isac_radar_channel_mex marshals onto isac_radar_channel_dev.  Noise: MATLAB's randn stream is not reproducible outside
% MATLAB, so the device draws Philox noise of the same variance (N0/2 per component, basicRadarChannel.m:67-69); pass
% randn(size(txWaveform)) + 1j*randn(size(txWaveform)) as a 4th argument for a bit-repeatable run.
    cfg = struct('fc', radarParams.fc, 'fs', radarParams.fs, 'N0', radarParams.N0, 'range', radarParams.range(:), ...
                 'velocity', radarParams.velocity(:), 'largeScaleFading', radarParams.largeScaleFading(:), ...
                 'steeringVec', radarParams.RxSteeringVec, 'los', int32(targetLoSConditions(:)));
    if nargin > 3
        rxWaveform = double(isac_radar_channel_mex(cfg, single(txWaveform), uint64(0), single(varargin{1})));
    else
        rxWaveform = double(isac_radar_channel_mex(cfg, single(txWaveform), uint64(randi(2^31))));
    end
end
