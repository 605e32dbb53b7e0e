function H = cdlChannelMatrix(channel, carrier, slotStartTime)
%CDLCHANNELMATRIX Frequency-domain channel matrix H[K x L x nRx x nTx] of one slot from an nrCDLChannel configuration.
% Device replacement of the pair "filter the waveform through nrCDLChannel (uePhy.m:731, gNBPhy.m:840), then recover H with
% nrChannelEstimate (uePhy.m:897, gNBPhy.m:1030)": call it in the CSI branch of uePhy.phyRxProcessing (uePhy.m:886-932) with
% the UE's channel object (built at +parameters/+channelModels/+communication/cdl.m:48-88, profile set by
% communication.channelModels.updateCDLModels.m:7-15) and hand H to riSelect / cqiSelect.  Only the properties cdl.m sets are
% read (DelayProfile, DelaySpread, CarrierFrequency, antenna array sizes) plus MaximumDopplerShift and Seed; the realisation
% is statistically equivalent to the toolbox object, not sample-identical (different random stream).
    prof = find(strcmpi(channel.DelayProfile, {'CDL-A','CDL-B','CDL-C','CDL-D','CDL-E'})) - 1;
    if isempty(prof), error('isac:cdlChannelMatrix:profile', 'unsupported DelayProfile %s', channel.DelayProfile); end
    ts = channel.TransmitAntennaArray.Size; rs = channel.ReceiveAntennaArray.Size;
    cfg = struct('profile', prof, 'delaySpread', channel.DelaySpread, 'fc', channel.CarrierFrequency, ...
                 'maxDoppler', channel.MaximumDopplerShift, 'txSize', int32(ts(1:3)), 'rxSize', int32(rs(1:3)), ...
                 'txPattern38901', 1, 'rxPattern38901', 0, 'seed', double(channel.Seed));
    info = nrOFDMInfo(carrier);
    starts = cumsum([0 info.SymbolLengths(1:carrier.SymbolsPerSlot-1)]) / info.SampleRate;   % symbol start times in the slot
    H = isac_cdl_mex(cfg, 12*carrier.NSizeGrid, carrier.SubcarrierSpacing*1e3, starts, slotStartTime);
end
