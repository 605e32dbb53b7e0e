function PMISet = isacPMISet(cfg, i1, i2, mp)
%ISACPMISET PMISet structure from the gateway outputs.  Type1SinglePanel: i1 = [i11 i12 i13], i2 per subband.
% Type1MultiPanel (cfg.nPanels >= 2): the gateway returns i1(3) and i2 as linear indices into the flattened index sets
% [i13 i141 i142 i143] and [i20 i21 i22] of the reported rank (lengths in mp); un-flatten them into the reference's
% i1 = [i11 i12 i13 i141 i142 i143], i2 = [i20; i21; i22] per subband (dlPMISelect.m:455-457, :489).
    PMISet.i1 = i1(:).'; PMISet.i2 = i2(:).';
    if cfg.nPanels < 2, return; end
    PMISet.i1 = NaN(1, 6); PMISet.i2 = NaN(3, numel(i2));
    if any(isnan(i1)) || ~any(mp), return; end
    [a, b, c, d] = ind2sub(mp(4:7), i1(3));
    PMISet.i1 = [i1(1) i1(2) a b c d];
    for sb = find(~isnan(i2(:).'))
        [a, b, c] = ind2sub(mp(1:3), i2(sb));
        PMISet.i2(:, sb) = [a; b; c];
    end
end
