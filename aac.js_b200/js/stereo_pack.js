/*
 * stereo_pack.js -- the host's share of the stereo tools of a channel pair element.
 *
 * aac.js runs processMS (src/decoder.js:379-404) and processIS (src/decoder.js:337-376) over the
 * 2 x 1024 dequantised coefficients of a CPEElement on the CPU.  With the B200 path those
 * coefficients are staged to the device anyway, so the host only WALKS the bands -- the same two
 * loops, same order, same conditions -- and records, per group of 4 coefficients (every
 * scalefactor-band edge is a multiple of 4, src/tables.js:34-124), what the device has to do:
 *      op = 0       untouched
 *      op = 1       l' = l + r, r' = l - r                  (decoder.js:395-397)
 *      op = 2 + k   r' = l * scale[k], scale = c * scaleFactors[idx]   (decoder.js:360-366)
 * into one 768-byte aacfb_stereo_ops record (include/aacfb.h): 256 op bytes + 128 float32 scales.
 * Python twin: aacjs_b200.pack_stereo (tested against the reference's own functions).
 */
var ICStream = require('aac/src/ics');

exports.RECORD_BYTES = 768;

// element: CPEElement; ops: Uint8Array view of the record's 256 op bytes; scales: Float32Array
// view of its 128 scales.  Returns true when the record holds at least one op.
exports.pack = function(element, ops, scales) {
    var present = false, g, i, w, k, idx, groupOff, ics, info, offsets;
    for (k = 0; k < 256; k++) ops[k] = 0;

    if (element.commonWindow && element.maskPresent) {          // decoder.js:295-296
        ics = element.left; info = ics.info; offsets = info.swbOffsets;
        var cbl = ics.bandTypes, cbr = element.right.bandTypes;
        groupOff = 0; idx = 0;
        for (g = 0; g < info.groupCount; g++) {
            for (i = 0; i < info.maxSFB; i++, idx++) {
                if (element.ms_used[idx] && cbl[idx] < ICStream.NOISE_BT && cbr[idx] < ICStream.NOISE_BT) {
                    for (w = 0; w < info.groupLength[g]; w++) {
                        var a = (groupOff + w * 128 + offsets[i]) >> 2, b = (groupOff + w * 128 + offsets[i + 1]) >> 2;
                        for (k = a; k < b; k++) ops[k] = 1;
                        present = true;
                    }
                }
            }
            groupOff += info.groupLength[g] * 128;
        }
    }

    ics = element.right; info = ics.info; offsets = info.swbOffsets;
    var bandTypes = ics.bandTypes, sectEnd = ics.sectEnd, scaleFactors = ics.scaleFactors, nScales = 0;
    idx = 0; groupOff = 0;
    for (g = 0; g < info.groupCount; g++) {
        for (i = 0; i < info.maxSFB;) {
            var end = sectEnd[idx];
            if (bandTypes[idx] === ICStream.INTENSITY_BT || bandTypes[idx] === ICStream.INTENSITY_BT2) {
                for (; i < end; i++, idx++) {
                    var c = bandTypes[idx] === ICStream.INTENSITY_BT ? 1 : -1;
                    if (element.maskPresent) c *= element.ms_used[idx] ? -1 : 1;
                    scales[nScales] = c * scaleFactors[idx];
                    for (w = 0; w < info.groupLength[g]; w++) {
                        var a2 = (groupOff + w * 128 + offsets[i]) >> 2, b2 = (groupOff + w * 128 + offsets[i + 1]) >> 2;
                        for (k = a2; k < b2; k++) ops[k] = 2 + nScales;
                        present = true;
                    }
                    nScales++;
                }
            } else {
                idx += end - i;
                i = end;
            }
        }
        groupOff += info.groupLength[g] * 128;
    }
    return present;
};
