/*
 * quant_pack.js -- the host's share of ICStream.decodeSpectralData when the B200 dequantises.
 *
 * aac.js interleaves the Huffman decode with the inverse quantisation and the scalefactor
 * multiplication (src/ics.js:243-256) and fills noise bands with its generator (src/ics.js:228-242),
 * all on the CPU, and the result -- 1024 Float32 per channel -- is what would have to cross PCIe.
 * With quantOnDevice the bit parse keeps only what is serial: the Huffman decode.  The integers it
 * yields (buf[j], ics.js:247) go into an Int16Array in data[] order, the band types and scalefactors
 * decodeBandTypes / decodeScaleFactors left behind become one 16-bit code per band, and the device
 * runs ics.js:203-266's arithmetic on the staged row: one aacfb_qframe record of 2304 bytes per
 * channel-frame (include/aacfb.h) instead of 4096.
 *
 *   decodeSpectralData   replacement for ICStream.prototype.decodeSpectralData, installed by
 *                        decoder_b200.js around the reference's own readChunk (and removed again):
 *                        the same walk over groups / bands / windows, the reference's own
 *                        Huffman.decodeSpectralData calls in the same order, no arithmetic.
 *   pack                 ICStream -> aacfb_qframe bytes.
 * Python twin of pack: aacjs_b200.pack_qframe.
 */
var ICStream = require('aac/src/ics');
var Huffman = require('aac/src/huffman');
var tables = require('aac/src/tables');

exports.RECORD_BYTES = 2304;
var BAND_ZERO = 0x0000, BAND_SPECTRAL = 0x4000, BAND_NOISE = 0x8000, BAND_UNDEFINED = 0x01ff;

// scaleFactors[idx] is +-SCALEFACTOR_TABLE[i] (ics.js:144,158,171): distinct powers 2^((i-200)/4), so the
// index is recovered exactly from the value; anything else (NaN: a read outside the table) -> UNDEFINED.
var sfIndex = null;
function scalefactorIndex(value) {
    if (sfIndex === null) {
        sfIndex = {};
        for (var i = 0; i < tables.SCALEFACTOR_TABLE.length; i++) sfIndex[tables.SCALEFACTOR_TABLE[i]] = i;
    }
    var i = sfIndex[Math.abs(value)];
    return i === undefined ? BAND_UNDEFINED : i;
}
exports.scalefactorIndex = scalefactorIndex;

// ics.js:203-266 without the arithmetic: this.quant[off + k + j] = buf[j].
exports.decodeSpectralData = function(stream) {
    var info = this.info, maxSFB = info.maxSFB, windowGroups = info.groupCount, offsets = info.swbOffsets,
        bandTypes = this.bandTypes, buf = this.specBuf;
    var quant = this.quant || (this.quant = new Int16Array(1024));
    var groupOff = 0, idx = 0;
    for (var g = 0; g < windowGroups; g++) {
        var groupLen = info.groupLength[g];
        for (var sfb = 0; sfb < maxSFB; sfb++, idx++) {
            var hcb = bandTypes[idx], off = groupOff + offsets[sfb], width = offsets[sfb + 1] - offsets[sfb];
            if (hcb === ICStream.ZERO_BT || hcb === ICStream.INTENSITY_BT || hcb === ICStream.INTENSITY_BT2 ||
                hcb === ICStream.NOISE_BT) continue;               // nothing in the bitstream for these bands
            for (var group = 0; group < groupLen; group++, off += 128) {
                var num = (hcb >= ICStream.FIRST_PAIR_BT) ? 2 : 4;
                for (var k = 0; k < width; k += num) {
                    Huffman.decodeSpectralData(stream, hcb, buf, 0);
                    for (var j = 0; j < num; j++) quant[off + k + j] = buf[j];
                }
            }
        }
        groupOff += groupLen << 7;
    }
    if (this.pulsePresent) throw new Error('TODO: add pulse data');   // ics.js:263-265
};

// One record at byte offset `at` of (Uint8Array bytes, DataView view) over the same buffer.
exports.pack = function(ics, bytes, view, at) {
    var info = ics.info, groups = info.groupCount, maxSFB = info.maxSFB, i, idx;
    for (i = 0; i < 8; i++) bytes[at + i] = i < groups ? info.groupLength[i] : 0;
    for (idx = 0; idx < 120; idx++) {
        var code = BAND_ZERO;
        if (idx < groups * maxSFB) {
            var t = ics.bandTypes[idx];
            if (t === ICStream.ZERO_BT || t === ICStream.INTENSITY_BT || t === ICStream.INTENSITY_BT2) code = BAND_ZERO;
            else if (t === ICStream.NOISE_BT) code = BAND_NOISE | scalefactorIndex(ics.scaleFactors[idx]);
            else code = BAND_SPECTRAL | scalefactorIndex(ics.scaleFactors[idx]);
        }
        view.setUint16(at + 8 + 2 * idx, code, true);
    }
    for (i = 0; i < 8; i++) bytes[at + 248 + i] = 0;
    var quant = ics.quant;
    for (i = 0; i < 1024; i++) view.setInt16(at + 256 + 2 * i, quant ? quant[i] : 0, true);
};
