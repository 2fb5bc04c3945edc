/*
 * filter_bank.js -- drop-in for aac.js src/filter_bank.js backed by the B200 library.
 *
 *   new FilterBank(smallFrames, channels)                 (reference :24-44)
 *   filterBank.process(info, input, output, channel)      (reference :88-204)
 *
 * Same contract: throws on smallFrames, reads info.windowSequence and
 * info.windowShape[0..1], does not touch `input`, overwrites `output`, keeps the
 * per-channel overlap between calls (on the GPU).  The per-frame path is the
 * correctness seam; throughput comes from the batched decoder (decoder_b200.js).
 */
var addon = require('./build/Release/aacfb.node');

function FilterBank(smallFrames, channels, sampleIndex) {
    // the library throws "WHA?? No small frames allowed." exactly like the reference
    this.handle = addon.create(0, 1, channels, sampleIndex === undefined ? 4 : sampleIndex, smallFrames ? 1 : 0, 0);
    this.length = 1024;
    this.shortLength = 128;
    this.infoBytes = new Uint8Array(8);
}

FilterBank.prototype.process = function(info, input, output, channel) {
    var b = this.infoBytes;
    b[0] = info.windowSequence;
    b[1] = info.windowShape[0];
    b[2] = info.windowShape[1];
    b[3] = info.maxSFB | 0;
    addon.filterbankProcess(this.handle, 0, channel, b, input, output);
};

module.exports = FilterBank;
