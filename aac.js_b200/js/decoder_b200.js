/*
 * decoder_b200.js -- AACDecoder with the filterbank-synthesis path on a B200.
 *
 * Keeps the Aurora.js Decoder plugin surface of aac.js src/decoder.js (init /
 * setCookie / readChunk returning interleaved Float32 PCM).  The serial
 * ADTS / Huffman / ICS parse stays exactly where it is (the reference's own code, unchanged,
 * on the CPU).  M/S and intensity stereo either stay there too (stereoOnDevice = false) or,
 * for plain stereo streams, become a 768-byte op record per frame that the device applies to
 * the staged spectra (stereo_pack.js, aacfb_process_stereo).  What changes is the tail of
 * the frame (src/decoder.js:263-269 / 309-319 and :204-213): instead of running
 * tns.process + filter_bank.process + the interleave per frame in JS, readChunk
 * parses up to K frames ahead, stages their spectra and side info into typed
 * arrays and makes ONE aacfb_process call; it returns K*1024*channels samples.
 * K = 1 is the PR-1 style correctness path.  Coupling elements are outside the accelerated path
 * (stage() throws).  The reference is used unmodified: the parse of each access unit is its own
 * readChunk with `process` intercepted (parseElements below).
 */
var AV = require('av');
var AACDecoder = require('aac/src/decoder');      // the unmodified reference
var ICStream = require('aac/src/ics');
var CPEElement = require('aac/src/cpe');
var addon = require('./build/Release/aacfb.node');
var tnsPack = require('./tns_pack');
var stereoPack = require('./stereo_pack');

var B200Decoder = AACDecoder.extend(function() {
    AV.Decoder.register('mp4a', this);
    AV.Decoder.register('aac ', this);

    this.prototype.framesPerChunk = 64;            // K
    this.prototype.tnsMode = tnsPack.AS_SHIPPED;   // literal parity with the reference by default
    this.prototype.stereoOnDevice = true;          // M/S + IS applied by the kernels (channel pairs of 2-channel streams)

    var setCookie = AACDecoder.prototype.setCookie;
    this.prototype.setCookie = function(buffer) {
        setCookie.call(this, buffer);              // parses the config, throws like the reference
        var C = this.config.chanConfig, K = this.framesPerChunk;
        this.handle = addon.create(0, 1, C, this.config.sampleIndex, 0, this.tnsMode);
        this.spectra = new Float32Array(K * C * 1024);
        this.info = new Uint8Array(K * C * 8);
        this.tnsOffsets = new Uint32Array(K * C + 1);
        this.tnsBuf = new ArrayBuffer(K * C * (8 + 8 * 4 * 84));
        this.tnsBytes = new Uint8Array(this.tnsBuf);
        this.tnsView = new DataView(this.tnsBuf);
        this.deviceStereo = this.stereoOnDevice && C === 2;
        if (this.deviceStereo) {                   // one aacfb_stereo_ops record per frame
            this.stereoBuf = new ArrayBuffer(K * stereoPack.RECORD_BYTES);
            this.stereoBytes = new Uint8Array(this.stereoBuf);
        }
    };

    // the part of process(elements) before TNS: M/S, IS (decoder.js:295-302), then stage
    this.prototype.stage = function(elements, t) {
        var C = this.config.chanConfig, channel = 0, self = this;
        function put(ics, ch) {
            var cf = t * C + ch;
            self.spectra.set(ics.data, cf * 1024);
            var b = self.info, o = cf * 8;
            b[o] = ics.info.windowSequence; b[o + 1] = ics.info.windowShape[0]; b[o + 2] = ics.info.windowShape[1];
            b[o + 3] = ics.info.maxSFB; b[o + 4] = ics.tnsPresent ? 1 : 0; b[o + 5] = b[o + 6] = b[o + 7] = 0;
            self.tnsOffsets[cf] = self.tnsLen;
            if (ics.tnsPresent) self.tnsLen = (tnsPack.writeBlock(ics.tns, self.tnsBytes, self.tnsView, self.tnsLen) + 3) & ~3;
        }
        for (var i = 0; i < elements.length && channel < C; i++) {
            var e = elements[i];
            if (e instanceof ICStream) { put(e, channel); channel += 1; }
            else if (e instanceof CPEElement) {
                var onDevice = false;
                if (this.deviceStereo) {
                    var at = t * stereoPack.RECORD_BYTES;
                    onDevice = stereoPack.pack(e, new Uint8Array(this.stereoBuf, at, 256),
                                               new Float32Array(this.stereoBuf, at + 256, 128));
                    if (onDevice) this.anyStereo = true;
                } else {
                    if (e.commonWindow && e.maskPresent) this.processMS(e, e.left.data, e.right.data);
                    this.processIS(e, e.left.data, e.right.data);
                }
                put(e.left, channel); put(e.right, channel + 1);
                if (onDevice) this.info[(t * C + channel) * 8 + 5] = 1;   // stereo_present, left channel
                channel += 2;
            } else throw new Error('coupling elements are outside the accelerated path');
        }
    };

    // The bit parse of ONE access unit, done by the reference itself: its readChunk (decoder.js:125-216)
    // reads the ADTS header, parses the elements, aligns, and then calls this.process(elements) --
    // which is intercepted here to capture the parsed elements instead of running TNS / filterbank on
    // the CPU; the interleave that follows finds no channel data and returns an empty array.  Nothing
    // in the reference has to be modified or refactored for this.
    var referenceReadChunk = AACDecoder.prototype.readChunk;
    this.prototype.parseElements = function() {
        var captured = null, self = this, own = this.process;
        this.process = function(elements) { captured = elements; self.data = []; };
        try { referenceReadChunk.call(this); } finally { this.process = own; }
        return captured;
    };

    this.prototype.readChunk = function() {
        var C = this.config.chanConfig, K = this.framesPerChunk, t = 0, stream = this.bitstream;
        this.tnsLen = 0;
        this.anyStereo = false;
        while (t < K) {
            var mark = stream.offset();
            try {
                var elements = this.parseElements();   // the reference's own parse (decoder.js:129-200)
                this.stage(elements, t++);
            } catch (err) {
                if (!(err instanceof AV.UnderflowError) || t === 0) throw err;
                stream.seek(mark);                      // keep the partial frame for the next call
                break;
            }
        }
        this.tnsOffsets[t * C] = this.tnsLen;
        var pcm = new Float32Array(t * 1024 * C);
        if (this.anyStereo)
            addon.processStereo(this.handle, this.spectra, this.info, this.stereoBytes, this.tnsLen ? this.tnsBytes : null,
                                this.tnsLen ? this.tnsOffsets : null, pcm, t);
        else
            addon.process(this.handle, this.spectra, this.info, this.tnsLen ? this.tnsBytes : null,
                          this.tnsLen ? this.tnsOffsets : null, pcm, t);
        return pcm;                                     // t frames of decoder.js:204-215 output, back to back
    };
});

module.exports = B200Decoder;
