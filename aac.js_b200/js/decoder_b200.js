/*
 * decoder_b200.js -- AACDecoder with the filterbank-synthesis path on a B200.
 *
 * Keeps the Aurora.js Decoder plugin surface of aac.js src/decoder.js (init /
 * setCookie / readChunk returning interleaved PCM).  The serial ADTS / Huffman / ICS parse
 * stays exactly where it is (the reference's own code, unchanged, on the CPU).  What changes
 * is the tail of the frame (src/decoder.js:263-269 / 309-319 and :204-213): instead of running
 * tns.process + filter_bank.process + the interleave per frame in JS, readChunk parses up to K
 * frames ahead, stages them into page-locked typed arrays and makes ONE library call; it
 * returns K*1024*channels samples.  K = 1 is the PR-1 style correctness path.
 *
 * What is staged per channel-frame (options, all per decoder instance):
 *   quantOnDevice (default)  the Huffman-decoded integers + one 16-bit code per band (an
 *       aacfb_qframe record, quant_pack.js): the inverse quantisation / scalefactors / PNS of
 *       ICStream.decodeSpectralData (src/ics.js:203-266) run on the device -- 2304 instead of
 *       4096 bytes over PCIe.  While this decoder parses, ICStream.prototype.decodeSpectralData is
 *       swapped for quant_pack's walk (same Huffman calls in the same order, no arithmetic).
 *   otherwise                ics.data, the dequantised Float32 spectrum, as round 1 did.
 *   stereoOnDevice (default, 2-channel streams)  M/S and intensity stereo become a 768-byte op
 *       record per frame (stereo_pack.js); otherwise the reference's processMS / processIS run here.
 *   pcmFormat 'f32' (default: readChunk's Float32 / 32768) or 's16' (Int16Array).
 *
 * Everything the reference's process() checks or applies and the device does not do takes the
 * reference's OWN CPU path for that frame (cpuFrame below): coupling channel elements
 * (decoder.js:159-163,261-274 -- they never appear in `elements`, they sit in this.cces), frames
 * whose elements do not cover every channel (the reference leaves those channels silent and
 * their overlap untouched), and the profile / gain-control / SBR throws of processSingle /
 * processPair fire exactly as in the reference.  The overlap state is handed over in both
 * directions (addon.getOverlap / setOverlap), so the stream stays continuous.
 */
var AV = require('av');
var AACDecoder = require('aac/src/decoder');      // the unmodified reference
var ICStream = require('aac/src/ics');
var CPEElement = require('aac/src/cpe');
var addon = require('./build/Release/aacfb.node');
var tnsPack = require('./tns_pack');
var stereoPack = require('./stereo_pack');
var quantPack = require('./quant_pack');

var AOT_AAC_MAIN = 1, AOT_AAC_LTP = 4;             // decoder.js:59-62
var IN_F32 = 0, IN_Q16 = 1, PCM_F32 = 0, PCM_S16 = 1;   // include/aacfb.h

var B200Decoder = AACDecoder.extend(function() {
    AV.Decoder.register('mp4a', this);
    AV.Decoder.register('aac ', this);

    this.prototype.framesPerChunk = 64;            // K
    this.prototype.tnsMode = tnsPack.AS_SHIPPED;   // literal parity with the reference by default
    this.prototype.stereoOnDevice = true;          // M/S + IS applied by the kernels (channel pairs of 2-channel streams)
    this.prototype.quantOnDevice = true;           // inverse quantisation on the device (aacfb_qframe input)
    this.prototype.pcmFormat = 'f32';

    // Page-locked staging when the addon offers it (aacfb_host_alloc behind an external ArrayBuffer):
    // the copies of aacfb_process* run at the PCIe rate only from such memory.
    function alloc(bytes) {
        return addon.allocPinned ? addon.allocPinned(bytes) : new ArrayBuffer(bytes);
    }

    var setCookie = AACDecoder.prototype.setCookie;
    this.prototype.setCookie = function(buffer) {
        setCookie.call(this, buffer);              // parses the config, throws like the reference
        var C = this.config.chanConfig, K = this.framesPerChunk;
        this.handle = addon.create(0, 1, C, this.config.sampleIndex, 0, this.tnsMode);
        this.inputBuf = alloc(K * C * 4096);       // Float32 rows or 2304-byte aacfb_qframe records
        this.spectra = new Float32Array(this.inputBuf);
        this.qBytes = new Uint8Array(this.inputBuf);
        this.qView = new DataView(this.inputBuf);
        this.info = new Uint8Array(alloc(K * C * 8));
        this.tnsOffsets = new Uint32Array(alloc((K * C + 1) * 4));
        this.tnsBuf = alloc(K * C * (8 + 8 * 4 * 84));
        this.tnsBytes = new Uint8Array(this.tnsBuf);
        this.tnsView = new DataView(this.tnsBuf);
        this.deviceStereo = this.stereoOnDevice && C === 2;
        // quantised staging leaves no spectra on the host, so it needs the stereo tools on the device too:
        // mono streams and 2-channel streams with deviceStereo; everything else stages Float32 spectra
        this.useQuant = this.quantOnDevice && (C === 1 || this.deviceStereo);
        if (this.deviceStereo) {                   // one aacfb_stereo_ops record per frame
            this.stereoBuf = alloc(K * stereoPack.RECORD_BYTES);
            this.stereoBytes = new Uint8Array(this.stereoBuf);
        }
        this.pcmBuf = alloc(K * C * 4096);         // the library writes here; readChunk hands out a copy
        this.overlapBuf = new ArrayBuffer(C * 4096);
        this.overlapState = new Float32Array(this.overlapBuf);
        this.adtsOut = new Uint32Array(3 * (K + 1));
    };

    // the reference's own refusals (decoder.js:256-259,277-280,296,303,324-332), before anything is staged
    this.prototype.guard = function(ics) {
        var profile = this.config.profile;
        if (profile === AOT_AAC_MAIN) throw new Error("Main prediction unimplemented");
        if (profile === AOT_AAC_LTP) throw new Error("LTP prediction unimplemented");
        if (ics.gainPresent) throw new Error("Gain control not implemented");
        if (this.sbrPresent) throw new Error("SBR not implemented");
    };

    // The part of process(elements) before TNS: M/S, IS (decoder.js:295-302), then stage frame t.
    // Returns false when the frame needs the reference's CPU path instead (nothing is staged then).
    this.prototype.stage = function(elements, t) {
        var C = this.config.chanConfig, channel = 0, self = this, i, e;
        if (this.cces && this.cces.length > 0) return false;        // coupling: decoder.js:261,266,274,307,315
        for (i = 0; i < elements.length && channel < C; i++) {       // the elements must cover every channel
            e = elements[i];
            if (e instanceof ICStream) channel += 1;
            else if (e instanceof CPEElement) channel += 2;
            else throw new Error("Unknown element found.");          // decoder.js:246
        }
        if (channel !== C) return false;
        function put(ics, ch) {
            var cf = t * C + ch;
            self.guard(ics);
            if (self.useQuant) quantPack.pack(ics, self.qBytes, self.qView, cf * quantPack.RECORD_BYTES);
            else self.spectra.set(ics.data, cf * 1024);
            var b = self.info, o = cf * 8;
            b[o] = ics.info.windowSequence; b[o + 1] = ics.info.windowShape[0]; b[o + 2] = ics.info.windowShape[1];
            b[o + 3] = ics.info.maxSFB; b[o + 4] = ics.tnsPresent ? 1 : 0; b[o + 5] = b[o + 6] = b[o + 7] = 0;
            self.tnsOffsets[cf] = self.tnsLen;
            if (ics.tnsPresent) self.tnsLen = (tnsPack.writeBlock(ics.tns, self.tnsBytes, self.tnsView, self.tnsLen) + 3) & ~3;
        }
        channel = 0;
        for (i = 0; i < elements.length && channel < C; i++) {
            e = elements[i];
            if (e instanceof ICStream) { put(e, channel); channel += 1; }
            else {
                var onDevice = false;
                if (this.deviceStereo) {
                    var at = (this.stereoBase || 0) + t * stereoPack.RECORD_BYTES;   // (stereoBase: decoder_pool.js)
                    onDevice = stereoPack.pack(e, new Uint8Array(this.stereoBuf, at, 256),
                                               new Float32Array(this.stereoBuf, at + 256, 128));
                    if (onDevice) this.anyStereo = true;
                } else {
                    // (useQuant is off in this case: the spectra exist on the host)
                    if (e.commonWindow && e.maskPresent) this.processMS(e, e.left.data, e.right.data);
                    this.processIS(e, e.left.data, e.right.data);
                }
                put(e.left, channel); put(e.right, channel + 1);
                if (onDevice) this.info[(t * C + channel) * 8 + 5] = 1;   // stereo_present, left channel
                channel += 2;
            }
        }
        return true;
    };

    // The bit parse of ONE access unit, done by the reference itself: its readChunk (decoder.js:125-216)
    // reads the ADTS header, parses the elements, aligns, and then calls this.process(elements) --
    // which is intercepted here to capture the parsed elements instead of running TNS / filterbank on
    // the CPU; the interleave that follows finds no channel data and returns an empty array.  Nothing
    // in the reference has to be modified or refactored for this.  With quantOnDevice the inverse
    // quantisation is taken out of the parse for its duration (quant_pack.decodeSpectralData).
    var referenceReadChunk = AACDecoder.prototype.readChunk;
    this.prototype.parseElements = function() {
        var captured = null, self = this, own = this.process;
        var stock = ICStream.prototype.decodeSpectralData;
        this.process = function(elements) { captured = elements; self.data = []; };
        if (this.useQuant) ICStream.prototype.decodeSpectralData = quantPack.decodeSpectralData;
        try { referenceReadChunk.call(this); }
        finally { this.process = own; ICStream.prototype.decodeSpectralData = stock; }
        return captured;
    };

    // How many complete access units the buffered bytes hold, from their ADTS headers alone
    // (addon.adtsIndex = aacfb_adts_index = ADTSDemuxer.readHeader hopping by frameLength,
    // adts_demuxer.js:28-52): readChunk then parses exactly that many instead of probing for the
    // underflow.  Raw (non-ADTS) streams and hosts without the byte-level stream API keep probing.
    this.prototype.framesAvailable = function(K) {
        var stream = this.bitstream, s = stream.stream;
        if (!addon.adtsIndex || !s || !s.peekBuffer || !s.remainingBytes) return K;
        if (!stream.available(12) || stream.peek(12) !== 0xfff) return K;
        var n = Math.min(s.remainingBytes(), K * 8191 + 9);
        var count = addon.adtsIndex(s.peekBuffer(0, n).data, this.adtsOut);
        return Math.min(K, Math.max(count, 1));     // none complete: let the parse underflow like the reference's
    };

    // One library call for the t frames staged so far; returns their PCM as a fresh typed array.
    this.prototype.flush = function(t) {
        var C = this.config.chanConfig, n = t * 1024 * C, s16 = this.pcmFormat === 's16';
        var out = s16 ? new Int16Array(n) : new Float32Array(n);
        if (t === 0) return out;
        this.tnsOffsets[t * C] = this.tnsLen;
        var staged = s16 ? new Int16Array(this.pcmBuf, 0, n) : new Float32Array(this.pcmBuf, 0, n);
        var tb = this.tnsLen ? this.tnsBytes : null, to = this.tnsLen ? this.tnsOffsets : null;
        if (this.useQuant || s16)
            addon.processIo(this.handle, this.useQuant ? this.qBytes : this.spectra, this.useQuant ? IN_Q16 : IN_F32,
                            this.info, this.anyStereo ? this.stereoBytes : null, tb, to, staged, s16 ? PCM_S16 : PCM_F32, t);
        else if (this.anyStereo)
            addon.processStereo(this.handle, this.spectra, this.info, this.stereoBytes, tb, to, staged, t);
        else
            addon.process(this.handle, this.spectra, this.info, tb, to, staged, t);
        out.set(staged);
        this.tnsLen = 0;
        this.anyStereo = false;
        return out;
    };

    // One access unit through the reference's own process() on the CPU (coupling, uncovered channels):
    // the device's overlap state goes into this.filter_bank.overlaps and back, so the stream continues
    // seamlessly on either side.  The frame is parsed again from `mark` with the stock
    // decodeSpectralData, because the quantised parse left no spectra on the host.
    this.prototype.cpuFrame = function(mark) {
        var C = this.config.chanConfig, ov = this.overlapState, fb = this.filter_bank, ch;
        this.bitstream.seek(mark);
        addon.getOverlap(this.handle, ov);
        for (ch = 0; ch < C; ch++) fb.overlaps[ch].set(new Float32Array(this.overlapBuf, ch * 4096, 1024));
        var pcm = referenceReadChunk.call(this);    // stock parse + process + interleave (decoder.js:125-216)
        for (ch = 0; ch < C; ch++) ov.set(fb.overlaps[ch], ch * 1024);
        addon.setOverlap(this.handle, ov);
        if (this.pcmFormat !== 's16') return pcm;
        var out = new Int16Array(pcm.length);       // AACFB_PCM_S16 of the un-normalised sample (include/aacfb.h)
        for (var i = 0; i < pcm.length; i++) out[i] = Math.max(-32768, Math.min(32767, Math.round(pcm[i] * 32768)));
        return out;
    };

    function concat(a, b) {
        if (a.length === 0) return b;
        var out = (a instanceof Int16Array) ? new Int16Array(a.length + b.length) : new Float32Array(a.length + b.length);
        out.set(a); out.set(b, a.length);
        return out;
    }

    this.prototype.readChunk = function() {
        var K = this.framesPerChunk, t = 0, done = 0, stream = this.bitstream, pcm = null;
        this.tnsLen = 0;
        this.anyStereo = false;
        var budget = this.framesAvailable(K);
        while (done + t < budget) {
            var mark = stream.offset();
            try {
                var elements = this.parseElements();   // the reference's own parse (decoder.js:129-200)
                if (this.stage(elements, t)) { t++; continue; }
            } catch (err) {
                // Underflow: the access unit is not complete yet.  Anything else: the reference would have
                // emitted the frames before it first -- do that, and let the next call hit the error again.
                if (done + t === 0) throw err;
                stream.seek(mark);
                break;
            }
            // this frame takes the reference's CPU path: flush what is staged, run it, carry on
            var head = this.flush(t);
            pcm = concat(pcm === null ? head : concat(pcm, head), this.cpuFrame(mark));
            done += t + 1;
            t = 0;
        }
        var tail = this.flush(t);
        return pcm === null ? tail : concat(pcm, tail);   // frames of decoder.js:204-215 output, back to back
    };
});

module.exports = B200Decoder;
