/*
 * decoder_pool.js -- S decoders in lock step behind ONE library context.
 *
 * decoder_b200.js is one Aurora decoder = one stream: K frames per library call, which is
 * latency-bound (INTEGRATION.md section 3h).  The library's batch is [S][T][C]; a host that decodes
 * many streams at once (a transcoding farm) fills it with this pool:
 *
 *     var pool = new DecoderPool(S, {framesPerChunk: 64, pcmFormat: 's16'});
 *     var d = pool.createDecoder();      // S times; each is an AACDecoder as far as Aurora can tell:
 *     d.setCookie(cookie);               //   the host wires demuxer / bitstream to it as it does for
 *     d.bitstream = ...;                 //   the reference's decoder
 *     var pcm = pool.readChunks();       // [S] typed arrays of T * 1024 * channels samples, or null
 *
 * readChunks parses T <= K access units of EVERY stream -- the reference's own bit parse, exactly as
 * decoder_b200.js runs it (parseElements / stage are inherited, they stage into this pool's arrays
 * through views) -- and makes one aacfb_process* call for all S * T frames on a context created
 * with n_streams = S.  T is what every stream can deliver: the smallest count of complete access
 * units (ADTS index), and if a parse underflows nevertheless every stream is rewound and the round
 * is redone with the shorter T; with no complete frame in some stream the call returns null and
 * nothing has been consumed.
 *
 * Limits, stated: all streams share channel configuration and sample rate (one context); a frame
 * that needs the reference's CPU path (coupling channel elements, elements that do not cover every
 * channel -- decoder_b200.js cpuFrame) is an Error naming the stream: decode that stream with a
 * decoder of its own.
 */
var AV = require('av');
var AACDecoder = require('aac/src/decoder');      // the unmodified reference
var B200Decoder = require('./decoder_b200');
var addon = require('./build/Release/aacfb.node');
var stereoPack = require('./stereo_pack');
var quantPack = require('./quant_pack');

var IN_F32 = 0, IN_Q16 = 1, PCM_F32 = 0, PCM_S16 = 1;   // include/aacfb.h

function alloc(bytes) {
    return addon.allocPinned ? addon.allocPinned(bytes) : new ArrayBuffer(bytes);
}

// A member parses and stages like a B200Decoder but owns neither a context nor staging arrays.
var Member = B200Decoder.extend(function() {
    this.prototype.setCookie = function(buffer) {
        AACDecoder.prototype.setCookie.call(this, buffer);   // parses the config, throws like the reference
        this.pool.attach(this);
    };
    this.prototype.readChunk = function() {
        throw new Error("a pooled decoder is read through DecoderPool.readChunks");
    };
});

function DecoderPool(nStreams, options) {
    options = options || {};
    this.S = nStreams;
    this.K = options.framesPerChunk || 64;
    this.tnsMode = options.tnsMode || 0;                 // tns_pack.AS_SHIPPED
    this.stereoOnDevice = options.stereoOnDevice !== false;
    this.quantOnDevice = options.quantOnDevice !== false;
    this.pcmFormat = options.pcmFormat || 'f32';
    this.members = [];
    this.attached = 0;
    this.chanConfig = 0;
    this.sampleIndex = 0;
    this.handle = null;
}

DecoderPool.prototype.createDecoder = function() {
    if (this.members.length === this.S) throw new Error("DecoderPool: all " + this.S + " decoders exist");
    var d = new Member();
    d.pool = this;
    d.index = this.members.length;
    d.framesPerChunk = this.K;
    d.pcmFormat = this.pcmFormat;
    d.adtsOut = new Uint32Array(3 * (this.K + 1));     // framesAvailable's scratch (decoder_b200.js)
    this.members.push(d);
    return d;
};

// Called from a member's setCookie: one context for all, once every stream's configuration is known.
DecoderPool.prototype.attach = function(d) {
    var C = d.config.chanConfig, si = d.config.sampleIndex;
    if (this.attached === 0) { this.chanConfig = C; this.sampleIndex = si; }
    else if (C !== this.chanConfig || si !== this.sampleIndex)
        throw new Error("DecoderPool: stream " + d.index + " differs in channel configuration or sample rate");
    this.attached += 1;
    if (this.attached < this.S) return;
    var S = this.S, K = this.K, n = S * K * C;
    this.handle = addon.create(0, S, C, si, 0, this.tnsMode);
    this.inputBuf = alloc(n * 4096);               // Float32 rows or 2304-byte aacfb_qframe records
    this.infoBuf = alloc(n * 8);
    this.tnsOffBuf = alloc((n + 1) * 4);
    this.tnsBuf = alloc(n * (8 + 8 * 4 * 84));
    this.tnsBytes = new Uint8Array(this.tnsBuf);
    this.tnsView = new DataView(this.tnsBuf);
    this.deviceStereo = this.stereoOnDevice && C === 2;
    this.useQuant = this.quantOnDevice && (C === 1 || this.deviceStereo);   // as in decoder_b200.js
    this.stereoBuf = this.deviceStereo ? alloc(S * K * stereoPack.RECORD_BYTES) : null;
    this.pcmBuf = alloc(n * 4096);
};

// Point member s at its [T][C] slice of the batch [S][T][C].
DecoderPool.prototype.bind = function(d, T) {
    var C = this.chanConfig, cf0 = d.index * T * C, rec = this.useQuant ? quantPack.RECORD_BYTES : 4096;
    d.useQuant = this.useQuant;
    d.deviceStereo = this.deviceStereo;
    d.spectra = new Float32Array(this.inputBuf, cf0 * 4096, T * C * 1024);   // (Float32 staging: 4096-byte rows)
    d.qBytes = new Uint8Array(this.inputBuf, cf0 * rec, T * C * rec);
    d.qView = new DataView(this.inputBuf, cf0 * rec, T * C * rec);
    d.info = new Uint8Array(this.infoBuf, cf0 * 8, T * C * 8);
    d.tnsOffsets = new Uint32Array(this.tnsOffBuf, cf0 * 4, T * C + 1);
    d.tnsBytes = this.tnsBytes;                    // one blob for the batch: offsets are absolute
    d.tnsView = this.tnsView;
    d.stereoBuf = this.stereoBuf;
    d.stereoBase = d.index * T * stereoPack.RECORD_BYTES;
    d.anyStereo = false;
};

// Parse and stage T frames of every stream.  Returns T on success; on a short stream every bitstream is back
// at its mark and the number of frames that stream could deliver (< T) is returned.
DecoderPool.prototype.stageRound = function(T, marks) {
    var S = this.S, s, t, d, tnsLen = 0, anyStereo = false;
    for (s = 0; s < S; s++) {
        d = this.members[s];
        this.bind(d, T);
        d.tnsLen = tnsLen;
        for (t = 0; t < T; t++) {
            var ok;
            try {
                ok = d.stage(d.parseElements(), t);
            } catch (err) {
                for (var r = 0; r <= s; r++) this.members[r].bitstream.seek(marks[r]);
                if (t > 0 && err instanceof AV.UnderflowError) return t;
                err.streamIndex = s;
                throw err;
            }
            if (!ok) {
                for (var q = 0; q <= s; q++) this.members[q].bitstream.seek(marks[q]);
                var e = new Error("DecoderPool: stream " + s + " has a frame that needs the reference's CPU path " +
                                  "(coupling / uncovered channels): decode it with a B200Decoder of its own");
                e.streamIndex = s;
                throw e;
            }
        }
        tnsLen = d.tnsLen;
        anyStereo = anyStereo || d.anyStereo;
    }
    this.tnsLen = tnsLen;
    this.anyStereo = anyStereo;
    return T;
};

// One round: [S] typed arrays (T * 1024 * channels samples each, T the same for all), or null when some
// stream has no complete access unit buffered (nothing is consumed then).
DecoderPool.prototype.readChunks = function() {
    if (this.handle === null) throw new Error("DecoderPool: setCookie has not run on every decoder");
    var S = this.S, C = this.chanConfig, s, T = this.K, marks = [];
    for (s = 0; s < S; s++) {
        var d = this.members[s];
        marks.push(d.bitstream.offset());
        T = Math.min(T, d.framesAvailable(this.K));
    }
    while (true) {
        var got;
        try {
            got = this.stageRound(T, marks);
        } catch (err) {
            if (err instanceof AV.UnderflowError) return null;   // not even one frame of that stream: feed it first
            throw err;
        }
        if (got === T) break;
        T = got;                                    // a stream ran short: the same round again with what it has
    }
    var n = S * T * C, s16 = this.pcmFormat === 's16';
    new Uint32Array(this.tnsOffBuf, 0, n + 1)[n] = this.tnsLen;
    var staged = s16 ? new Int16Array(this.pcmBuf, 0, n * 1024) : new Float32Array(this.pcmBuf, 0, n * 1024);
    var info = new Uint8Array(this.infoBuf, 0, n * 8);
    var tb = this.tnsLen ? this.tnsBytes : null, to = this.tnsLen ? new Uint32Array(this.tnsOffBuf, 0, n + 1) : null;
    var stereo = this.anyStereo ? new Uint8Array(this.stereoBuf, 0, S * T * stereoPack.RECORD_BYTES) : null;
    var input = this.useQuant ? new Uint8Array(this.inputBuf, 0, n * quantPack.RECORD_BYTES)
                              : new Float32Array(this.inputBuf, 0, n * 1024);
    if (this.useQuant || s16)
        addon.processIo(this.handle, input, this.useQuant ? IN_Q16 : IN_F32, info, stereo, tb, to, staged,
                        s16 ? PCM_S16 : PCM_F32, T);
    else if (stereo !== null)
        addon.processStereo(this.handle, input, info, stereo, tb, to, staged, T);
    else
        addon.process(this.handle, input, info, tb, to, staged, T);
    var out = [], per = T * 1024 * C;
    for (s = 0; s < S; s++) {
        var chunk = s16 ? new Int16Array(per) : new Float32Array(per);
        chunk.set(s16 ? new Int16Array(this.pcmBuf, 2 * s * per, per) : new Float32Array(this.pcmBuf, 4 * s * per, per));
        out.push(chunk);
    }
    return out;
};

module.exports = DecoderPool;
