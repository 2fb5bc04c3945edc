/*
 * tns_pack.js -- serialise the reference's TNS object (src/tns.js:22-44: nFilt[8],
 * length[8][4], order[8][4], direction[8][4], coef[8][4][20]) into one block of the
 * aacfb.h TNS blob, and the drop-in  tns.process(ics, data, decode)  (src/tns.js:105).
 */
var addon = require('./build/Release/aacfb.node');

exports.AS_SHIPPED = 0;   // identity: what src/tns.js:122 does today (`tmp` read for `top`)
exports.FIXED_AR = 1;     // decode = true  branch, src/tns.js:156-162
exports.FIXED_MA = 2;     // decode = false branch, src/tns.js:163-174

exports.blockSize = function(tns) {
    var n = 8;
    for (var w = 0; w < 8; w++)
        for (var f = 0; f < tns.nFilt[w]; f++) n += 4 + 4 * tns.order[w][f];
    return n;
};

// writes the block at byte offset `at` (4-byte aligned) of Uint8Array/DataView pair
exports.writeBlock = function(tns, bytes, view, at) {
    for (var w = 0; w < 8; w++) bytes[at + w] = tns.nFilt[w];
    var p = at + 8;
    for (var w = 0; w < 8; w++) {
        for (var f = 0; f < tns.nFilt[w]; f++) {
            var order = tns.order[w][f];
            bytes[p] = tns.length[w][f]; bytes[p + 1] = order; bytes[p + 2] = tns.direction[w][f] ? 1 : 0; bytes[p + 3] = 0;
            for (var i = 0; i < order; i++) view.setFloat32(p + 4 + 4 * i, tns.coef[w][f][i], true);
            p += 4 + 4 * order;
        }
    }
    return p;
};

// drop-in for TNS.prototype.process; `mode` defaults to the reference as shipped
exports.process = function(handle, tns, ics, data, decode, mode) {
    if (mode === undefined) mode = exports.AS_SHIPPED;
    else if (mode !== exports.AS_SHIPPED) mode = decode ? exports.FIXED_AR : exports.FIXED_MA;
    var n = exports.blockSize(tns), buf = new ArrayBuffer(n), bytes = new Uint8Array(buf);
    exports.writeBlock(tns, bytes, new DataView(buf), 0);
    var info = new Uint8Array(8);
    info[0] = ics.info.windowSequence; info[1] = ics.info.windowShape[0]; info[2] = ics.info.windowShape[1];
    // tns.js:106 reads ics.maxSFB, which ICStream never sets (it lives on ics.info): take the evident intent
    info[3] = ics.maxSFB !== undefined ? ics.maxSFB : ics.info.maxSFB; info[4] = 1;
    addon.tnsProcess(handle, info, bytes, data, mode);
};
