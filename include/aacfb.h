/*
 * aacfb.h -- C ABI of the B200-native AAC-LC filterbank-synthesis path.
 *
 * This is the drop-in boundary for ONE hot path of audiocogs/aac.js:
 *
 *     TNS.process            (reference src/tns.js:105-177)
 *  -> FilterBank.process     (reference src/filter_bank.js:88-204)
 *       -> MDCT.process      (reference src/mdct.js:62-115)
 *            -> FFT.process  (reference src/fft.js:105-192)
 *  -> interleave + /32768    (reference src/decoder.js:204-213)
 *
 * Everything above it (ADTS / Huffman / ICS bit parse, dequantisation) stays in
 * the JavaScript host.  The stereo tools that sit between the parse and TNS
 * (processMS / processIS, decoder.js:337-404) can optionally run on the device
 * too: aacfb_process_stereo.  The entry points below are exactly what an N-API addon
 * for that path binds; they use plain pointers and sizes
 * only -- no torch, no C++ types.  The library has NO CPU fallback: every
 * compute entry point fails with AACFB_ERR_CUDA when no sm_100 device/kernel
 * image is usable.
 *
 * Vocabulary (the reference's): a *frame* is one AAC access unit = 1024
 * samples for each of C channels; a *channel-frame* is one channel of one
 * frame = 1024 dequantised spectral coefficients in, 1024 PCM samples out.
 * A *stream* is one decoder instance (one FilterBank with its per-channel
 * `overlaps`, filter_bank.js:38-41).  A context batches S independent streams
 * of identical channel count that are decoded in lockstep.
 */
#ifndef AACFB_H_
#define AACFB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AACFB_FRAME_LEN 1024   /* config.frameLength, decoder.js:86            */
#define AACFB_SHORT_LEN 128    /* FilterBank.shortLength, filter_bank.js:30    */
#define AACFB_MAX_CHANNELS 8
#define AACFB_TNS_MAX_ORDER 20 /* tns.js:46                                    */

/* ICStream window sequences, reference src/ics.js:44-47 */
enum {
    AACFB_ONLY_LONG_SEQUENCE   = 0,
    AACFB_LONG_START_SEQUENCE  = 1,
    AACFB_EIGHT_SHORT_SEQUENCE = 2,
    AACFB_LONG_STOP_SEQUENCE   = 3
};

/* window shapes: index into LONG_WINDOWS / SHORT_WINDOWS, filter_bank.js:85-86 */
enum { AACFB_SHAPE_SINE = 0, AACFB_SHAPE_KBD = 1 };

/* error codes (negative); 0 = success */
enum {
    AACFB_OK            =  0,
    AACFB_ERR_ARG       = -1,  /* bad argument (null, range)                    */
    AACFB_ERR_SMALL     = -2,  /* smallFrames requested: filter_bank.js:25-27   */
    AACFB_ERR_SEQUENCE  = -3,  /* window_sequence > 3                           */
    AACFB_ERR_TNS       = -4,  /* TNS order > 20 (tns.js:84-85) or bad blob     */
    AACFB_ERR_CUDA      = -5,  /* CUDA runtime failure / no usable device       */
    AACFB_ERR_NOMEM     = -6,
    AACFB_ERR_ADTS      = -7   /* "Invalid ADTS header." adts_demuxer.js:29-30  */
};

/* create() flags ---------------------------------------------------------- */
/* TNS behaviour (reference defect C1, tns.js:122 reads `tmp` for `top`):
 *   AS_SHIPPED : identity, literally what the reference does today.
 *   FIXED_AR   : `tmp`->`top`, decode=true  -> all-pole filter tns.js:156-162
 *   FIXED_MA   : `tmp`->`top`, decode=false -> FIR branch     tns.js:163-174
 *                (what decoder.js:264,310,313 would select once C1 is fixed) */
#define AACFB_TNS_AS_SHIPPED 0u
#define AACFB_TNS_FIXED_AR   1u
#define AACFB_TNS_FIXED_MA   2u
#define AACFB_TNS_MODE_MASK  3u

/* One per channel-frame, laid out [S][T][C].  Mirrors the ICSInfo fields that
 * FilterBank.process and TNS.process read (filter_bank.js:89-91,104;
 * tns.js:106,113; ics.js:278-307). */
typedef struct aacfb_frame_info {
    uint8_t window_sequence; /* info.windowSequence, 0..3                       */
    uint8_t shape_prev;      /* info.windowShape[0]                              */
    uint8_t shape_cur;       /* info.windowShape[1]                              */
    uint8_t max_sfb;         /* ics.maxSFB (TNS band clamp, tns.js:106)          */
    uint8_t tns_present;     /* ics.tnsPresent (decoder.js:263,309,312)          */
    uint8_t stereo_present;  /* LEFT channel of a pair only: the pair-frame has an
                              * aacfb_stereo_ops record (M/S and/or intensity)    */
    uint8_t reserved[2];     /* must be 0                                        */
} aacfb_frame_info;

/* Stereo tools of a channel pair element, applied to the dequantised spectra of
 * (left, right) BEFORE TNS:
 *     mid/side          processMS, reference src/decoder.js:379-404
 *     intensity stereo  processIS, reference src/decoder.js:337-376
 * The host keeps the band walk (ms_used / bandTypes / sectEnd / scaleFactors are
 * bit-parse results) and hands over WHAT to do per group of 4 coefficients --
 * scalefactor-band edges are multiples of 4 for every sample rate and window
 * length (tables.js:34-124) -- instead of touching the 2 x 1024 coefficients:
 *     op[i] (coefficients 4i .. 4i+3):  0  untouched
 *                                       1  l' = l + r, r' = l - r      (decoder.js:395-397)
 *                                       2+k  r' = l * scale[k]         (decoder.js:360-366,
 *                                            scale = c * scaleFactors[idx])
 * Supported for pairs that are channels (2j, 2j+1) of a stream with an even
 * channel count (stereo: the CPE).  One record per channel PAIR-frame, laid out
 * [S][T][C/2]; only records whose left channel has stereo_present != 0 are read. */
#define AACFB_STEREO_NONE 0
#define AACFB_STEREO_MS   1
#define AACFB_STEREO_IS   2   /* + index into scale[] */
typedef struct aacfb_stereo_ops {
    uint8_t op[256];
    float   scale[128];
} aacfb_stereo_ops;             /* 768 bytes */

/* TNS side info is a packed blob, one block per channel-frame that has
 * tns_present != 0, located by tns_offsets[] (byte offsets, 4-byte aligned,
 * [S*T*C + 1] entries; an empty block has offsets[i+1]==offsets[i]).
 * Block layout (mirrors TNS.nFilt/length/order/direction/coef, tns.js:24-40):
 *     uint8_t n_filt[8];                     // per window w (only w<windowCount used)
 *     then for w in 0..7, for filt in 0..n_filt[w]-1:
 *         aacfb_tns_filter hdr;              // 4 bytes
 *         float coef[hdr.order];             // the dequantised reflection
 *                                            // coefficients, tns.js:97      */
typedef struct aacfb_tns_filter {
    uint8_t length;     /* length[w][filt]    in scalefactor bands              */
    uint8_t order;      /* order[w][filt]     0..20                             */
    uint8_t direction;  /* direction[w][filt] 0 = upward, 1 = downward          */
    uint8_t reserved;
} aacfb_tns_filter;

/* ---- inverse quantisation + scalefactors + PNS on the device ---------------
 * ICStream.decodeSpectralData (reference src/ics.js:203-266) interleaves the Huffman decode with
 *     data[i] = +-IQ_TABLE[|q|];  data[i] *= scaleFactors[idx]          (ics.js:247-254)
 * A host that ships the Huffman-decoded INTEGERS and the per-band scalefactor indices instead of
 * the Float32 spectrum moves 2304 instead of 4104 bytes per channel-frame over PCIe; the device
 * dequantises the row while it sits in shared memory.  One record per channel-frame, [S][T][C]:
 *   q[i]        buf[j] of ics.js:247 at data index i (window-grouped order, as the reference
 *               stores it); |q| <= 8191 (IQ_TABLE has 8191 entries: |q| = 8191 reads `undefined`
 *               there -> NaN, reproduced).  Coefficients of ZERO / intensity / noise bands and
 *               everything at or above swbOffsets[maxSFB] are ignored.
 *   band[idx]   idx = g * maxSFB + sfb as in ics.js:213-219: kind | index i into
 *               SCALEFACTOR_TABLE (tables.js:168-176, 428 entries; i > 427 = `undefined` -> NaN)
 *                 AACFB_BAND_ZERO      ZERO_BT, INTENSITY_BT, INTENSITY_BT2: data = 0   (ics.js:222-227)
 *                 AACFB_BAND_SPECTRAL  scaleFactors[idx] =  SCALEFACTOR_TABLE[i]          (ics.js:166-172)
 *                 AACFB_BAND_NOISE     scaleFactors[idx] = -SCALEFACTOR_TABLE[i]; perceptual noise
 *                                      substitution AS SHIPPED (ics.js:228-242): the generator
 *                                      randomState = (randomState * (1664525 + 1013904223))|0 starts
 *                                      at 0x1F2E3D4C in every new ICStream (one per element per frame,
 *                                      decoder.js:146,154), yields 11 non-zero values and then 0 for
 *                                      ever, so all but the first noise coefficients of a
 *                                      channel-frame become 0 * (sf / sqrt(0)) = NaN.  Reproduced.
 *   group_len   info.groupLength[0 .. groupCount-1], zero-terminated (ONLY_LONG etc.: {1})
 * window_sequence and maxSFB come from the channel-frame's aacfb_frame_info. */
#define AACFB_BAND_ZERO       0x0000u
#define AACFB_BAND_SPECTRAL   0x4000u
#define AACFB_BAND_NOISE      0x8000u
#define AACFB_BAND_KIND_MASK  0xc000u
#define AACFB_BAND_INDEX_MASK 0x01ffu
#define AACFB_BAND_UNDEFINED  0x01ffu   /* scalefactor index outside the table: NaN           */
typedef struct aacfb_qframe {
    uint8_t  group_len[8];
    uint16_t band[120];          /* MAX_SECTIONS, ics.js:49                                */
    uint8_t  reserved[8];        /* must be 0                                              */
    int16_t  q[1024];
} aacfb_qframe;                  /* 2304 bytes, 16-byte aligned rows                        */

/* PCM sample formats of the *_io entry points.  S16 is what Aurora's sinks
 * make of readChunk's Float32 output OUTSIDE the reference tree (parity unpinned there); here it
 * is defined as  Int16Array[i] = max(-32768, min(32767, Math.round(data[ch][k])))  on the
 * un-normalised sample (decoder.js:210 divides by 32768 only to normalise): round half up,
 * saturate, NaN -> 0.  Pinned by known-answer tests. */
#define AACFB_PCM_F32 0u
#define AACFB_PCM_S16 1u
/* input formats */
#define AACFB_IN_F32  0u         /* float spectra [S][T][C][1024] (ics.data)               */
#define AACFB_IN_Q16  1u         /* aacfb_qframe  [S][T][C]                                */

typedef struct aacfb_ctx aacfb_ctx;

/* Construct.  Replaces `new FilterBank(smallFrames, channels)`
 * (filter_bank.js:24-44, called at decoder.js:112) for `n_streams` decoder
 * instances at once, and carries config.sampleIndex for TNS (tns.js:23).
 * Overlap state is zero-initialised (filter_bank.js:38-41). */
int aacfb_create(aacfb_ctx **out, int device, int n_streams, int channels,
                 int sample_index, int small_frames, uint32_t flags);
int aacfb_destroy(aacfb_ctx *ctx);

/* Zero the overlap state of every stream/channel. */
int aacfb_reset(aacfb_ctx *ctx);

/* The batched hot path with HOST buffers (what the N-API addon calls from the
 * decoder's readChunk).  For each stream s, frame t, channel c:
 *     TNS.process (per ctx TNS mode)  -> FilterBank.process -> interleave,/32768
 *   spectra     [S][T][C][1024] float  (ics.data after M/S, IS, coupling)
 *   info        [S][T][C]
 *   tns_blob / tns_offsets: see above; both may be NULL when no frame has TNS
 *   pcm         [S][T][1024][C] float  (readChunk's return value, per frame)
 * Blocking: returns when pcm is complete.  Does not modify spectra. */
int aacfb_process(aacfb_ctx *ctx, const float *spectra,
                  const aacfb_frame_info *info,
                  const uint8_t *tns_blob, const uint32_t *tns_offsets,
                  float *pcm, int n_frames);

/* aacfb_process with the stereo tools of the pair elements run on the device as
 * well: spectra are ics.data BEFORE processMS / processIS (decoder.js:300-307),
 * stereo_ops [S][T][C/2] as described above (NULL = none: plain aacfb_process). */
int aacfb_process_stereo(aacfb_ctx *ctx, const float *spectra,
                         const aacfb_frame_info *info,
                         const aacfb_stereo_ops *stereo_ops,
                         const uint8_t *tns_blob, const uint32_t *tns_offsets,
                         float *pcm, int n_frames);

/* Same contract with DEVICE pointers, enqueued on `stream` (a cudaStream_t
 * passed as void*; NULL = the legacy default stream).  Asynchronous.
 * tns_blob_bytes is the size of the device blob (0 if none). */
int aacfb_process_device(aacfb_ctx *ctx, const float *d_spectra,
                         const aacfb_frame_info *d_info,
                         const uint8_t *d_tns_blob, const uint32_t *d_tns_offsets,
                         size_t tns_blob_bytes,
                         float *d_pcm, int n_frames, void *stream);

int aacfb_process_device_stereo(aacfb_ctx *ctx, const float *d_spectra,
                                const aacfb_frame_info *d_info,
                                const aacfb_stereo_ops *d_stereo_ops,
                                const uint8_t *d_tns_blob, const uint32_t *d_tns_offsets,
                                size_t tns_blob_bytes,
                                float *d_pcm, int n_frames, void *stream);

/* The general form of the four calls above: `input` is float spectra (AACFB_IN_F32) or
 * aacfb_qframe records (AACFB_IN_Q16: the device runs ICStream.decodeSpectralData's inverse
 * quantisation, ics.js:203-266, first); `pcm` receives [S][T][1024][C] samples as Float32 / 32768
 * (AACFB_PCM_F32, readChunk's output) or as int16 (AACFB_PCM_S16).  stereo_ops, tns_blob and
 * tns_offsets may be NULL.  Host buffers, blocking. */
int aacfb_process_io(aacfb_ctx *ctx, const void *input, uint32_t in_format,
                     const aacfb_frame_info *info,
                     const aacfb_stereo_ops *stereo_ops,
                     const uint8_t *tns_blob, const uint32_t *tns_offsets,
                     void *pcm, uint32_t pcm_format, int n_frames);
/* Same with DEVICE pointers, asynchronous on `stream`. */
int aacfb_process_device_io(aacfb_ctx *ctx, const void *d_input, uint32_t in_format,
                            const aacfb_frame_info *d_info,
                            const aacfb_stereo_ops *d_stereo_ops,
                            const uint8_t *d_tns_blob, const uint32_t *d_tns_offsets,
                            size_t tns_blob_bytes,
                            void *d_pcm, uint32_t pcm_format, int n_frames, void *stream);

/* Page-locked host memory for the buffers of the host-buffer calls.  The copies of
 * aacfb_process* run at the PCIe rate only from page-locked memory; a host that keeps its
 * staging arrays for the life of the decoder (the N-API addon's typed arrays) either allocates
 * them here (aacfb_host_alloc: the addon wraps the pointer in an external ArrayBuffer) or
 * registers existing memory once (aacfb_host_register; `bytes` > 0, any alignment).  Buffers
 * that are neither still work (pageable copies, slower). */
void *aacfb_host_alloc(size_t bytes);
int aacfb_host_free(void *p);
int aacfb_host_register(void *p, size_t bytes);
int aacfb_host_unregister(void *p);

/* The inner seam, one channel-frame at a time, HOST buffers:
 *     filterBank.process(info, input, output, channel)  filter_bank.js:88
 * `output` is the 1024 un-scaled, un-interleaved samples of this.data[channel]
 * (decoder.js:269,318-319).  Stream index selects the decoder instance. */
int aacfb_filterbank_process(aacfb_ctx *ctx, int stream, int channel,
                             const aacfb_frame_info *info,
                             const float *input, float *output);

/* The inner seam  tns.process(ics, data, decode)  tns.js:105: filters `data`
 * (1024 floats, HOST) in place.  `mode` is one of AACFB_TNS_*; `tns_block`
 * is one block of the blob format above (`block_bytes` long). */
int aacfb_tns_process(aacfb_ctx *ctx, const aacfb_frame_info *info,
                      const uint8_t *tns_block, size_t block_bytes,
                      float *data, uint32_t mode);

/* Overlap state hand-off, HOST buffers [S][C][1024] (FilterBank.overlaps). */
int aacfb_get_overlap(aacfb_ctx *ctx, float *overlap);
int aacfb_set_overlap(aacfb_ctx *ctx, const float *overlap);

/* Introspection */
const char *aacfb_last_error(const aacfb_ctx *ctx); /* ctx may be NULL: global */
int aacfb_version(void);
/* number of kernels this context has launched since creation */
uint64_t aacfb_launch_count(const aacfb_ctx *ctx);
/* Copy one of the constant tables the kernels use to a HOST buffer (for
 * table-parity tests against the oracle).  which: 0 FFT roots 512 (re,im)
 * [1024 f32], 1 FFT roots 64 [128], 2 MDCT twiddles 2048 (c,s) [1024],
 * 3 MDCT twiddles 256 [128], 4 sine1024, 5 kbd1024, 6 sine128, 7 kbd128.
 * Returns the number of floats written or a negative error. */
int aacfb_get_table(int which, float *dst, int capacity);
/* which: 8 IQ_TABLE [8192 f32, entry 8191 = NaN for the reference's out-of-table read]
 * (tables.js:181-191), 9 SCALEFACTOR_TABLE [428] (tables.js:168-176), 10 the PNS generator's
 * output as shipped, as floats [32] (ics.js:234-235).  Needs capacity >= 8192 for 8. */
/* Scalefactor-band offsets info.swbOffsets (tables.js:126-154, ics.js:301,307) of one
 * sample rate: is_short = 0 -> SWB_OFFSET_1024[sample_index], 1 -> SWB_OFFSET_128.
 * Returns the number of bands (swbCount); dst receives swbCount + 1 offsets. */
int aacfb_get_swb_offsets(int sample_index, int is_short, uint16_t *dst, int capacity);

/* ADTS frame index (host-side, no device work) -- SURVEY section 8(f) row 3.
 * The reference reads one ADTS header per access unit inside the serial decode
 * loop (decoder.js:129-130 -> ADTSDemuxer.readHeader, adts_demuxer.js:28-52), but
 * the header alone says where the next access unit starts (frameLength), so a
 * host can locate every frame of a buffer WITHOUT entropy decoding and parse the
 * frames of a batch independently (one parser per core) before one aacfb_process
 * call.  Field for field what readHeader returns, plus the byte offset:
 *     12 bits 0xfff | 3 skipped | protectionAbsent | profile-1 (2) | samplingIndex (4)
 *     | 1 skipped | chanConfig (3) | 4 skipped | frameLength (13) | 11 skipped
 *     | numFrames-1 (2) | 16 more bits (CRC) if !protectionAbsent                    */
typedef struct aacfb_adts_frame {
    uint64_t offset;          /* byte offset of the syncword in `data`             */
    uint32_t frame_length;    /* ret.frameLength: bytes to the next syncword        */
    uint8_t  header_bytes;    /* 7, or 9 with CRC (what readHeader consumes)        */
    uint8_t  profile;         /* ret.profile = field + 1                            */
    uint8_t  sampling_index;  /* ret.samplingIndex                                  */
    uint8_t  chan_config;     /* ret.chanConfig                                     */
    uint8_t  num_frames;      /* ret.numFrames = field + 1                          */
    uint8_t  reserved[7];     /* sizeof == 24                                       */
} aacfb_adts_frame;
/* Walks data[0, size) from a syncword at offset 0.  Fills up to `capacity` frames
 * (frames may be NULL to count only) and stops at the first frame that is not
 * completely inside the buffer; *consumed (may be NULL) = offset of that frame =
 * where the next buffer has to start (the batching decoder's rewind point).
 * Returns the number of complete frames, or AACFB_ERR_ADTS when a header does not
 * start with 0xfff or announces a frame shorter than its own header. */
int aacfb_adts_index(const uint8_t *data, size_t size, aacfb_adts_frame *frames, int capacity, size_t *consumed);

#ifdef __cplusplus
}
#endif
#endif /* AACFB_H_ */
