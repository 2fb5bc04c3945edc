// aacfb_kernels.h -- launch interface between the C-ABI layer and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/aacfb.h"
#include "aacfb_geometry.h"
#include "aacfb_tables.h"

namespace aacfb {

#ifndef AACFB_WORKERS
#define AACFB_WORKERS 6
#endif
#ifndef AACFB_STAGES
#define AACFB_STAGES 2
#endif
constexpr int kWorkers = AACFB_WORKERS;     // workers per CTA (6 x 64 threads -> 168 registers/thread, no spills)
constexpr int kStages = AACFB_STAGES;       // TMA ring depth per worker
constexpr int kWorkersGeneric = 6;          // generic instantiation (items with EIGHT_SHORT frames)
constexpr int kStagesGeneric = 2;
constexpr int kTnsWarps = 4;                // autonomous warps per tns_kernel CTA, 32 rows each

struct SynthParams {
    const float *spectra;           // [S][T][nc][1024]
    const float *scratch;           // TNS-filtered rows (same layout, valid on `ranges` only) or nullptr
    const uint32_t *ranges;         // [S][T][nc] lo4 | hi4 << 16: the float4 interval of a row that lives in scratch
    const aacfb_frame_info *info;   // [S][T][nc]
    const aacfb_stereo_ops *stereo; // [S][T][nc/2] stereo tools of the pair-frames, or nullptr
    float *pcm;                     // [S][T][1024][nc]
    const float *ovl_in;            // overlap state read by chunks starting at t = 0
    float *ovl_out;                 // overlap state written by chunks ending at t = T
    const SynthTables *tab;         // device copy of the tables
    Geometry g;
    const uint8_t *qframes;         // [S][T][nc] aacfb_qframe records INSTEAD of `spectra` (AACFB_IN_Q16), or nullptr
    const DequantTables *dq;        // device tables of the inverse quantisation (needed with qframes)
    int pcm_s16;                    // 1: `pcm` is int16_t [S][T][1024][nc] (AACFB_PCM_S16; scale must be 1)
    unsigned *counter;              // zeroed before launch (one per kernel instantiation)
    unsigned *short_items;          // zeroed; number of items with EIGHT_SHORT frames (set by the long-only pass)
    float scale;
    // synth_tns_kernel only (TNS filtered inside the synthesis kernel, no pre-pass):
    float *tns_ring;                // per CTA 2 x kFusedRingRows rows of 1024 floats: the filtered rows in flight
    const uint8_t *tns_blob;        // packed TNS blocks + offsets, as in TnsParams
    const uint32_t *tns_offsets;
    size_t tns_blob_bytes;
    const TnsBandTables *tns_bands;
    int tns_ar;                     // 1: all-pole branch, 0: MA branch
    int sample_index;
};

struct TnsParams {
    const float *spectra;
    float *scratch;
    uint32_t *ranges;               // out: per channel-frame interval of scratch that was written
    const aacfb_frame_info *info;
    const uint8_t *blob;
    const uint32_t *offsets;
    size_t blob_bytes;
    size_t n_cf;
    int sample_index;
    int ar;                         // 1: all-pole branch (decode=true), 0: MA branch
    const TnsBandTables *bands;
    const unsigned *gate;           // != nullptr: run only if *gate != 0 (items the fused kernel left to the generic pass)
};

struct StereoParams {            // stereo tools as a pre-pass (only when TNS has to run between them and the IMDCT)
    const float *spectra;
    float *out;                     // same layout: rows with the ops applied (others copied)
    const aacfb_frame_info *info;
    const aacfb_stereo_ops *stereo;
    size_t n_pairs_frames;          // S * T * nc / 2
};
struct DequantParams {           // inverse quantisation as a pre-pass (TNS modes only)
    const uint8_t *qframes;         // [n_cf] aacfb_qframe
    const aacfb_frame_info *info;
    const DequantTables *dq;
    float *out;                     // [n_cf][1024]
    size_t n_cf;
};
cudaError_t launch_dequant(const DequantParams &P, cudaStream_t stream);
cudaError_t launch_stereo(const StereoParams &P, cudaStream_t stream);
cudaError_t launch_synth(const SynthParams &P, int num_sms, bool generic, cudaStream_t stream);
cudaError_t launch_tns(const TnsParams &P, cudaStream_t stream);
// TNS + synthesis in one kernel (float rows, two channels, items without EIGHT_SHORT frames; the others are
// counted in P.short_items and left to the pre-pass + generic instantiation).
constexpr int kFusedClients = kWorkers - 1;   // synthesis workers per CTA; the last worker filters
constexpr int kFusedRingRows = 128;           // rows per ring slot (two per lane of the filtering worker)
cudaError_t launch_synth_tns(const SynthParams &P, int num_sms, cudaStream_t stream);
size_t tns_ring_floats(int num_sms);
int synth_smem_bytes();

}  // namespace aacfb
