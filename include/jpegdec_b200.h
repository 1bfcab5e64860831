/*
 * jpegdec_b200.h -- batch / device-resident extension of the JPEGDEC C API and the
 * thin host->CUDA FFI underneath it.  Plain C ABI: pointers and sizes only.
 *
 * Why it exists: the reference's hot path -- DecodeJPEG (src/jpeg.inl:4946-5357)
 * driving JPEGDecodeMCU (:2090-2274), JPEGIDCT (:2278-2798), JPEGPutMCU* (:2799-4868)
 * and JPEGDither (:4871-4940) -- decodes one image per call on one core.  A B200 needs
 * ~10^5 independent restart segments in flight, so the throughput path takes a *batch*
 * of images per call.  JPEG_decode() (include/JPEGDEC.h) is this API with n = 1 plus
 * the reference's callback / framebuffer semantics replayed on the host.
 *
 * Per stage, which reference function it replaces:
 *   jdk_prescan        <- JPEGFilter marker handling (:1431-1540) + restart bookkeeping (:5337-5348)
 *   jdk_entropy        <- JPEGDecodeMCU (:2090-2274) incl. the 64-bit window behaviour; for progressive files the DC part
 *                         of JPEGDecodeMCU_P (:1819-1884); parse-only / low-frequency-only variants for 1/8 and 1/4 scale
 *   jdk_stitch/_patch  <- cross-segment window phase of the same function
 *   jdk_unstuff, jdk_chunk_* <- the same for scans without restart markers (chunk-parallel, self-synchronising)
 *   jdk_idct_tb, jdk_idct_color <- JPEGIDCT + DC-only shortcut (:5146-5154) + JPEGPutMCU22/11/12/21/Gray/8BitGray
 *   jdk_scaled         <- the 1/4 and 1/8 paths of the above (:2305-2326, :3323-3396, :3627-3748)
 *   jdk_dither         <- JPEGDither (:4871-4940)
 */
#ifndef JPEGDEC_B200_H
#define JPEGDEC_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct JPEGB200_CTX JPEGB200_CTX;     /* one per (process, GPU) */
typedef struct JPEGB200_BATCH JPEGB200_BATCH; /* one decode job: n images, one pixel type / option set */

/* batch flags */
#define JPEGB200_OUT_DEVICE 1   /* output pointers are device pointers (pixels stay in HBM) */

/* stage indices for JPEGB200_batchGetTimings (milliseconds, CUDA events on the batch stream) */
enum {
    JPEGB200_T_H2D = 0,
    JPEGB200_T_PRESCAN,
    JPEGB200_T_ENTROPY,
    JPEGB200_T_STITCH,
    JPEGB200_T_IDCT,
    JPEGB200_T_DITHER,
    JPEGB200_T_D2H,
    JPEGB200_T_TOTAL,
    JPEGB200_NUM_TIMINGS
};

/* counters for JPEGB200_batchGetCounters */
enum {
    JPEGB200_C_LAUNCHES = 0,   /* kernels launched by the last batchDecode */
    JPEGB200_C_SEGMENTS,
    JPEGB200_C_BLOCKS,
    JPEGB200_C_EVENTS,         /* coefficients rewritten because the reference reads them through a truncated bit window */
    JPEGB200_C_COMPRESSED_BYTES,
    JPEGB200_C_OUTPUT_BYTES,
    JPEGB200_C_RECORD_BYTES,   /* coefficient-record bytes written by the entropy kernel (restart-segment path) */
    JPEGB200_C_H2D_BYTES,
    JPEGB200_C_D2H_BYTES,
    JPEGB200_C_EVENT_CANDIDATES, /* reads that are truncated for at least one possible window phase (examined, not all applied) */
    JPEGB200_NUM_COUNTERS
};

/* ---- context ---- */
JPEGB200_CTX *JPEGB200_create(int device, int arith_mode);
void JPEGB200_destroy(JPEGB200_CTX *ctx);
const char *JPEGB200_lastErrorString(JPEGB200_CTX *ctx);
int JPEGB200_deviceCount(void);
void *JPEGB200_hostAlloc(size_t bytes);   /* pinned host memory for inputs/outputs */
void JPEGB200_hostFree(void *p);
/* Host placement on multi-socket boxes (optional, never needed for correctness).  numaNode: the host NUMA node the
 * context's GPU hangs off (-1 unknown).  bindHostToDevice: pins the CALLING thread to that node's CPUs (within the CPUs the
 * process may use) and prefers the node for its allocations, so that pinned buffers allocated afterwards and the thread
 * that drives the copies sit next to the GPU; returns the number of CPUs in the new mask, 0 = nothing changed.
 * The reference has no counterpart: its decoder runs where the caller's thread runs (src/JPEGDEC.cpp:157-224). */
int JPEGB200_numaNode(JPEGB200_CTX *ctx);
int JPEGB200_bindHostToDevice(JPEGB200_CTX *ctx);
/* device memory for JPEGB200_OUT_DEVICE outputs (plain cudaMalloc / cudaFree / synchronous cudaMemcpy on the context's GPU) */
void *JPEGB200_deviceAlloc(JPEGB200_CTX *ctx, size_t bytes);
void JPEGB200_deviceFree(JPEGB200_CTX *ctx, void *p);
int JPEGB200_deviceRead(JPEGB200_CTX *ctx, void *host_dst, const void *dev_src, size_t bytes);
/* 64-bit digests of n device byte ranges (starts 8-byte aligned), computed on the GPU: digest = sum over the 8-byte
 * little-endian words w[i] (tail zero padded) of mix64(w[i] ^ i * 0x9E3779B97F4A7C15) mod 2^64, mix64 = splitmix64's
 * finaliser.  Lets a caller check device-resident pixels against digests of reference output without a D2H of the pixels. */
int JPEGB200_digestDevice(JPEGB200_CTX *ctx, const void *const *dev_ptrs, const int64_t *lengths, int n, uint64_t *digests);

/* ---- batch job ---- */
/* Parses the n headers on the host (no GPU work).  datas[i]/sizes[i]: JPEG files in host memory
 * (pinned memory makes the upload a straight DMA; files that sit back to back are uploaded with one copy).
 * pixel_type / options as in JPEGDEC.h.  At most 3 GiB of compressed bytes per batchCreate (JPEGB200_decodeBatch
 * takes any amount and splits it); a single file may be at most 512 MiB.
 * Progressive files are accepted when options has JPEG_SCALE_EIGHTH: like the reference (src/jpeg.inl:4964-4966,
 * JPEGDecodeMCU_P :1819-1884) only the DC coefficients of the first scan are decoded; otherwise that image's status
 * is JPEG_UNSUPPORTED_FEATURE. */
JPEGB200_BATCH *JPEGB200_batchCreate(JPEGB200_CTX *ctx, const uint8_t *const *datas, const int32_t *sizes,
                                     int n, int pixel_type, int options);
void JPEGB200_batchDestroy(JPEGB200_BATCH *b);
int JPEGB200_batchCount(JPEGB200_BATCH *b);
/* per-image facts after batchCreate: status is JPEG_SUCCESS or the open() error the reference would give */
int JPEGB200_batchImageInfo(JPEGB200_BATCH *b, int i, int32_t *width, int32_t *height, int32_t *subsample,
                            int32_t *out_w, int32_t *out_h, int32_t *status);
/* bytes needed for a tightly packed output of image i (out_h rows of out_pitch bytes) */
int64_t JPEGB200_batchOutputBytes(JPEGB200_BATCH *b, int i, int64_t *pitch_bytes);
/* Destination for image i.  Device pointer if the batch is decoded with JPEGB200_OUT_DEVICE, else host
 * (pinned recommended).  pitch_bytes = 0 -> tight. */
int JPEGB200_batchSetOutput(JPEGB200_BATCH *b, int i, void *out, int64_t pitch_bytes);
/* Let the library own a device output arena (tight images back to back, 256-B aligned). */
int JPEGB200_batchAllocDeviceOutput(JPEGB200_BATCH *b);
int JPEGB200_batchGetDeviceOutput(JPEGB200_BATCH *b, int i, void **devptr, int64_t *pitch_bytes);
int JPEGB200_batchReadOutput(JPEGB200_BATCH *b, int i, void *host_dst); /* synchronous D2H of one image from the arena */
int JPEGB200_batchErrMcu(JPEGB200_BATCH *b, int i);                      /* first undecodable MCU of image i, -1 if none */
/* dither needs no extra buffers from the caller: packed rows are written to the output. */

int JPEGB200_batchUpload(JPEGB200_BATCH *b);            /* H2D: compressed bytes + descriptors (async) */
int JPEGB200_batchDecode(JPEGB200_BATCH *b, int flags); /* kernel launches (async) */
int JPEGB200_batchDownload(JPEGB200_BATCH *b);          /* D2H of pixels for host outputs (async) */
int JPEGB200_batchWait(JPEGB200_BATCH *b, int32_t *status /* n entries, may be NULL */);
int JPEGB200_batchGetTimings(JPEGB200_BATCH *b, float *ms /* JPEGB200_NUM_TIMINGS */);
int JPEGB200_batchGetCounters(JPEGB200_BATCH *b, int64_t *counters /* JPEGB200_NUM_COUNTERS */);
void *JPEGB200_batchStream(JPEGB200_BATCH *b);          /* cudaStream_t the job runs on */

/* One call for a whole batch of ANY size: create + upload + decode + (download) + wait + destroy.
 * outs[i]: destination (host, or the caller's device memory with JPEGB200_OUT_DEVICE); pitches may be NULL (tight).
 * The batch is run as a pipeline of jobs on separate streams: with host outputs the pixels of one job cross PCIe while the
 * next job's kernels run; with device outputs a job holds at most 192 MiB of compressed bytes, which bounds the transient
 * device memory however large the batch is (the reference equivalent is a loop of JPEG_openRAM + JPEG_decode,
 * src/JPEGDEC.cpp:157-224, which streams any input through a 2 KB window, src/jpeg.inl:1544-1566).
 * Returns 1 = all images decoded, 2 = some images failed (see status[]), 0 = call failed. */
int JPEGB200_decodeBatch(JPEGB200_CTX *ctx, const uint8_t *const *datas, const int32_t *sizes, int n,
                         int pixel_type, int options, void *const *outs, const int64_t *pitches,
                         int flags, int32_t *status);
/* JPEGB200_NUM_COUNTERS counters summed over the jobs of the last JPEGB200_decodeBatch on this context */
int JPEGB200_lastCallCounters(JPEGB200_CTX *ctx, int64_t *counters);
/* CUDA-event stage times (JPEGB200_NUM_TIMINGS, ms) summed over those jobs, and how many jobs there were.  Jobs overlap
 * on the GPU unless the pipeline depth is 1, so the sums are an upper bound of the stage's share of the call. */
int JPEGB200_lastCallTimings(JPEGB200_CTX *ctx, float *ms, int *jobs);
/* jobs JPEGB200_decodeBatch keeps in flight (0 = default: 6 with host outputs, 3 with device outputs; 1 = strictly serial) */
int JPEGB200_setPipelineDepth(JPEGB200_CTX *ctx, int jobs_in_flight);

/* ---- shared-table blob (multi-GPU: rank 0 exports, NCCL broadcast, other ranks import) ---- */
#define JPEGB200_TABLE_BLOB_BYTES (10496 * 2 + 3 * 64 * 2 + 16)
int JPEGB200_exportTables(const uint8_t *jpeg, int size, uint8_t *blob /* JPEGB200_TABLE_BLOB_BYTES */);
int JPEGB200_setSharedTables(JPEGB200_CTX *ctx, const uint8_t *blob);
int JPEGB200_sharedTableHits(JPEGB200_CTX *ctx);

#ifdef __cplusplus
}
#endif
#endif /* JPEGDEC_B200_H */
