/*
 * jd_kernels.cuh -- hand-written sm_100a kernels of the decode pipeline.
 *
 *   jdk_prescan       one CTA per image: finds RSTn markers, writes per-segment byte offsets
 *   jdk_unstuff_segs  one warp per restart segment: FF00 -> FF into 16-byte aligned clean streams
 *   jdk_entropy       one thread per restart segment: Huffman walk (jd_core.h jd_decode_segment), tables + a stream ring
 *                     + a record staging chunk per walker in shared memory; compact coefficient records + block headers
 *   jdk_stitch        one thread per image: resolves the reference's bit-window phase across segments and folds the
 *                     per-segment status into the image status
 *   jdk_patch         one thread per truncation event: rewrites the affected record
 *   jdk_unstuff<count/write>, jdk_chunk_parse / _prefix / _emit / _stitch
 *                     scans without restart markers: chunk-parallel entropy decode (jd_chunk.h)
 *   jdk_idct_tb       fused record-expand + dequant + 8x8 integer IDCT + colour conversion for 4:2:0 colour at full size:
 *                     blocks binned by class, one thread per block for the common classes, planes staged in shared
 *                     memory, 128-bit coalesced scanline stores
 *   jdk_idct_p        the same for every other sampling / pixel type / half scale (SSE2-build arithmetic): one thread per
 *                     block, two columns per register
 *   jdk_idct_color    those cases in scalar-build arithmetic: 8 lanes per block
 *   jdk_scaled        1/4 and 1/8 decode (DC / 2x2 butterfly), one thread per MCU
 *   jdk_dither        (jd_device.cu) Floyd-Steinberg 1/2/4-bpp, one warp-lane per image row wavefront
 *   jdk_digest        (jd_device.cu) 64-bit digest of device-resident pixels (verification aid)
 *
 * No tensor cores: this is integer, byte-granular, HBM-bound work.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "jd_core.h"
#include "jd_chunk.h"
#include "jd_internal.h"

#define JD_NONE 0xFFFFFFFFu
#ifndef JD_ENTROPY_THREADS
#define JD_ENTROPY_THREADS 128
#endif
#define JD_RING_STRIDE 36   /* words between two walkers' rings: 32 + 4 keeps 16-byte alignment and spreads the banks */

__constant__ uint8_t c_tpos[64] = JD_TPOS_INIT;

/* ------------------------------------------------------------------------------------ */
/* prescan                                                                                */
/* ------------------------------------------------------------------------------------ */
/* 0x80 in every byte of the result where the byte of x is zero (exact per byte) */
__device__ __forceinline__ uint32_t jd_zero_bytes(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu); }

/* One CTA per image walks its scan 16 KB at a time: 64 contiguous bytes per thread (four 16-byte loads in flight), RSTn
 * markers (FF D0..D7) found four bytes at a time with byte-flag words, then a block-wide scan of the counts gives every
 * marker its index.  (4 KB per iteration with byte compares: 0.23 ms per 1024 HD files, 0.71 ms per 512 UHD files, all of it
 * the latency of ~70 / ~270 dependent iterations.) */
__global__ void __launch_bounds__(256) jdk_prescan(const uint8_t *__restrict__ data, const JDImageDesc *__restrict__ imgs,
                                                    uint32_t *__restrict__ seg_start)
{
    const JDImageDesc &im = imgs[blockIdx.x];
    const uint32_t nseg = im.nseg, base = im.seg_base;
    const uint32_t tid = threadIdx.x;
    __shared__ uint32_t s_wtot[2][8];
    if (nseg == 0) return; /* header rejected on the host: owns no segment slots */
    if (tid == 0) seg_start[base] = im.scan_off;
    for (uint32_t i = 1 + tid; i < nseg; i += 256) seg_start[base + i] = JD_NONE;
    if (nseg <= 1) return;
    __syncthreads();
    const uint32_t lo = im.scan_off, hi = im.scan_end;
    uint32_t found = 0; /* markers found so far (block-uniform) */
    int buf = 0;
    for (uint32_t p0 = lo & ~15u; p0 < hi && found + 1 < nseg; p0 += 256 * 64) {
        const uint32_t p = p0 + tid * 64;
        unsigned long long m = 0;
        if (p < hi) {
            uint32_t w[17];
            const uint4 zero4 = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint4 v = (p + 16u * j < hi) ? *reinterpret_cast<const uint4 *>(data + p + 16 * j) : zero4;
                w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
            }
            w[16] = (p + 64 < hi) ? (uint32_t)data[p + 64] : 0u;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t nx = __funnelshift_r(w[i], w[i + 1], 8);                       /* the four bytes one further */
                const uint32_t f = jd_zero_bytes(~w[i]) & jd_zero_bytes((nx & 0xF8F8F8F8u) ^ 0xD0D0D0D0u);
                m |= (unsigned long long)((((f >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * i);
            }
            /* only markers whose two bytes lie inside [lo, hi) */
            if (p < lo) m &= ~0ull << (lo - p);
            if (p + 64 >= hi) { const uint32_t nvalid = hi - 1u - p; m &= (nvalid >= 64u) ? ~0ull : ((1ull << nvalid) - 1ull); }
        }
        if (!__syncthreads_or(m != 0ull)) continue;
        /* block-wide exclusive scan of popc(m) */
        const uint32_t cnt = (uint32_t)__popcll(m);
        uint32_t x = cnt;
        const uint32_t lane = tid & 31, wid = tid >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += y; }
        if (lane == 31) s_wtot[buf][wid] = x;
        __syncthreads();
        uint32_t wbase = 0, tot = 0;
#pragma unroll
        for (int w2 = 0; w2 < 8; w2++) { uint32_t t = s_wtot[buf][w2]; if ((uint32_t)w2 < wid) wbase += t; tot += t; }
        uint32_t k = found + wbase + x - cnt; /* index of this thread's first marker */
        while (m) {
            const int bit = __ffsll((long long)m) - 1;
            m &= m - 1;
            if (k + 1 < nseg) seg_start[base + k + 1] = p + bit + 2;
            k++;
        }
        found += tot;
        buf ^= 1;
    }
}

/* ------------------------------------------------------------------------------------ */
/* entropy decode                                                                         */
/* ------------------------------------------------------------------------------------ */
struct JDEventSinkDev {
    JDEvent *events;
    uint32_t *count;
    uint32_t cap;
    __device__ __forceinline__ void push(const JDEvent &e)
    {
        uint32_t i = atomicAdd(count, 1u);
        if (i < cap) events[i] = e;
    }
};

struct JDEntropyArgs {
    const uint8_t *data;          /* compressed batch buffer (4-byte aligned base) */
    const JDImageDesc *imgs;
    const uint16_t *luts;         /* lut sets, JD_LUT_ENTRIES each */
    const uint32_t *work;         /* padded work list: global segment index or JD_NONE */
    const uint32_t *cta_lut;      /* LUT set per CTA */
    const uint32_t *seg_img;      /* image of each segment */
    const uint32_t *seg_start;    /* from prescan */
    jd_u64 *blk_hdr;
    uint16_t *rec;
    uint32_t *seg_jmap;
    uint32_t *seg_status;         /* 0 ok, else (code << 28) | local err mcu */
    uint32_t *seg_nrec;
    JDEvent *events;
    uint32_t *event_count;
    uint32_t event_cap;
    uint32_t nwork;
    uint32_t dc_output;           /* 1: 1/8-scale job, only DC values are consumed downstream; 2: 1/4-scale job, zigzag 1..4 */
    const uint8_t *clean;         /* un-stuffed segments (jdk_unstuff_segs) and their lengths, when that stage ran */
    const uint32_t *seg_clen;
};

/* Un-stuffed copy of a restart segment (JD_ENTROPY_CLEAN pipeline): segment `seg` whose raw bytes start at `start` is
 * written at this 16-byte aligned offset of the clean buffer; consecutive segments are at least (raw length + 19) apart,
 * which covers the un-stuffed bytes rounded down to 16 plus one zero-padded 16-byte chunk. */
__host__ __device__ __forceinline__ uint32_t jd_clean_off(uint32_t start, uint32_t seg) { return (start & ~15u) + 32u * seg; }

template <bool CLEAN>
__device__ __forceinline__ void jd_entropy_body(const JDEntropyArgs &a, const uint16_t *s_lut, const uint32_t *s_tpos, uint32_t *s_ring, uint16_t *s_stage)
{
    const uint32_t wi = blockIdx.x * JD_ENTROPY_THREADS + threadIdx.x;
    if (wi >= a.nwork) return;
    const uint32_t seg = a.work[wi];
    if (seg == JD_NONE) return;
    const uint32_t img = a.seg_img[seg];
    const JDImageDesc &im = a.imgs[img];
    const uint32_t sl = seg - im.seg_base; /* local segment index */
    const uint32_t total_mcus = (uint32_t)im.mcus_x * im.mcus_y;
    const uint32_t m0 = sl * im.mcus_per_seg;
    JDSegIn in;
    in.data = a.data;
    in.start = a.seg_start[seg];
    in.end = im.scan_end;
    in.nmcu = (m0 + im.mcus_per_seg <= total_mcus) ? im.mcus_per_seg : total_mcus - m0;
    in.bpm = im.bpm;
    in.ncomp = im.ncomp;
    in.tsel = im.tsel;
    in.seg = seg;
    in.img = img;
    in.ring = s_ring + threadIdx.x * JD_RING_STRIDE;
    in.stage = s_stage + threadIdx.x * 8;
    jd_segin_whole_interval(&in);
    in.blk0 = im.blk_base + m0 * im.bpm;
    jd_u64 *hdr = a.blk_hdr + im.blk_base + (size_t)m0 * im.bpm;
    JDSegOut so;
    if (in.start == JD_NONE || in.start < im.scan_off || in.start > im.scan_end) {
        /* restart marker missing: everything from here on is undecodable */
        for (uint32_t b = 0; b < in.nmcu * in.bpm; b++) hdr[b] = 0ull;
        a.seg_jmap[seg] = JD_JW_INIT;
        a.seg_status[seg] = ((uint32_t)JD_SEG_MISSING << 28);
        a.seg_nrec[seg] = 0;
        return;
    }
    /* record area of this segment (jd_core.h JD_REC_INDEX): no prefix sum over segments is needed */
    const uint32_t next = (sl + 1 < im.nseg) ? a.seg_start[seg + 1] : JD_NONE;
    const uint32_t seg_end = (next != JD_NONE) ? next : im.scan_end;
    in.rec_index0 = JD_REC_INDEX(in.start - im.comp_off, sl);
    in.rec_cap = JD_REC_CAP(seg_end > in.start ? seg_end - in.start : 0u);
    uint16_t *rec = a.rec + im.rec_base + in.rec_index0;
    if (CLEAN) {
        in.data = a.clean;
        in.start = jd_clean_off(in.start, seg);     /* the host keeps the clean buffer below 4 GiB */
        in.end = in.start + a.seg_clen[seg];
    }
    JDEventSinkDev sink{a.events, a.event_count, a.event_cap};
    in.al = (im.prog >> 8) & 15u;
    if (im.prog & 1u) jd_decode_segment<JDEventSinkDev, JD_MODE_DC_SCAN, CLEAN>(in, s_lut, s_tpos, hdr, rec, sink, so);
    else if (a.dc_output == 1u) jd_decode_segment<JDEventSinkDev, JD_MODE_PARSE_AC, CLEAN>(in, s_lut, s_tpos, hdr, rec, sink, so);
    else if (a.dc_output == 2u) jd_decode_segment<JDEventSinkDev, JD_MODE_STORE_LOW, CLEAN>(in, s_lut, s_tpos, hdr, rec, sink, so);
    else jd_decode_segment<JDEventSinkDev, JD_MODE_BASELINE, CLEAN>(in, s_lut, s_tpos, hdr, rec, sink, so);
    a.seg_jmap[seg] = so.jmap;
    a.seg_status[seg] = (so.err_mcu < 0) ? 0u : (((uint32_t)so.status << 28) | ((uint32_t)so.err_mcu & 0x0FFFFFFFu));
    a.seg_nrec[seg] = so.nrec;
}

template <bool CLEAN>
__global__ void __launch_bounds__(JD_ENTROPY_THREADS) jdk_entropy(const JDEntropyArgs a)
{
    __shared__ __align__(16) uint16_t s_lut[JD_LUT_ENTRIES];
    __shared__ uint32_t s_tpos[64];
    __shared__ __align__(16) uint32_t s_ring[CLEAN ? JD_ENTROPY_THREADS * JD_RING_STRIDE : 4];   /* per-walker stream rings (jd_core.h) */
    __shared__ __align__(16) uint16_t s_stage[JD_ENTROPY_THREADS * 8];                           /* per-walker record staging chunks */
    for (int i = threadIdx.x; i < 64; i += JD_ENTROPY_THREADS) s_tpos[i] = jd_tposw(c_tpos[i]);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.luts + (size_t)a.cta_lut[blockIdx.x] * JD_LUT_ENTRIES);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (int i = threadIdx.x; i < JD_LUT_ENTRIES * 2 / 16; i += JD_ENTROPY_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    jd_entropy_body<CLEAN>(a, s_lut, s_tpos, s_ring, s_stage);
}

/* ------------------------------------------------------------------------------------ */
/* un-stuff the restart segments (JPEGFilter, src/jpeg.inl:1431-1540: FF00 -> FF, the stream    */
/* of a segment ends at the first FFxx marker) so that the entropy kernel's bit reader is a     */
/* plain word stream.  One warp per segment: 512 raw bytes per iteration (16 per lane, aligned  */
/* 16-byte loads), kept bytes compacted through a shared-memory staging line and written out as */
/* aligned 16-byte stores.                                                                      */
/* ------------------------------------------------------------------------------------ */
#define JD_UNSTUFF_WARPS 4
__global__ void __launch_bounds__(JD_UNSTUFF_WARPS * 32) jdk_unstuff_segs(const uint8_t *__restrict__ data, const JDImageDesc *__restrict__ imgs,
                                                                         const uint32_t *__restrict__ seg_img, const uint32_t *__restrict__ seg_start,
                                                                         uint32_t nseg, uint8_t *__restrict__ clean, uint32_t *__restrict__ seg_clen)
{
    __shared__ __align__(16) uint8_t s_stage[JD_UNSTUFF_WARPS][512 + 32];
    const uint32_t seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (seg >= nseg) return;
    const JDImageDesc &im = imgs[seg_img[seg]];
    if (im.nch != 0u || im.nseg == 0u) return;           /* restart-free scans take the chunk path; rejected headers own nothing */
    const uint32_t sl = seg - im.seg_base;
    const uint32_t start = seg_start[seg];
    if (start == JD_NONE || start < im.scan_off || start > im.scan_end) { if (lane == 0) seg_clen[seg] = 0; return; }
    const uint32_t next = (sl + 1 < im.nseg) ? seg_start[seg + 1] : JD_NONE;
    /* the RSTn marker that ends the segment sits in the two bytes before the next segment's start */
    const uint32_t end = (next != JD_NONE && next >= start + 2u) ? next - 2u : im.scan_end;
    uint8_t *stage = s_stage[threadIdx.x >> 5];
    uint8_t *dst = clean + jd_clean_off(start, seg);
    uint32_t fill = 0, outpos = 0;
    uint32_t carry_ff = 0;                               /* the byte before this iteration's first byte was a kept 0xFF */
    bool done = false;
    for (uint32_t p0 = start & ~15u; p0 < end && !done; p0 += 512u) {
        const uint32_t p = p0 + lane * 16u;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (p < end) v = *reinterpret_cast<const uint4 *>(data + p);    /* the batch buffer is padded past its last file */
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        /* byte after my 16 (the next lane's first; lane 31 reads it) and byte before (the previous lane's last) */
        uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1) & 0xFFu;
        if (lane == 31) nxt = (p + 16u < end) ? (uint32_t)data[p + 16u] : 0xD9u;
        uint32_t prv = __shfl_up_sync(0xffffffffu, v.w, 1) >> 24;
        if (lane == 0) prv = carry_ff ? 0xFFu : 0u;
        uint32_t keep = 0, endpos = JD_NONE;
        if (p < end && p + 16u > start) {
            const uint32_t ffm = (((~w[0]) - 0x01010101u) & w[0] & 0x80808080u) | (((~w[1]) - 0x01010101u) & w[1] & 0x80808080u) |
                                 (((~w[2]) - 0x01010101u) & w[2] & 0x80808080u) | (((~w[3]) - 0x01010101u) & w[3] & 0x80808080u);
            if (ffm == 0u && prv != 0xFFu && p >= start && p + 16u <= end) {
                keep = 0xFFFFu;                          /* no 0xFF in sight: every byte is data (has-zero-byte test on ~w is exact for "any") */
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const uint32_t pos = p + (uint32_t)i;
                    const uint32_t b0 = (w[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
                    const uint32_t bn = (i < 15) ? ((w[(i + 1) >> 2] >> (((i + 1) & 3) * 8)) & 0xFFu) : nxt;
                    const uint32_t bp = (i > 0) ? ((w[(i - 1) >> 2] >> (((i - 1) & 3) * 8)) & 0xFFu) : prv;
                    if (pos < start || pos >= end) continue;
                    const bool after_ff = (bp == 0xFFu) && (pos > start);
                    if (b0 == 0xFFu && !after_ff) {
                        /* FF00 keeps the FF; FF + anything else (or FF as the last byte) ends the data */
                        if (pos + 1u < end && bn == 0u) keep |= 1u << i; else if (endpos == JD_NONE) endpos = pos;
                    } else if (!(after_ff && b0 == 0u)) {
                        /* an FF directly after a kept FF00 pair's zero is handled above (after_ff is false for it:
                         * the byte before it is 00); a byte after an FF that is not 00 never gets here (stream ended) */
                        keep |= 1u << i;
                    }
                }
            }
        }
        const uint32_t stop = __reduce_min_sync(0xffffffffu, endpos);
        if (stop != JD_NONE) {
            done = true;
#pragma unroll
            for (int i = 0; i < 16; i++) if (p + (uint32_t)i >= stop) keep &= ~(1u << i);
        }
        /* does the next iteration start right after a kept 0xFF?  (only lane 31's last byte matters) */
        carry_ff = __shfl_sync(0xffffffffu, ((keep >> 15) & 1u) & (uint32_t)((v.w >> 24) == 0xFFu), 31);
        const uint32_t cnt = __popc(keep);
        uint32_t x = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += y; }
        uint32_t o = fill + x - cnt;
#pragma unroll
        for (int i = 0; i < 16; i++) if (keep & (1u << i)) stage[o++] = (uint8_t)(w[i >> 2] >> ((i & 3) * 8));
        const uint32_t total = fill + __shfl_sync(0xffffffffu, x, 31);
        __syncwarp();
        const uint32_t nflush = total >> 4;
        uint4 chunk = make_uint4(0, 0, 0, 0);
        if (lane < nflush) chunk = *reinterpret_cast<const uint4 *>(stage + 16u * lane);
        uint32_t rem_b = 0;
        const uint32_t rem = total & 15u;
        if (lane < rem) rem_b = stage[16u * nflush + lane];
        __syncwarp();
        if (lane < nflush) *reinterpret_cast<uint4 *>(dst + outpos + 16u * lane) = chunk;
        if (lane < rem) stage[lane] = (uint8_t)rem_b;
        __syncwarp();
        outpos += 16u * nflush;
        fill = rem;
    }
    /* tail: the last partial chunk, zero padded (the reader's last word must end in zeros) */
    if (lane >= fill && lane < 16u) stage[lane] = 0;
    __syncwarp();
    if (lane == 0u) *reinterpret_cast<uint4 *>(dst + outpos) = *reinterpret_cast<const uint4 *>(stage);
    if (lane == 0) seg_clen[seg] = outpos + fill;
}

/* ------------------------------------------------------------------------------------ */
/* stitch + patch                                                                          */
/* ------------------------------------------------------------------------------------ */
__global__ void jdk_stitch(JDImageDesc *imgs, uint32_t nimg, const uint32_t *__restrict__ seg_jmap,
                           const uint32_t *__restrict__ seg_status, uint32_t *__restrict__ seg_phase,
                           const uint32_t *__restrict__ seg_nrec, unsigned long long *__restrict__ rec_count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nimg) return;
    JDImageDesc &im = imgs[i];
    uint32_t c = 0, status = 0, err_mcu = 0;
    unsigned long long nrec = 0;
    for (uint32_t s = 0; s < im.nseg; s++) {
        if (im.nch == 0u) nrec += seg_nrec[im.seg_base + s];
        const uint32_t g = im.seg_base + s;
        seg_phase[g] = c;
        const uint32_t j = (seg_jmap[g] >> (4 * c)) & 15u;
        c = (j >= 6u) ? 0u : j;
        const uint32_t st = seg_status[g];
        if (st != 0u && status == 0u) { status = st >> 28; err_mcu = s * im.mcus_per_seg + (st & 0x0FFFFFFFu); }
    }
    im.status = status;
    im.err_mcu = err_mcu;
    if (nrec) atomicAdd(rec_count, nrec);
}

__global__ void jdk_patch(const JDImageDesc *__restrict__ imgs, const JDEvent *__restrict__ events, const uint32_t *__restrict__ event_count, uint32_t cap,
                          const uint32_t *__restrict__ seg_phase, const jd_u64 *__restrict__ blk_hdr, uint16_t *__restrict__ rec,
                          uint32_t *__restrict__ applied)
{
    uint32_t n = *event_count;
    if (n > cap) n = cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const JDEvent e = events[i];
        const uint32_t jc = (e.j1 >> (4 * seg_phase[e.seg])) & 15u;
        if (8 * (int)jc + e.p7 + e.s > 64) {
            jd_patch_record(rec + imgs[e.img].rec_base, blk_hdr[e.blk], e.ord, jd_event_value(&e, jc));
            atomicAdd(applied, 1u);
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* restart-free scans: un-stuff, then chunk-parallel entropy decode (jd_chunk.h)            */
/* ------------------------------------------------------------------------------------ */
struct JDChunkArgs {
    const uint8_t *comp;           /* raw batch buffer */
    uint8_t *filt;                 /* un-stuffed copy (same offsets) */
    JDImageDesc *imgs;
    const uint16_t *luts;
    const uint32_t *cimg_list;     /* indices of the chunked images */
    uint32_t ncimg;
    uint32_t *flen;                /* per image: un-stuffed scan length */
    uint32_t nchunks;
    uint32_t *X_in, *X_out;        /* exit state of every chunk = entry state of its right neighbour (double buffered across passes) */
    uint32_t *Ep;                  /* entry state each chunk was last parsed from */
    uint32_t max_nch;              /* chunks of the longest scan (grid x = ceil / 128) */
    uint32_t *cn, *cpre, *cjmap, *cstatus, *cnown;
    uint32_t *cfirst;              /* per chunk: first block that starts in it (jd_chunk_parse) | invalid-code flag << 31 */
    int32_t *cdcs, *cpe;           /* per chunk x 3: DC sums (parse pass) / DC predictors at the chunk's first block (prefix) */
    uint32_t *changed;
    jd_u64 *blk_hdr;
    uint16_t *rec;
    JDEvent *events;
    uint32_t *event_count;
    uint32_t event_cap;
    uint32_t *seg_phase, *seg_jmap, *seg_status;
    uint32_t nseg_total;           /* phase slot of chunk g = nseg_total + g */
};

/* Restart-free scans: FF00 -> FF, stop at the first marker (JPEGFilter, jpeg.inl:1431-1540).  One warp per 4 KB piece of a
 * scan, two passes: count the bytes each piece keeps (and whether a marker ends the scan inside it), then every piece sums
 * the counts to its left and writes its kept bytes there.  (The first version walked a whole scan with one warp: 1.9 ms per
 * 1024 HD images, all of it latency.) */
#define JD_UNSTUFF_PIECE 4096u
/* flags (0x80 per byte) for bytes [lo, hi) of a word, 0 <= lo, hi <= 4 */
__device__ __forceinline__ uint32_t jd_byte_range(uint32_t lo, uint32_t hi)
{
    const uint32_t below_hi = (hi >= 4u) ? 0x80808080u : (0x80808080u & ((1u << (8u * hi)) - 1u));
    const uint32_t below_lo = (lo >= 4u) ? 0x80808080u : (0x80808080u & ((1u << (8u * lo)) - 1u));
    return below_hi & ~below_lo;
}

template <bool WRITE>
__device__ __forceinline__ uint32_t jd_unstuff_piece(const uint8_t *__restrict__ src, uint32_t len, uint32_t p0, uint32_t p1, uint8_t *dst,
                                                     uint32_t base, bool &stopped, uint32_t lane)
{
    /* 128 bytes per iteration, one ALIGNED 32-bit word per lane, classified four bytes at a time with byte-flag words: the walk
     * runs over word addresses, so the first word of a piece may begin up to 3 bytes before p0 (those bytes belong to the piece
     * on the left and are masked off), and pieces end on the same grid.  The byte before / after a word comes from the
     * neighbour lane; across iterations from the previous iteration's lane 31 / the next iteration's word, loaded one
     * iteration ahead. */
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t *wsrc = reinterpret_cast<const uint32_t *>(src - mis);
    /* piece [p0, p1) in scan offsets = word-grid offsets [g0, g1) where grid offset = scan offset + mis */
    const uint32_t g0 = (p0 == 0u) ? 0u : ((p0 + mis) & ~3u), g1 = (p1 >= len) ? (len + mis) : ((p1 + mis) & ~3u);
    const uint32_t vlo = (g0 > mis) ? g0 : mis;                     /* this piece's bytes of the scan: grid offsets [vlo, g1) */
    const uint32_t wlimit = (len + mis + 3u) >> 2;                  /* words that hold scan bytes */
    uint32_t kept = 0;
    stopped = false;
    if (g0 >= g1) return 0u;
    uint32_t carry = (g0 > mis) ? (uint32_t)__ldg(src + (g0 - mis) - 1u) : 0u;    /* byte before the piece */
    uint32_t wn = ((g0 >> 2) + lane < wlimit) ? __ldg(wsrc + (g0 >> 2) + lane) : 0xD9D9D9D9u;
    for (uint32_t q0 = g0; q0 < g1 && !stopped; q0 += 128) {
        const uint32_t gq = q0 + lane * 4;                         /* grid offset of this lane's word */
        const uint32_t w = wn;
        { const uint32_t wi = ((q0 + 128u) >> 2) + lane; wn = (q0 + 128u < g1 + 4u && wi < wlimit) ? __ldg(wsrc + wi) : 0xD9D9D9D9u; }
        uint32_t before = __shfl_up_sync(0xffffffffu, w >> 24, 1), after = __shfl_down_sync(0xffffffffu, w & 0xFFu, 1);
        const uint32_t first_next = __shfl_sync(0xffffffffu, wn & 0xFFu, 0);
        if (lane == 0) before = carry;
        if (lane == 31) after = first_next;
        carry = __shfl_sync(0xffffffffu, w >> 24, 31);
        /* byte flags */
        const uint32_t ff = jd_zero_bytes(~w), zz = jd_zero_bytes(w);
        const uint32_t prev_ff = (ff << 8) | ((before == 0xFFu) ? 0x80u : 0u);
        const uint32_t next_zz = (zz >> 8) | ((after == 0u) ? 0x80000000u : 0u);
        const uint32_t lo = (vlo > gq) ? ((vlo - gq < 4u) ? vlo - gq : 4u) : 0u;
        const uint32_t hi = (g1 > gq) ? ((g1 - gq < 4u) ? g1 - gq : 4u) : 0u;
        const uint32_t mine = jd_byte_range(lo, hi);
        /* the first byte of the scan has no byte before it */
        const uint32_t noprev = (gq <= mis && mis < gq + 4u) ? (0x80u << (8u * (mis - gq))) : 0u;
        uint32_t keep = mine & ~(zz & prev_ff & ~noprev);           /* everything but the zero that follows an FF */
        const uint32_t mk = mine & ff & ~next_zz;                   /* FF followed by something else: a marker ends the scan */
        const uint32_t anymk = __ballot_sync(0xffffffffu, mk != 0u);
        if (anymk) {
            stopped = true;
            const uint32_t fl = (uint32_t)__ffs((int)anymk) - 1u;  /* first lane with a marker */
            if (lane > fl) keep = 0u;
            else if (lane == fl) keep &= ((mk & (0u - mk)) - 1u);   /* bytes below its first marker byte */
        }
        const uint32_t cnt = __popc(keep);
        if (WRITE) {
            uint32_t x = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += y; }
            uint8_t *o = dst + base + kept + x - cnt;
            /* kept bytes to the front, in order */
            uint32_t v = w, m = keep;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (m & 0x80u) *o++ = (uint8_t)v;
                v >>= 8; m >>= 8;
            }
            kept += __shfl_sync(0xffffffffu, x, 31);
        } else {
            kept += cnt;                                            /* per lane; summed after the loop */
        }
    }
    if (!WRITE) kept = __reduce_add_sync(0xffffffffu, kept);
    return kept;
}

/* grid: x = groups of 4 pieces, y = position in cimg_list; per piece scratch = cn / cpre at the piece's first chunk */
template <bool WRITE>
__global__ void __launch_bounds__(128) jdk_unstuff(const JDChunkArgs a)
{
    const uint32_t piece = blockIdx.x * 4u + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    const uint32_t ii = a.cimg_list[blockIdx.y];
    const JDImageDesc &im = a.imgs[ii];
    const uint32_t len = im.scan_end - im.scan_off;
    const uint32_t p0 = piece * JD_UNSTUFF_PIECE;
    if (p0 >= len && !(len == 0u && piece == 0u)) return;
    const uint32_t p1 = (p0 + JD_UNSTUFF_PIECE < len) ? p0 + JD_UNSTUFF_PIECE : len;
    const uint8_t *src = a.comp + im.scan_off;
    uint8_t *dst = a.filt + im.scan_off;
    const uint32_t slot = im.chunk_base + piece * (JD_UNSTUFF_PIECE / JD_CHUNK_BYTES);   /* nch >= len / 512 + 1 */
    bool stopped;
    if (!WRITE) {
        const uint32_t kept = jd_unstuff_piece<false>(src, len, p0, p1, dst, 0u, stopped, lane);
        if (lane == 0) { a.cn[slot] = kept; a.cpre[slot] = stopped ? 1u : 0u; }
        return;
    }
    /* bytes kept to the left, and whether the scan already ended there */
    uint32_t base = 0, ended = 0;
    for (uint32_t q = lane; q < piece; q += 32u) {
        const uint32_t sl = im.chunk_base + q * (JD_UNSTUFF_PIECE / JD_CHUNK_BYTES);
        base += a.cn[sl]; ended |= a.cpre[sl];
    }
    base = __reduce_add_sync(0xffffffffu, base);
    ended = __reduce_or_sync(0xffffffffu, ended);
    if (ended) return;
    const uint32_t kept = jd_unstuff_piece<true>(src, len, p0, p1, dst, base, stopped, lane);
    if (stopped || p1 >= len) {
        /* the scan ends in this piece: its un-stuffed length, and zeros for the reads that run a few bytes past the end */
        if (lane < 24) dst[base + kept + lane] = 0;
        if (lane == 0) a.flen[ii] = base + kept;
    }
}

__device__ __forceinline__ JDScanIn jd_scan_of(const JDChunkArgs &a, const JDImageDesc &im, uint32_t ii)
{
    JDScanIn sc;
    sc.filt = a.filt; sc.f0 = im.scan_off; sc.flen = a.flen[ii];
    sc.bpm = im.bpm; sc.ncomp = im.ncomp; sc.tsel = im.tsel;
    sc.total_blocks = (uint32_t)im.mcus_x * im.mcus_y * im.bpm;
    return sc;
}

/* The chunk kernels run one CTA per 128 consecutive chunks of ONE image (blockIdx.y = position in cimg_list), so the
 * image's Huffman table set sits in shared memory like in jdk_entropy. */
__device__ __forceinline__ void jd_load_lut_set(uint16_t *s_lut, const uint16_t *g_lut)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(g_lut);
    uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
    for (uint32_t i = threadIdx.x; i < (uint32_t)(JD_LUT_ENTRIES / 8); i += blockDim.x) dst[i] = __ldg(src + i);
}

/* One speculative pass.  The entry state of chunk c is the exit state chunk c-1 produced in the previous pass (X_in);
 * a chunk whose entry state is the one it was last parsed from keeps its results, so after the first two passes only the
 * few chunks whose left neighbour had not re-synchronised are parsed again.
 * (Measured and dropped: staging the CTA's 64 KB of stream in shared memory -- 8 resident warps per SM instead of 40, 1.84 ->
 * 3.95 ms -- and a 16-word stream ring per parser topped up at block starts like jdk_entropy's -- 28 warps, 1.44 -> 1.50 ms:
 * with 40 resident warps the loads straight from global memory are hidden well enough.) */
__global__ void __launch_bounds__(128) jdk_chunk_parse(const JDChunkArgs a)
{
    __shared__ __align__(16) uint16_t s_lut[JD_LUT_ENTRIES];
    const uint32_t ii = a.cimg_list[blockIdx.y];
    const JDImageDesc &im = a.imgs[ii];
    const uint32_t cb = blockIdx.x * 128u, c = cb + threadIdx.x;
    if (cb >= im.nch) return;
    const uint32_t g = im.chunk_base + c;
    const bool live = c < im.nch;
    uint32_t entry = 0;
    bool need = false;
    if (live) {
        entry = (c == 0) ? JD_CS_PACK(0, 0, 0) : a.X_in[g - 1];
        need = entry != a.Ep[g];
        if (!need) a.X_out[g] = a.X_in[g];
    }
    if (!__syncthreads_or(need ? 1 : 0)) return;
    jd_load_lut_set(s_lut, a.luts + (size_t)im.lutset * JD_LUT_ENTRIES);
    __syncthreads();
    if (!need) return;
    const JDScanIn sc = jd_scan_of(a, im, ii);
    uint32_t nstart, bad, first;
    int32_t dcs[3];
    const uint32_t ex = jd_chunk_parse(sc, s_lut, c, entry, &nstart, &bad, dcs, &first);
    a.cn[g] = nstart;
    a.cfirst[g] = first | (bad << 31);
    a.cdcs[3 * g] = dcs[0]; a.cdcs[3 * g + 1] = dcs[1]; a.cdcs[3 * g + 2] = dcs[2];
    a.Ep[g] = entry;
    if (ex != a.X_in[g]) atomicOr(a.changed, 1u);
    a.X_out[g] = ex;
}

/* per restart-free scan (one warp): blocks started before each chunk and the DC predictors at each chunk's first block */
__global__ void __launch_bounds__(128) jdk_chunk_prefix(const JDChunkArgs a)
{
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (w >= a.ncimg) return;
    const JDImageDesc &im = a.imgs[a.cimg_list[w]];
    uint32_t run = 0;
    int r0 = 0, r1 = 0, r2 = 0;
    for (uint32_t cb = 0; cb < im.nch; cb += 32u) {
        const uint32_t c = cb + lane, g = im.chunk_base + c;
        const bool live = c < im.nch;
        const uint32_t n = live ? a.cn[g] : 0u;
        const int d0 = live ? a.cdcs[3 * g] : 0, d1 = live ? a.cdcs[3 * g + 1] : 0, d2 = live ? a.cdcs[3 * g + 2] : 0;
        uint32_t x = n;
        int y0 = d0, y1 = d1, y2 = d2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t xv = __shfl_up_sync(0xffffffffu, x, d);
            const int v0 = __shfl_up_sync(0xffffffffu, y0, d), v1 = __shfl_up_sync(0xffffffffu, y1, d), v2 = __shfl_up_sync(0xffffffffu, y2, d);
            if (lane >= (uint32_t)d) { x += xv; y0 += v0; y1 += v1; y2 += v2; }
        }
        if (live) {
            a.cpre[g] = run + x - n;
            a.cpe[3 * g] = r0 + y0 - d0; a.cpe[3 * g + 1] = r1 + y1 - d1; a.cpe[3 * g + 2] = r2 + y2 - d2;
        }
        run += __shfl_sync(0xffffffffu, x, 31);
        r0 += __shfl_sync(0xffffffffu, y0, 31); r1 += __shfl_sync(0xffffffffu, y1, 31); r2 += __shfl_sync(0xffffffffu, y2, 31);
    }
}

/* every chunk decodes the blocks that start in it: the entropy walk of jdk_entropy (jd_decode_segment, CLEAN reader: stream
 * ring and record staging in shared memory) started in the middle of the stream */
__global__ void __launch_bounds__(128) jdk_chunk_emit(const JDChunkArgs a)
{
    __shared__ __align__(16) uint16_t s_lut[JD_LUT_ENTRIES];
    __shared__ uint32_t s_tpos[64];
    __shared__ __align__(16) uint32_t s_ring[128 * JD_RING_STRIDE];
    __shared__ __align__(16) uint16_t s_stage[128 * 8];
    const uint32_t ii = a.cimg_list[blockIdx.y];
    const JDImageDesc &im = a.imgs[ii];
    if (blockIdx.x * 128u >= im.nch) return;
    if (threadIdx.x < 64) s_tpos[threadIdx.x] = jd_tposw(c_tpos[threadIdx.x]);
    jd_load_lut_set(s_lut, a.luts + (size_t)im.lutset * JD_LUT_ENTRIES);
    __syncthreads();
    const uint32_t c = blockIdx.x * 128u + threadIdx.x;
    if (c >= im.nch) return;
    const uint32_t g = im.chunk_base + c;
    const uint32_t total_blocks = (uint32_t)im.mcus_x * im.mcus_y * im.bpm;
    const uint32_t pre = a.cpre[g], fb = a.cfirst[g];
    uint32_t n = a.cn[g];
    n = (pre >= total_blocks) ? 0u : ((n < total_blocks - pre) ? n : total_blocks - pre);   /* bits after the last block are not blocks */
    uint32_t jmap = JD_JW_INIT, status = JD_SEG_OK, done = 0;
    if (n != 0u) {
        const uint32_t P0 = c * JD_CHUNK_BYTES * 8u + (fb & 0xFFFFu);       /* first block's first bit, relative to the scan */
        const uint32_t byte0 = im.scan_off + (P0 >> 3);
        JDSegIn in;
        in.data = a.filt;
        in.start = byte0 & ~15u;
        in.end = im.scan_off + a.flen[ii];
        in.nmcu = 0; in.bpm = im.bpm; in.ncomp = im.ncomp; in.tsel = im.tsel;
        in.skip_bits = (byte0 - in.start) * 8u + (P0 & 7u);
        in.blk_first = (fb >> 16) & 0xFu;
        in.nblk = n;
        in.midstream = 1;
        in.pred[0] = a.cpe[3 * g]; in.pred[1] = a.cpe[3 * g + 1]; in.pred[2] = a.cpe[3 * g + 2];
        /* image-relative record slot: the scan's one "segment" owns slot 0..nseg-1, its chunks follow */
        in.rec_index0 = JD_REC_INDEX(im.scan_off - im.comp_off + c * JD_CHUNK_BYTES, im.nseg + c);
        in.rec_cap = JD_REC_CAP(JD_CHUNK_BYTES);
        in.seg = a.nseg_total + g;             /* phase slot of this walk (events) */
        in.img = ii;
        in.blk0 = im.blk_base + pre;
        in.al = 0;
        in.ring = s_ring + threadIdx.x * JD_RING_STRIDE;
        in.stage = s_stage + threadIdx.x * 8;
        JDEventSinkDev sink{a.events, a.event_count, a.event_cap};
        JDSegOut so;
        jd_decode_segment<JDEventSinkDev, JD_MODE_BASELINE, true>(in, s_lut, s_tpos, a.blk_hdr + im.blk_base + pre, a.rec + im.rec_base + in.rec_index0, sink, so);
        jmap = so.jmap; status = so.status; done = so.nblk_done;
    }
    if (status == JD_SEG_OK && (fb >> 31) != 0u) status = JD_SEG_BADCODE;    /* the parse pass met an invalid code after these blocks */
    a.cjmap[g] = jmap;
    a.cstatus[g] = status;
    a.cnown[g] = done;
}

/* phase map composition: first `a`, then `b` (six nibbles: next phase for each current phase; both normalised to 0..5) */
__device__ __forceinline__ uint32_t jd_jmap_compose(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
#pragma unroll
    for (int p = 0; p < 6; p++) r |= ((b >> (4u * ((a >> (4 * p)) & 15u))) & 15u) << (4 * p);
    return r;
}

/* per restart-free scan (one warp): true window phase at each chunk entry = the composition of the phase maps of the chunks
 * to its left applied to phase 0 (a scan over the chunks, 32 at a time); folds the chunk statuses */
__global__ void __launch_bounds__(128) jdk_chunk_stitch(const JDChunkArgs a)
{
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (w >= a.ncimg) return;
    const JDImageDesc &im = a.imgs[a.cimg_list[w]];
    const uint32_t ident = 0x543210u;
    uint32_t carry = ident;              /* composition of every chunk before this group of 32 */
    uint32_t first_bad = 0xFFFFFFFFu;
    for (uint32_t cb = 0; cb < im.nch; cb += 32u) {
        const uint32_t c = cb + lane, g = im.chunk_base + c;
        const bool live = c < im.nch;
        const uint32_t m = live ? jd_jw_ckpt(a.cjmap[g]) : ident;      /* phases >= 6 restart at 0 */
        uint32_t x = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= (uint32_t)d) x = jd_jmap_compose(y, x);
        }
        uint32_t excl = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0) excl = ident;
        if (live) {
            a.seg_phase[a.nseg_total + g] = jd_jmap_compose(carry, excl) & 15u;
            if (a.cstatus[g] != 0u && c < first_bad) first_bad = c;
        }
        carry = jd_jmap_compose(carry, __shfl_sync(0xffffffffu, x, 31));
    }
    first_bad = __reduce_min_sync(0xffffffffu, first_bad);
    if (lane == 0) {
        uint32_t status = 0, err_mcu = 0;
        if (first_bad != 0xFFFFFFFFu) {
            const uint32_t g = im.chunk_base + first_bad;
            status = a.cstatus[g]; err_mcu = (a.cpre[g] + a.cnown[g]) / im.bpm;
        }
        /* the scan is one "segment" for the per-image stitch (jdk_stitch) */
        a.seg_jmap[im.seg_base] = JD_JW_INIT;
        a.seg_status[im.seg_base] = status ? ((status << 28) | (err_mcu & 0x0FFFFFFFu)) : 0u;
    }
}

/* ------------------------------------------------------------------------------------ */
/* fused expand + dequant + IDCT + colour                                                  */
/* ------------------------------------------------------------------------------------ */
#define JD_PT_565 0
#define JD_PT_8888 1
#define JD_PT_GRAY 2

struct JDIdctArgs {
    const JDImageDesc *imgs;
    const jd_u64 *blk_hdr;
    const uint16_t *rec;
    const int32_t *quant;   /* [img][3][64] int32, column-major per component: [c * 8 + r] */
    uint8_t *out;           /* output base */
    /* geometry shared by every image of this launch (the host groups images by size / sampling) */
    uint32_t mcus_x, mcus_y, width, height, bpm;
    uint32_t img0;          /* first image of this launch (blockIdx.z offset) */
    uint32_t big_endian;    /* RGB565_BIG_ENDIAN requested */
    uint32_t padded;        /* 1: write the whole MCU-aligned area (dither intermediate / callback replay) */
};

template <int HS, int VS, int NC, int MPB>
struct JDGeo {
    static constexpr int BPMEFF = HS * VS + (NC == 3 ? 2 : 0);
    static constexpr int NB = MPB * BPMEFF;
    static constexpr int THREADS = NB * 8;
    static constexpr int WCTA = MPB * HS * 8;
    static constexpr int HCTA = VS * 8;
    static constexpr int YSTRIDE = WCTA + 16; /* keeps 16-byte row alignment, spreads banks */
    static constexpr int CSTRIDE = MPB * 8 + 8;
    static constexpr int TSTRIDE = 72; /* halfwords per coefficient tile (64 + pad: bank spread) */
};

__device__ __forceinline__ void jd_unpack8(const uint4 v, int m[8])
{
    m[0] = (int)(short)(v.x & 0xFFFF); m[1] = (int)v.x >> 16;
    m[2] = (int)(short)(v.y & 0xFFFF); m[3] = (int)v.y >> 16;
    m[4] = (int)(short)(v.z & 0xFFFF); m[5] = (int)v.z >> 16;
    m[6] = (int)(short)(v.w & 0xFFFF); m[7] = (int)v.w >> 16;
}

/* d = { c[15:0] << 16 | sat_u8(a) << 8 | sat_u8(b) } : two saturating byte packs build a clamped pixel */
__device__ __forceinline__ uint32_t jd_pack_sat(int a, int b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

/* x / B for x < 4096 without a high multiply (IMAD.HI is a slow instruction on this part) */
template <int B>
__device__ __forceinline__ uint32_t jd_div_small(uint32_t x) { return (x * (uint32_t)(65536 / B + 1)) >> 16; }

__device__ __forceinline__ uint32_t jd_byte(uint32_t w, int i) { return (w >> (8 * i)) & 0xFFu; }

/* SSE2-build chroma terms with the >>16 of the int16 mulhi folded: ((C-128)<<8) * K >> 16 == ((C-128) * K) >> 8 */
__device__ __forceinline__ void jd_chroma_terms_sse(uint32_t Cb, uint32_t Cr, int &tr, int &tg, int &tb)
{
    tr = ((int)Cr * 5742 - 128 * 5742) >> 8;
    tg = (((int)Cr * -2925 + 128 * 2925) >> 8) + (((int)Cb * -1409 + 128 * 1409) >> 8);
    tb = ((int)Cb * 7258 - 128 * 7258) >> 8;
}

template <int PT>
__device__ __forceinline__ uint32_t jd_pixel_sse(uint32_t Y, int tr, int tg, int tb)
{
    const int Y4 = (int)Y << 4;
    const int R = (Y4 + tr) >> 4, G = (Y4 + tg) >> 4, B = (Y4 + tb) >> 4;
    if (PT == JD_PT_8888) return jd_pack_sat(G, B, jd_pack_sat(255, R, 0u));          /* bytes B,G,R,A */
    const uint32_t w = jd_pack_sat(G, B, jd_pack_sat(0, R, 0u));                      /* bytes B,G,R,0 */
    return ((w >> 8) & 0xF800u) | ((w >> 5) & 0x07E0u) | ((w >> 3) & 0x001Fu);
}

template <int PT>
__device__ __forceinline__ uint32_t jd_pixel_scalar(int Y12, int cb, int cr, bool big_endian)
{
    /* JPEGPixelLE/BE/RGB (jpeg.inl:3101-3278): cb, cr already minus 128 */
    const int B = (7258 * cb + Y12) >> 12, G = (-1409 * cb - 2925 * cr + Y12) >> 12, R = (5742 * cr + Y12) >> 12;
    if (PT == JD_PT_8888) return jd_pack_sat(G, R, jd_pack_sat(255, B, 0u));          /* bytes R,G,B,A */
    uint32_t v = ((jd_rt(R) >> 3) << 11) | ((jd_rt(G) >> 2) << 5) | (jd_rt(B) >> 3);
    if (big_endian) v = jd_bswap16(v);
    return v;
}

/* Phase C of the fused kernels (full size): colour conversion of the staged planes + coalesced 16-byte scanline stores.
 * s_y: (VS*8) rows x YSTRIDE luma bytes, s_cb/s_cr: 8 rows x CSTRIDE chroma bytes, covering WCTA pixels of MCU row `my`. */
template <int HS, int VS, int NC, int PT, int ARITH, int WCTA, int YSTRIDE, int CSTRIDE, int NTHREADS, bool INTERIOR = false>
__device__ __forceinline__ void jd_phase_c_full(const JDIdctArgs &a, const uint8_t *s_y, const uint8_t *s_cb, const uint8_t *s_cr,
                                                uint32_t strip, uint32_t my, uint32_t tid, uint32_t W, uint32_t H,
                                                uint8_t *outbase, uint32_t pitch)
{
    constexpr int BYPP = (PT == JD_PT_565) ? 2 : (PT == JD_PT_8888 ? 4 : 1);
    /* one item = PXI pixels (OWN 16-byte stores) in each of the VS rows that share chroma */
#ifndef JD_PXI_8888
#define JD_PXI_8888 4   /* 8-pixel items (two stores per row) measured slower: 3.37 -> 3.64 ms */
#endif
    constexpr int PXI = (PT == JD_PT_8888) ? JD_PXI_8888 : 16 / BYPP;   /* 8 (8888: two stores), 8 (565), 16 (gray) */
    constexpr int OWN = PXI * BYPP / 16;
    constexpr int IPR = WCTA / PXI;          /* items per row */
    constexpr int NITEM = IPR * 8;              /* x (HCTA / VS) row groups */
    constexpr bool SSE_PATH = (ARITH == JPEG_ARITH_SSE2) && (HS == VS); /* jpeg.inl:3409-3517, :4006-4308 */
    /* item (rg, xg): PXI pixels at x = xg * PXI in the VS rows of row group rg */
    auto item = [&](const uint32_t rg, const uint32_t xg) {
        const uint32_t gx = strip * WCTA + xg * PXI;
        if (!INTERIOR && gx >= W) return;
        const bool full = INTERIOR || (gx + PXI <= W);
        /* chroma samples covering these PXI pixels: PXI / HS of each */
        uint32_t cbw[2] = {0, 0}, crw[2] = {0, 0};
        if (NC == 3 && PT != JD_PT_GRAY) {
            constexpr int NCH = PXI / HS; /* 2, 4 or 8 bytes */
            const uint8_t *pb = s_cb + rg * CSTRIDE + xg * NCH, *pr = s_cr + rg * CSTRIDE + xg * NCH;
            if (NCH == 2) { cbw[0] = *reinterpret_cast<const uint16_t *>(pb); crw[0] = *reinterpret_cast<const uint16_t *>(pr); }
            else if (NCH == 4) { cbw[0] = *reinterpret_cast<const uint32_t *>(pb); crw[0] = *reinterpret_cast<const uint32_t *>(pr); }
            else { const uint2 u = *reinterpret_cast<const uint2 *>(pb), v = *reinterpret_cast<const uint2 *>(pr); cbw[0] = u.x; cbw[1] = u.y; crw[0] = v.x; crw[1] = v.y; }
        }
        /* SSE2-build path: packed (two-pixel) chroma terms, shared by the VS rows of this item */
        uint32_t tpk[PXI][3];
        if (NC == 3 && PT != JD_PT_GRAY && SSE_PATH) {
#pragma unroll
            for (int j = 0; j < PXI / 2; j++) {
                /* pixel pair j uses chroma sample j (HS == 2) or samples 2j, 2j+1 (HS == 1) */
                const int c0 = (HS == 2) ? j : 2 * j, c1 = (HS == 2) ? j : 2 * j + 1;
                const int tj = (HS == 2) ? j : 2 * j;
                int tr0, tg0, tb0, tr1, tg1, tb1;
                jd_chroma_terms_sse(jd_byte(cbw[c0 >> 2], c0 & 3), jd_byte(crw[c0 >> 2], c0 & 3), tr0, tg0, tb0);
                if (HS == 2) { tr1 = tr0; tg1 = tg0; tb1 = tb0; }
                else jd_chroma_terms_sse(jd_byte(cbw[c1 >> 2], c1 & 3), jd_byte(crw[c1 >> 2], c1 & 3), tr1, tg1, tb1);
                tpk[tj][0] = __byte_perm((uint32_t)tr0, (uint32_t)tr1, 0x5410);
                tpk[tj][1] = __byte_perm((uint32_t)tg0, (uint32_t)tg1, 0x5410);
                tpk[tj][2] = __byte_perm((uint32_t)tb0, (uint32_t)tb1, 0x5410);
            }
        }
#pragma unroll
        for (int vr = 0; vr < VS; vr++) {
            const uint32_t row = rg * VS + vr;
            const uint32_t gy = my * (VS * 8) + row;
            if (!INTERIOR && gy >= H) continue;
            uint32_t yw[4];
            {
                const uint8_t *py = s_y + row * YSTRIDE + xg * PXI;
                if (PXI == 4) yw[0] = *reinterpret_cast<const uint32_t *>(py);
                else if (PXI == 8) { const uint2 u = *reinterpret_cast<const uint2 *>(py); yw[0] = u.x; yw[1] = u.y; }
                else { const uint4 u = *reinterpret_cast<const uint4 *>(py); yw[0] = u.x; yw[1] = u.y; yw[2] = u.z; yw[3] = u.w; }
            }
            uint32_t ow[4 * OWN]; /* the output bytes of this row */
            if (PT == JD_PT_GRAY) {
                ow[0] = yw[0]; ow[1] = yw[1]; ow[2] = yw[2]; ow[3] = yw[3];
            } else {
                if (NC == 3 && SSE_PATH) {
                    /* SSE2-build arithmetic, two pixels per instruction: (Y<<4 + t) clamped to [0,4095] by one
                     * VIADDMNMX.S16x2.RELU per channel, then >>4 (== packus((Y4 + t) >> 4)); tpk[] computed above */
#pragma unroll
                    for (int j = 0; j < PXI / 2; j++) {
                        const uint32_t ywj = yw[j >> 1];
                        const uint32_t y4 = __byte_perm(ywj, 0, (j & 1) ? 0x4342 : 0x4140) << 4; /* Y(2j)<<4 | Y(2j+1)<<4 << 16 */
                        const int tj = (HS == 2) ? j : 2 * j;  /* index into the packed chroma terms */
                        const uint32_t r12 = __viaddmin_s16x2_relu(y4, tpk[tj][0], 0x0FFF0FFFu);
                        const uint32_t g12 = __viaddmin_s16x2_relu(y4, tpk[tj][1], 0x0FFF0FFFu);
                        const uint32_t b12 = __viaddmin_s16x2_relu(y4, tpk[tj][2], 0x0FFF0FFFu);
                        if (PT == JD_PT_8888) {
                            /* 12-bit values << 4: bytes 1 and 3 hold the two pixels (a left shift issues on the FMA pipe,
                             * the ALU pipe is this kernel's bound) */
                            const uint32_t rs = r12 << 4, gs = g12 << 4, bs = b12 << 4;
                            const uint32_t bg = __byte_perm(bs, gs, 0x7351);             /* B0 G0 B1 G1 */
                            const uint32_t ra = __byte_perm(rs, 0xFFFFFFFFu, 0x4341);    /* R0 FF R1 FF */
                            ow[2 * j] = __byte_perm(bg, ra, 0x5410);
                            ow[2 * j + 1] = __byte_perm(bg, ra, 0x7632);
                        } else {
                            ow[j] = ((r12 << 4) & 0xF800F800u) | ((g12 >> 1) & 0x07E007E0u) | ((b12 >> 7) & 0x001F001Fu);
                        }
                    }
                } else {
                uint32_t pix[PXI];
#pragma unroll
                for (int i = 0; i < PXI; i++) {
                    const uint32_t Y = jd_byte(yw[i >> 2], i & 3);
                    if (NC == 1) {
                        uint32_t v = jd_gray565(Y);
                        if (a.big_endian) v = jd_bswap16(v);
                        pix[i] = v;
                    } else {
                        const int ci = i / HS;
                        const uint32_t Cb = jd_byte(cbw[ci >> 2], ci & 3), Cr = jd_byte(crw[ci >> 2], ci & 3);
                        pix[i] = jd_pixel_scalar<PT>((int)Y << 12, (int)Cb - 128, (int)Cr - 128, a.big_endian != 0u);
                    }
                }
                if (PT == JD_PT_8888) {
#pragma unroll
                    for (int i = 0; i < PXI; i++) ow[i] = pix[i];
                }
                else {
#pragma unroll
                    for (int i = 0; i < 4; i++) ow[i] = pix[(2 * i) % PXI] | (pix[(2 * i + 1) % PXI] << 16);
                }
                }
            }
            uint8_t *dst = outbase + (size_t)gy * pitch + (size_t)gx * BYPP;
            if (INTERIOR || (full && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0))) {
#pragma unroll
                for (int q = 0; q < OWN; q++) reinterpret_cast<uint4 *>(dst)[q] = make_uint4(ow[4 * q], ow[4 * q + 1], ow[4 * q + 2], ow[4 * q + 3]);
            } else {
                for (uint32_t i = 0; i < (uint32_t)PXI && gx + i < W; i++) {
                    if (BYPP == 4) reinterpret_cast<uint32_t *>(dst)[i] = ow[i % (4 * OWN)];
                    else if (BYPP == 2) reinterpret_cast<uint16_t *>(dst)[i] = (uint16_t)(ow[(i >> 1) & 3] >> ((i & 1) * 16));
                    else dst[i] = (uint8_t)(ow[(i >> 2) & 3] >> ((i & 3) * 8));
                }
            }
        }
        };
    if (NTHREADS % IPR == 0) {
        /* the thread keeps its x position; only the row group advances (no division, x addressing hoisted) */
        const uint32_t xg = tid % IPR;
        for (uint32_t rg = tid / IPR; rg < 8u; rg += NTHREADS / IPR) item(rg, xg);
    } else {
        for (uint32_t it = tid; it < (uint32_t)NITEM; it += NTHREADS) { const uint32_t rg = jd_div_small<IPR>(it); item(rg, it - rg * IPR); }
    }
}

/* Phase C at 1/2 scale: 2x2 luma sums; scalar colour code in both builds (jpeg.inl:3297-3322, :3577-3626) */
template <int HS, int VS, int NC, int PT, int WCTA, int HCTA, int YSTRIDE, int CSTRIDE, int NTHREADS>
__device__ __forceinline__ void jd_phase_c_half(const JDIdctArgs &a, const uint8_t *s_y, const uint8_t *s_cb, const uint8_t *s_cr,
                                                uint32_t strip, uint32_t my, uint32_t tid, uint32_t W, uint32_t H,
                                                uint8_t *outbase, uint32_t pitch)
{
    constexpr int BYPP = (PT == JD_PT_565) ? 2 : (PT == JD_PT_8888 ? 4 : 1);
    const uint32_t OW = (W + 1) >> 1, OH = (H + 1) >> 1;
    constexpr int OWC = WCTA / 2, OHC = HCTA / 2;
    for (uint32_t it = tid; it < (uint32_t)(OWC * OHC); it += NTHREADS) {
        const uint32_t oy = it / OWC, ox = it - oy * OWC;
        const uint32_t gy = my * OHC + oy, gx = strip * OWC + ox;
        if (gy >= OH || gx >= OW) continue;
        const uint8_t *yp = s_y + (2 * oy) * YSTRIDE + 2 * ox;
        const int sum = yp[0] + yp[1] + yp[YSTRIDE] + yp[YSTRIDE + 1];
        uint8_t *dst = outbase + (size_t)gy * pitch + (size_t)gx * BYPP;
        if (PT == JD_PT_GRAY) {
            *dst = (uint8_t)((sum + 2) >> 2);
        } else if (NC == 1) {
            uint32_t v = jd_gray565((uint32_t)((sum + 2) >> 2));
            if (a.big_endian) v = jd_bswap16(v);
            *reinterpret_cast<uint16_t *>(dst) = (uint16_t)v;
        } else {
            int Cb, Cr;
            if (HS == 2 && VS == 2) {
                Cb = s_cb[oy * CSTRIDE + ox]; Cr = s_cr[oy * CSTRIDE + ox];
            } else if (HS == 1 && VS == 1) {
                const uint8_t *p1 = s_cb + (2 * oy) * CSTRIDE + 2 * ox, *p2 = s_cr + (2 * oy) * CSTRIDE + 2 * ox;
                Cb = (p1[0] + p1[1] + p1[CSTRIDE] + p1[CSTRIDE + 1] + 2) >> 2;
                Cr = (p2[0] + p2[1] + p2[CSTRIDE] + p2[CSTRIDE + 1] + 2) >> 2;
            } else if (HS == 2) {
                Cb = (s_cb[(2 * oy) * CSTRIDE + ox] + s_cb[(2 * oy + 1) * CSTRIDE + ox] + 1) >> 1;
                Cr = (s_cr[(2 * oy) * CSTRIDE + ox] + s_cr[(2 * oy + 1) * CSTRIDE + ox] + 1) >> 1;
            } else {
                Cb = (s_cb[oy * CSTRIDE + 2 * ox] + s_cb[oy * CSTRIDE + 2 * ox + 1] + 1) >> 1;
                Cr = (s_cr[oy * CSTRIDE + 2 * ox] + s_cr[oy * CSTRIDE + 2 * ox + 1] + 1) >> 1;
            }
            const uint32_t v = jd_pixel_scalar<PT>(sum << 10, Cb - 128, Cr - 128, a.big_endian != 0u);
            if (PT == JD_PT_8888) *reinterpret_cast<uint32_t *>(dst) = v;
            else *reinterpret_cast<uint16_t *>(dst) = (uint16_t)v;
        }
    }
}

template <int HS, int VS, int NC, int MPB, int PT, int ARITH, bool HALF>
__global__ void __launch_bounds__(JDGeo<HS, VS, NC, MPB>::THREADS)
jdk_idct_color(const JDIdctArgs a)
{
    using G = JDGeo<HS, VS, NC, MPB>;
    __shared__ __align__(16) int16_t s_tile[G::NB * G::TSTRIDE];
    __shared__ __align__(16) uint8_t s_y[G::HCTA * G::YSTRIDE];
    __shared__ __align__(16) uint8_t s_c[(NC == 3 ? 2 : 1) * 8 * G::CSTRIDE];

    const uint32_t img_i = a.img0 + blockIdx.z;
    const JDImageDesc &im = a.imgs[img_i];
    const uint32_t strip = blockIdx.x, my = blockIdx.y;
    const uint32_t tid = threadIdx.x;

    /* ---- phase A: expand this block's records into a column-major coefficient tile ---- */
    const uint32_t gb = tid >> 3, c = tid & 7;           /* block within CTA, lane within block */
    const uint32_t ml = jd_div_small<G::BPMEFF>(gb), blk = gb - ml * G::BPMEFF;
    const uint32_t mx = strip * MPB + ml;
    const uint32_t comp = (blk < (uint32_t)(HS * VS)) ? 0u : blk - HS * VS + 1u;
    jd_u64 h = 0;
    /* (block order inside an MCU in the stream = luma blocks, Cb, Cr = our blk numbering) */
    if (mx < a.mcus_x) h = __ldg(a.blk_hdr + im.blk_base + (my * a.mcus_x + mx) * a.bpm + blk);
    const uint16_t *const irec = a.rec + im.rec_base;
    const uint32_t ri = JD_HDR_REC(h);
    const int dc = JD_HDR_DC(h);
    const uint32_t ncoef = JD_HDR_NCOEF(h);
    int16_t *tile = s_tile + gb * G::TSTRIDE;
    uint32_t px0, px1; /* 8 output bytes of row `c` of this block */
    const int32_t *qg = a.quant + (size_t)img_i * 192 + comp * 64;
    if (__builtin_expect(__all_sync(0xffffffffu, ncoef == 0u), 0)) {
        /* no stored AC coefficient in any of the warp's 4 blocks: DC-only fill (jpeg.inl:5146-5154) */
        px0 = px1 = jd_range(dc * __ldg(qg)) * 0x01010101u;
    } else {
        *reinterpret_cast<uint4 *>(tile + c * 8) = make_uint4(0, 0, 0, 0);
        const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(qg + c * 8));
        const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(qg + c * 8 + 4));
        __syncwarp();
        if (!JD_HDR_BIG(h)) {
            const uint16_t *rp = irec + ri + c;
            if (c < ncoef) { const uint32_t r = __ldg(rp); tile[r >> 10] = (int16_t)((int)(r << 22) >> 22); }
            if (c + 8 < ncoef) { const uint32_t r = __ldg(rp + 8); tile[r >> 10] = (int16_t)((int)(r << 22) >> 22); }
            for (uint32_t i = c + 16; i < ncoef; i += 8) {
                const uint32_t r = __ldg(irec + ri + i);
                tile[r >> 10] = (int16_t)((int)(r << 22) >> 22);
            }
        } else {
            for (uint32_t i = c; i < ncoef; i += 8) {
                const uint32_t t = __ldg(irec + ri + 2 * i) & 63u;
                tile[t] = (int16_t)__ldg(irec + ri + 2 * i + 1);
            }
        }
        __syncwarp();
        /* ---- phase B: dequant + column pass (lane = column), row pass (lane = row) ---- */
        int m[8], o[8];
        const int qq[8] = {(int)q0.x, (int)q0.y, (int)q0.z, (int)q0.w, (int)q1.x, (int)q1.y, (int)q1.z, (int)q1.w};
        jd_unpack8(*reinterpret_cast<const uint4 *>(tile + c * 8), m);
        if (c == 0) m[0] = dc;
        const bool r47 = JD_HDR_HI(h) == 0u;
        if (ARITH == JPEG_ARITH_SSE2) {
#pragma unroll
            for (int r = 0; r < 8; r++) m[r] *= qq[r];
            jd_col_sse16(m, r47, o);
        } else {
            jd_col_scalar(m, qq, r47, o);
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; r++) tile[r * 8 + c] = (int16_t)o[r];
        __syncwarp();
        int p[8];
        uint32_t ob[8];
        jd_unpack8(*reinterpret_cast<const uint4 *>(tile + c * 8), p);
        jd_row_raw(p, JD_HDR_COLMASK(h), (int *)ob);
        /* ucRangeTable as arithmetic + two saturating packs per 4 bytes */
#pragma unroll
        for (int i = 0; i < 8; i++) ob[i] = (uint32_t)((((int)ob[i] << 17) >> 22) + 128);
        px0 = jd_pack_sat((int)ob[1], (int)ob[0], jd_pack_sat((int)ob[3], (int)ob[2], 0u));
        px1 = jd_pack_sat((int)ob[5], (int)ob[4], jd_pack_sat((int)ob[7], (int)ob[6], 0u));
    }
    /* stage the pixel bytes */
    if (comp == 0) {
        const uint32_t lx = (HS == 2) ? (blk & 1u) : 0u;
        const uint32_t ly = (HS == 2 && VS == 2) ? (blk >> 1) : ((VS == 2) ? blk : 0u);
        *reinterpret_cast<uint2 *>(s_y + (ly * 8 + c) * G::YSTRIDE + (ml * HS + lx) * 8) = make_uint2(px0, px1);
    } else {
        *reinterpret_cast<uint2 *>(s_c + ((comp - 1) * 8 + c) * G::CSTRIDE + ml * 8) = make_uint2(px0, px1);
    }
    __syncthreads();

    /* ---- phase C: colour conversion + coalesced 128-bit scanline stores ---- */
    const uint8_t *s_cb = s_c, *s_cr = s_c + 8 * G::CSTRIDE;
    const uint32_t W = a.padded ? a.mcus_x * HS * 8 : a.width;
    const uint32_t H = a.padded ? a.mcus_y * VS * 8 : a.height;
    uint8_t *outbase = a.out + im.out_off;
    const uint32_t pitch = im.out_pitch;
    constexpr int BYPP = (PT == JD_PT_565) ? 2 : (PT == JD_PT_8888 ? 4 : 1);

    if (!HALF) {
        jd_phase_c_full<HS, VS, NC, PT, ARITH, G::WCTA, G::YSTRIDE, G::CSTRIDE, G::THREADS>(a, s_y, s_cb, s_cr, strip, my, tid, W, H, outbase, pitch);
    } else {
        jd_phase_c_half<HS, VS, NC, PT, G::WCTA, G::HCTA, G::YSTRIDE, G::CSTRIDE, G::THREADS>(a, s_y, s_cb, s_cr, strip, my, tid, W, H, outbase, pitch);
    }
}

/* ------------------------------------------------------------------------------------ */
/* fused expand + dequant + IDCT + colour, one THREAD per 8x8 block                          */
/*                                                                                          */
/* At the qualities the benchmark uses ~90 % of the blocks hold coefficients only in their  */
/* top-left 4x4 (rows 4-7 empty, columns 4-7 empty).  With 8 lanes per block half the lanes  */
/* then transform empty columns.  Here a thread owns a block: it expands the records into   */
/* its private tile, runs the column pass only over populated columns (4 of them for the    */
/* common class, results kept in registers -- no transpose through shared memory, no warp    */
/* syncs) and the 8 row passes.  Blocks are first binned by class inside the CTA so that     */
/* the lanes of a warp take the same path.  Colour phase shared with jdk_idct_color.         */
/* ------------------------------------------------------------------------------------ */
template <int HS, int VS, int NC, int MPB>
struct JDGeoTB {
    static constexpr int BPMEFF = HS * VS + (NC == 3 ? 2 : 0);
    static constexpr int NB = MPB * BPMEFF;                       /* blocks per CTA */
    static constexpr int NW = (NB + 31) / 32 > 4 ? (NB + 31) / 32 : 4;
    static constexpr int THREADS = NW * 32;
    static constexpr int WCTA = MPB * HS * 8;
    static constexpr int HCTA = VS * 8;
    static constexpr int YSTRIDE = WCTA + 16;
    static constexpr int CSTRIDE = MPB * 8 + 8;
    static constexpr int TSTRIDE = 72;                            /* int16 per full tile (144 B: conflict-free LDS.128 per quarter warp) */
    static constexpr int T0STRIDE = 40;                           /* int16 per 4-column tile (80 B: 16-byte aligned, conflict-free LDS.128 per quarter warp) */
};

__device__ __forceinline__ void jd_unpack4(const uint2 v, int m[4])
{
    m[0] = (int)(short)(v.x & 0xFFFF); m[1] = (int)v.x >> 16;
    m[2] = (int)(short)(v.y & 0xFFFF); m[3] = (int)v.y >> 16;
}

/* the 8 butterflies that end a row pass + the ucRangeTable clamp, two pixels per instruction: jd_core.h jd_row_finish2 */
__device__ __forceinline__ uint2 jd_row_finish_packed(const int t[8])
{
    uint2 r;
    jd_row_finish2(t, &r.x, &r.y);
    return r;
}

#ifndef JD_TB_MINB
#define JD_TB_MINB 10   /* 48 registers: 10 CTAs per SM measured faster than 56 registers / 9 CTAs and than 40 / 12 */
#endif
template <int HS, int VS, int NC, int MPB, int PT, int ARITH>
__global__ void __launch_bounds__(JDGeoTB<HS, VS, NC, MPB>::THREADS, JD_TB_MINB)
jdk_idct_tb(const JDIdctArgs a)
{
    using G = JDGeoTB<HS, VS, NC, MPB>;
    __shared__ __align__(16) int16_t s_tile0[G::NB * G::T0STRIDE];          /* common class: 4 columns x 4 rows, one per thread */
    __shared__ __align__(16) int16_t s_tileL[G::NW * 4 * G::TSTRIDE];       /* 8-lane mode: 4 full tiles per warp */
    __shared__ __align__(16) uint8_t s_y[G::HCTA * G::YSTRIDE];
    __shared__ __align__(16) uint8_t s_c[(NC == 3 ? 2 : 1) * 8 * G::CSTRIDE];
    __shared__ jd_u64 s_hdr[G::NB];
    __shared__ uint16_t s_perm[G::NB];
    __shared__ uint32_t s_wc[3][G::NW];

    const uint32_t img_i = a.img0 + blockIdx.z;
    const JDImageDesc &im = a.imgs[img_i];
    const uint32_t strip = blockIdx.x, my = blockIdx.y;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint16_t *const irec = a.rec + im.rec_base;

    /* ---- headers + binning.  Thread-per-block classes: coefficients in columns 0-3 only (3-4 columns occupied), split by
     * whether rows 4-7 are empty (the reference picks its reduced column pass on that flag, jpeg.inl:2330): class 0 = rows
     * 4-7 empty, class 1 = not.  Everything else (class 2: > 4 columns, <= 2 columns, DC only) goes to the 8-lane passes. ---- */
    uint32_t cls = 3;
    if (tid < (uint32_t)G::NB) {
        const uint32_t ml = jd_div_small<G::BPMEFF>(tid), blk = tid - ml * G::BPMEFF;
        const uint32_t mx = strip * MPB + ml;
        if (mx < a.mcus_x) {
            const jd_u64 h = __ldg(a.blk_hdr + im.blk_base + (my * a.mcus_x + mx) * a.bpm + blk);
            s_hdr[tid] = h;
            const uint32_t n = JD_HDR_NCOEF(h), cm = JD_HDR_COLMASK(h);
            cls = (n != 0u && (cm & 0xF0u) == 0u && (cm & 0xFCu) != 0u) ? JD_HDR_HI(h) : 2u;
        }
    }
    const uint32_t b0 = __ballot_sync(0xffffffffu, cls == 0u), b1 = __ballot_sync(0xffffffffu, cls == 1u), b2 = __ballot_sync(0xffffffffu, cls == 2u);
    if (lane == 0) { s_wc[0][wid] = __popc(b0); s_wc[1][wid] = __popc(b1); s_wc[2][wid] = __popc(b2); }
    __syncthreads();
    uint32_t n0a = 0, n0b = 0, n1 = 0, pre = 0;
    {
        uint32_t before0 = 0, before1 = 0, before2 = 0;
#pragma unroll
        for (int w2 = 0; w2 < G::NW; w2++) {
            const uint32_t c0 = s_wc[0][w2], c1 = s_wc[1][w2], c2 = s_wc[2][w2];
            if ((uint32_t)w2 < wid) { before0 += c0; before1 += c1; before2 += c2; }
            n0a += c0; n0b += c1; n1 += c2;
        }
        const uint32_t lt = (1u << lane) - 1u;
        pre = (cls == 0u) ? before0 + __popc(b0 & lt) : (cls == 1u) ? n0a + before1 + __popc(b1 & lt) : n0a + n0b + before2 + __popc(b2 & lt);
    }
    if (cls < 3u) s_perm[pre] = (uint16_t)tid;
    __syncthreads();
    const uint32_t n0 = n0a + n0b;

    /* ---- phases A + B, common class: this thread's block ---- */
    if (tid < n0) {
        const uint32_t pb = s_perm[tid];
        const jd_u64 h = s_hdr[pb];
        const uint32_t ml = jd_div_small<G::BPMEFF>(pb), blk = pb - ml * G::BPMEFF;
        const uint32_t comp = (blk < (uint32_t)(HS * VS)) ? 0u : blk - HS * VS + 1u;
        const uint32_t ri = JD_HDR_REC(h), ncoef = JD_HDR_NCOEF(h);
        const int dc = JD_HDR_DC(h);
        const int32_t *q = a.quant + (size_t)img_i * 192 + comp * 64;   /* L1-resident */
        int16_t *tile = s_tile0 + tid * G::T0STRIDE;   /* positions c * 8 + r, c < 4 */
        const bool r47 = tid < n0a;                    /* rows 4-7 empty: the reduced column pass */
        uint8_t *prow;  /* first output row of this block in the staged plane */
        uint32_t pstride;
        if (comp == 0) {
            const uint32_t lx = (HS == 2) ? (blk & 1u) : 0u;
            const uint32_t ly = (HS == 2 && VS == 2) ? (blk >> 1) : ((VS == 2) ? blk : 0u);
            prow = s_y + (ly * 8) * G::YSTRIDE + (ml * HS + lx) * 8; pstride = G::YSTRIDE;
        } else {
            prow = s_c + ((comp - 1) * 8) * G::CSTRIDE + ml * 8; pstride = G::CSTRIDE;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) *reinterpret_cast<uint4 *>(tile + c * 8) = make_uint4(0, 0, 0, 0);
        if (!JD_HDR_BIG(h)) {
            /* the first 10 halfwords that cover the records come in as five independent aligned 32-bit loads (one round
             * trip instead of a chain of 2-byte loads); longer blocks finish in the loop below */
            const uint32_t off = ri & 1u, total = off + ncoef;
            const uint32_t *w32 = reinterpret_cast<const uint32_t *>(irec + (ri - off));
            uint32_t v[5];
#pragma unroll
            for (int w = 0; w < 5; w++) v[w] = ((uint32_t)(2 * w) < total) ? __ldg(w32 + w) : 0u;
#pragma unroll
            for (int hh = 0; hh < 10; hh++) {
                if ((uint32_t)hh >= off && (uint32_t)hh < total) {
                    const uint32_t r = (hh & 1) ? (v[hh >> 1] >> 16) : (v[hh >> 1] & 0xFFFFu);
                    tile[r >> 10] = (int16_t)((int)(r << 22) >> 22);
                }
            }
            for (uint32_t i = 10u - off; i < ncoef; i++) { const uint32_t r = __ldg(irec + ri + i); tile[r >> 10] = (int16_t)((int)(r << 22) >> 22); }
        } else {
            for (uint32_t i = 0; i < ncoef; i++) tile[__ldg(irec + ri + 2 * i) & 63u] = (int16_t)__ldg(irec + ri + 2 * i + 1);
        }
        int cr[8][4]; /* column-pass results (as int16 values), [row][column] */
        if (r47) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int m[8] = {0, 0, 0, 0, 0, 0, 0, 0}, qq[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o[8];
                jd_unpack4(*reinterpret_cast<const uint2 *>(tile + c * 8), m);
                { const uint4 qv = __ldg(reinterpret_cast<const uint4 *>(q + c * 8)); qq[0] = (int)qv.x; qq[1] = (int)qv.y; qq[2] = (int)qv.z; qq[3] = (int)qv.w; }
                if (c == 0) m[0] = dc;
                if (ARITH == JPEG_ARITH_SSE2) {
#pragma unroll
                    for (int r = 0; r < 4; r++) m[r] *= qq[r];
                    if (c == 0) m[0] += JD_ROW_BIAS;   /* mod 2^16, additive through both passes */
                    jd_col_sse16(m, true, o);
                } else {
                    jd_col_scalar(m, qq, true, o);
                }
#pragma unroll
                for (int r = 0; r < 8; r++) cr[r][c] = (c == 0) ? o[r] : (int)(short)o[r];   /* column 0 only ever gets added: mod 2^16 is enough */
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int m[8], qq[8], o[8];
                jd_unpack8(*reinterpret_cast<const uint4 *>(tile + c * 8), m);
                {
                    const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(q + c * 8)), q1 = __ldg(reinterpret_cast<const uint4 *>(q + c * 8 + 4));
                    qq[0] = (int)q0.x; qq[1] = (int)q0.y; qq[2] = (int)q0.z; qq[3] = (int)q0.w;
                    qq[4] = (int)q1.x; qq[5] = (int)q1.y; qq[6] = (int)q1.z; qq[7] = (int)q1.w;
                }
                if (c == 0) m[0] = dc;
                if (ARITH == JPEG_ARITH_SSE2) {
#pragma unroll
                    for (int r = 0; r < 8; r++) m[r] *= qq[r];
                    if (c == 0) m[0] += JD_ROW_BIAS;
                    jd_col_sse16(m, false, o);
                } else {
                    jd_col_scalar(m, qq, false, o);
                }
#pragma unroll
                for (int r = 0; r < 8; r++) cr[r][c] = (c == 0) ? o[r] : (int)(short)o[r];   /* column 0 only ever gets added: mod 2^16 is enough */
            }
        }
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int p8[8] = {cr[r][0] + (ARITH == JPEG_ARITH_SSE2 ? 0 : JD_ROW_BIAS), cr[r][1], cr[r][2], cr[r][3], 0, 0, 0, 0};
            int t8[8];
            jd_row_terms(p8, 0x0Fu, t8);     /* 4-column variant (jpeg.inl:2698-2718) */
            *reinterpret_cast<uint2 *>(prow + r * pstride) = jd_row_finish_packed(t8);
        }
    }
    /* ---- phases A + B, every other block: 8 lanes per block (lane = column, then row), 4 blocks per warp pass.
     * The passes go to the warps that hold no common-class block when there are such warps (they would otherwise idle
     * at the barrier), else round-robin over all warps. ---- */
    {
        const uint32_t npass = (n1 + 3u) >> 2;
        const uint32_t busy = (n0 + 31u) >> 5;
        const uint32_t nfree = (busy < (uint32_t)G::NW) ? (uint32_t)G::NW - busy : 0u;
        /* a thread-per-block warp runs ~720 instructions, a pass ~200: the free warps take up to 4 passes each first,
         * what remains goes round-robin over all warps */
        const uint32_t base = (npass < 4u * nfree) ? npass : 4u * nfree;
        {
            const uint32_t c = lane & 7u;
            /* this warp's passes: (free warps only) wid - busy, + nfree, ... below `base`, then base + wid, + NW, ...
             * (one loop on purpose: two loops around a shared body measured 12 % slower on UHD q85) */
            bool first = wid >= busy;
            for (uint32_t j = first ? wid - busy : base + wid;; j += first ? nfree : (uint32_t)G::NW) {
                if (first && j >= base) { first = false; j = base + wid; }
                if (!first && j >= npass) break;
                const uint32_t oi = j * 4u + (lane >> 3);
                const bool valid = oi < n1;
                const uint32_t pb = valid ? s_perm[n0 + oi] : 0u;
                const jd_u64 h = valid ? s_hdr[pb] : 0;
                const uint32_t ml = jd_div_small<G::BPMEFF>(pb), blk = pb - ml * G::BPMEFF;
                const uint32_t comp = (blk < (uint32_t)(HS * VS)) ? 0u : blk - HS * VS + 1u;
                const uint32_t ri = JD_HDR_REC(h), ncoef = JD_HDR_NCOEF(h);
                const int dc = JD_HDR_DC(h);
                const int32_t *qg = a.quant + (size_t)img_i * 192 + comp * 64;
                int16_t *tile = s_tileL + (wid * 4u + (lane >> 3)) * G::TSTRIDE;
                uint2 px;
                if (__all_sync(0xffffffffu, ncoef == 0u)) {
                    /* DC only (jpeg.inl:5146-5154) */
                    px.x = px.y = jd_range(dc * __ldg(qg)) * 0x01010101u;
                } else {
                    if (valid) *reinterpret_cast<uint4 *>(tile + c * 8) = make_uint4(0, 0, 0, 0);
                    const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(qg + c * 8));
                    const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(qg + c * 8 + 4));
                    __syncwarp();
                    if (!JD_HDR_BIG(h)) {
                        for (uint32_t i = c; i < ncoef; i += 8) { const uint32_t r = __ldg(irec + ri + i); tile[r >> 10] = (int16_t)((int)(r << 22) >> 22); }
                    } else {
                        for (uint32_t i = c; i < ncoef; i += 8) tile[__ldg(irec + ri + 2 * i) & 63u] = (int16_t)__ldg(irec + ri + 2 * i + 1);
                    }
                    __syncwarp();
                    int m[8], o[8];
                    const int qq[8] = {(int)q0.x, (int)q0.y, (int)q0.z, (int)q0.w, (int)q1.x, (int)q1.y, (int)q1.z, (int)q1.w};
                    if (valid) jd_unpack8(*reinterpret_cast<const uint4 *>(tile + c * 8), m);
                    else { for (int r = 0; r < 8; r++) m[r] = 0; }
                    if (c == 0) m[0] = dc;
                    const bool r47 = JD_HDR_HI(h) == 0u;
                    if (ARITH == JPEG_ARITH_SSE2) {
#pragma unroll
                        for (int r = 0; r < 8; r++) m[r] *= qq[r];
                        jd_col_sse16(m, r47, o);
                    } else {
                        jd_col_scalar(m, qq, r47, o);
                    }
                    __syncwarp();
                    if (valid) {
#pragma unroll
                        for (int r = 0; r < 8; r++) tile[r * 8 + c] = (int16_t)o[r];
                    }
                    __syncwarp();
                    int p8[8], t8[8];
                    if (valid) jd_unpack8(*reinterpret_cast<const uint4 *>(tile + c * 8), p8);
                    else { for (int r = 0; r < 8; r++) p8[r] = 0; }
                    p8[0] += JD_ROW_BIAS;
                    jd_row_terms(p8, JD_HDR_COLMASK(h), t8);
                    px = jd_row_finish_packed(t8);
                }
                if (valid) {
                    if (comp == 0) {
                        const uint32_t lx = (HS == 2) ? (blk & 1u) : 0u;
                        const uint32_t ly = (HS == 2 && VS == 2) ? (blk >> 1) : ((VS == 2) ? blk : 0u);
                        *reinterpret_cast<uint2 *>(s_y + (ly * 8 + c) * G::YSTRIDE + (ml * HS + lx) * 8) = px;
                    } else {
                        *reinterpret_cast<uint2 *>(s_c + ((comp - 1) * 8 + c) * G::CSTRIDE + ml * 8) = px;
                    }
                }
            }
        }
    }
    __syncthreads();

    /* ---- phase C ---- */
    const uint32_t W = a.padded ? a.mcus_x * HS * 8 : a.width;
    const uint32_t H = a.padded ? a.mcus_y * VS * 8 : a.height;
    uint8_t *outbase = a.out + im.out_off;
    const uint32_t pitch = im.out_pitch;
    if ((strip + 1) * G::WCTA <= W && (my + 1) * G::HCTA <= H && ((reinterpret_cast<uintptr_t>(outbase) | pitch) & 15u) == 0u)
        jd_phase_c_full<HS, VS, NC, PT, ARITH, G::WCTA, G::YSTRIDE, G::CSTRIDE, G::THREADS, true>(a, s_y, s_c, s_c + 8 * G::CSTRIDE, strip, my, tid, W, H, outbase, pitch);
    else
        jd_phase_c_full<HS, VS, NC, PT, ARITH, G::WCTA, G::YSTRIDE, G::CSTRIDE, G::THREADS, false>(a, s_y, s_c, s_c + 8 * G::CSTRIDE, strip, my, tid, W, H, outbase, pitch);
}

/* ------------------------------------------------------------------------------------ */
/* fused expand + dequant + IDCT + colour, one THREAD per 8x8 block, two columns per         */
/* register (SSE2-build arithmetic; every sampling, full and half size).                     */
/*                                                                                          */
/* CTA = a strip of MPB MCUs of one MCU row, at most 128 blocks, one per thread.  The blocks */
/* are first binned inside the CTA -- rows 4-7 empty / populated (the reference's two column  */
/* pass variants, jpeg.inl:2330) x coefficients within columns 0-3 / beyond -- so that the    */
/* lanes of a warp mostly run the same code.  A thread then expands its block's records into  */
/* a private row-major tile in shared memory, dequantising on the way (int16 wrap, like the   */
/* reference's _mm_mullo_epi16), pulls the tile into registers as 8 rows x 2 or 4 column       */
/* pairs, and runs jd_idct_block_packed (jd_core.h): packed column pass in place, row pass,   */
/* pixels into the CTA's planes.  No transpose through shared memory, no warp                 */
/* synchronisation.  A warp that holds only blocks confined to columns 0-3 uses the           */
/* two-pair instantiation (half the registers and column passes); a mixed warp takes the       */
/* four-pair one for all its lanes, which gives identical results.  Colour phase as in the    */
/* other fused kernels.                                                                       */
/* ------------------------------------------------------------------------------------ */
template <int HS, int VS, int NC, int MPB>
struct JDGeoP {
    static constexpr int BPMEFF = HS * VS + (NC == 3 ? 2 : 0);
    static constexpr int NB = MPB * BPMEFF;                       /* blocks per CTA (<= THREADS) */
    static constexpr int THREADS = 128;
    static constexpr int WCTA = MPB * HS * 8;
    static constexpr int HCTA = VS * 8;
    static constexpr int YSTRIDE = WCTA + 16;
    static constexpr int CSTRIDE = MPB * 8 + 8;
    static constexpr int TWORDS = 36;                             /* words per private tile: 32 + 4 (conflict-free LDS.128 per quarter warp) */
};

#ifndef JD_P_MINB
#define JD_P_MINB 7
#endif
template <int HS, int VS, int NC, int MPB, int PT, bool HALF>
__global__ void __launch_bounds__(128, JD_P_MINB)
jdk_idct_p(const JDIdctArgs a)
{
    using G = JDGeoP<HS, VS, NC, MPB>;
    static_assert(G::NB <= G::THREADS, "one thread per block");
    __shared__ __align__(16) uint32_t s_tile[G::NB * G::TWORDS];
    __shared__ __align__(16) uint8_t s_y[G::HCTA * G::YSTRIDE];
    __shared__ __align__(16) uint8_t s_c[(NC == 3 ? 2 : 1) * 8 * G::CSTRIDE];
    __shared__ jd_u64 s_hdr[G::NB];
    __shared__ __align__(16) uint32_t s_wc[4][4];
    __shared__ uint16_t s_q[3 * 64];                              /* prescaled quant, natural order, low 16 bits */
    __shared__ uint8_t s_perm[G::THREADS];

    const uint32_t img_i = a.img0 + blockIdx.z;
    const JDImageDesc &im = a.imgs[img_i];
    const uint32_t strip = blockIdx.x, my = blockIdx.y;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint16_t *const irec = a.rec + im.rec_base;

    /* ---- headers, quant, binning: bin 0 = rows 4-7 empty & columns 0-3, 1 = rows 4-7 populated & columns 0-3,
     * 2 = populated & beyond column 3, 3 = empty & beyond column 3 (one boundary between the 2-pair and the 4-pair blocks,
     * two between the column-pass variants) ---- */
    uint32_t key = 4;
    if (tid < (uint32_t)G::NB) {
        const uint32_t ml = jd_div_small<G::BPMEFF>(tid), blk = tid - ml * G::BPMEFF;
        const uint32_t mx = strip * MPB + ml;
        if (mx < a.mcus_x) {
            const jd_u64 h = __ldg(a.blk_hdr + im.blk_base + (my * a.mcus_x + mx) * a.bpm + blk);
            s_hdr[tid] = h;
            const uint32_t wide = (JD_HDR_COLMASK(h) & 0xF0u) != 0u, hi = JD_HDR_HI(h);
            key = wide ? (hi ? 2u : 3u) : hi;
        }
    }
    for (uint32_t i = tid; i < (NC == 3 ? 192u : 64u); i += G::THREADS) {
        const uint32_t n = i & 63u;
        s_q[i] = (uint16_t)__ldg(a.quant + (size_t)img_i * 192 + (i & ~63u) + (n & 7u) * 8u + (n >> 3));   /* stored column-major */
    }
    uint32_t bal[4];
#pragma unroll
    for (int k = 0; k < 4; k++) bal[k] = __ballot_sync(0xffffffffu, key == (uint32_t)k);
    if (lane < 4u) s_wc[lane][wid] = __popc(lane == 0u ? bal[0] : lane == 1u ? bal[1] : lane == 2u ? bal[2] : bal[3]);
    __syncthreads();
    uint32_t nbin[4], pre = 0;
    {
        uint32_t before = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 c = *reinterpret_cast<const uint4 *>(s_wc[k]);
            nbin[k] = c.x + c.y + c.z + c.w;
            const uint32_t inwarps = (wid > 0u ? c.x : 0u) + (wid > 1u ? c.y : 0u) + (wid > 2u ? c.z : 0u);
            if (key == (uint32_t)k) pre = before + inwarps + __popc(bal[k] & ((1u << lane) - 1u));
            before += nbin[k];
        }
    }
    if (key < 4u) s_perm[pre] = (uint8_t)tid;
    __syncthreads();
    const uint32_t n2 = nbin[0] + nbin[1], nall = n2 + nbin[2] + nbin[3];

    /* ---- this thread's block ---- */
    const bool have = tid < nall;
    const bool wide = have && tid >= n2;
    const bool wide_any = __any_sync(0xffffffffu, wide);
    if (have) {
        const uint32_t pb = s_perm[tid];
        const jd_u64 h = s_hdr[pb];
        const uint32_t ml = jd_div_small<G::BPMEFF>(pb), blk = pb - ml * G::BPMEFF;
        const uint32_t comp = (blk < (uint32_t)(HS * VS)) ? 0u : blk - HS * VS + 1u;
        const uint32_t ri = JD_HDR_REC(h), ncoef = JD_HDR_NCOEF(h);
        const bool hi = JD_HDR_HI(h) != 0u;
        const uint16_t *q = s_q + comp * 64u;
        uint32_t *tile = s_tile + tid * G::TWORDS;
        uint16_t *t16 = reinterpret_cast<uint16_t *>(tile);
        uint8_t *prow;  /* first output row of this block in the staged plane */
        uint32_t pstride;
        if (comp == 0) {
            const uint32_t lx = (HS == 2) ? (blk & 1u) : 0u;
            const uint32_t ly = (HS == 2 && VS == 2) ? (blk >> 1) : ((VS == 2) ? blk : 0u);
            prow = s_y + (ly * 8) * G::YSTRIDE + (ml * HS + lx) * 8; pstride = G::YSTRIDE;
        } else {
            prow = s_c + ((comp - 1) * 8) * G::CSTRIDE + ml * 8; pstride = G::CSTRIDE;
        }
        /* expand + dequantise: d = (int16)(coefficient * quant) like _mm_mullo_epi16 (jpeg.inl:2338) */
        const uint4 z4 = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int r = 0; r < 4; r++) *reinterpret_cast<uint4 *>(tile + 4 * r) = z4;
        if (hi) {
#pragma unroll
            for (int r = 4; r < 8; r++) *reinterpret_cast<uint4 *>(tile + 4 * r) = z4;
        }
        if (!JD_HDR_BIG(h)) {
            /* the first 10 halfwords that cover the records come in as five independent aligned 32-bit loads (one round
             * trip instead of a chain of 2-byte loads); longer blocks finish in the loop below */
            const uint32_t off = ri & 1u, total = off + ncoef;
            const uint32_t *w32 = reinterpret_cast<const uint32_t *>(irec + (ri - off));
            uint32_t v[5];
#pragma unroll
            for (int w = 0; w < 5; w++) v[w] = ((uint32_t)(2 * w) < total) ? __ldg(w32 + w) : 0u;
#pragma unroll
            for (int hh = 0; hh < 10; hh++) {
                if ((uint32_t)hh >= off && (uint32_t)hh < total) {
                    const uint32_t r = (hh & 1) ? (v[hh >> 1] >> 16) : (v[hh >> 1] & 0xFFFFu);
                    const uint32_t n = JD_TRANSPOSE6(r >> 10);      /* records carry column-major positions */
                    t16[n] = (uint16_t)(((int)(r << 22) >> 22) * (int)q[n]);
                }
            }
            for (uint32_t i = 10u - off; i < ncoef; i++) {
                const uint32_t r = __ldg(irec + ri + i), n = JD_TRANSPOSE6(r >> 10);
                t16[n] = (uint16_t)(((int)(r << 22) >> 22) * (int)q[n]);
            }
        } else {
            for (uint32_t i = 0; i < ncoef; i++) {
                const uint32_t t = __ldg(irec + ri + 2 * i) & 63u, n = JD_TRANSPOSE6(t);
                t16[n] = (uint16_t)((int)(short)__ldg(irec + ri + 2 * i + 1) * (int)q[n]);
            }
        }
        t16[0] = (uint16_t)(JD_HDR_DC(h) * (int)q[0] + JD_ROW_BIAS);
        const uint32_t cm = JD_HDR_COLMASK(h);
        if (!wide_any) {
            uint32_t x[8][2];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                if (r < 4 || hi) { const uint2 w = *reinterpret_cast<const uint2 *>(tile + 4 * r); x[r][0] = w.x; x[r][1] = w.y; }
                else { x[r][0] = 0u; x[r][1] = 0u; }
            }
            jd_idct_block_packed<2>(x, hi, cm, prow, pstride);
        } else {
            uint32_t x[8][4];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                if (r < 4 || hi) { const uint4 w = *reinterpret_cast<const uint4 *>(tile + 4 * r); x[r][0] = w.x; x[r][1] = w.y; x[r][2] = w.z; x[r][3] = w.w; }
                else { x[r][0] = 0u; x[r][1] = 0u; x[r][2] = 0u; x[r][3] = 0u; }
            }
            jd_idct_block_packed<4>(x, hi, cm, prow, pstride);
        }
    }
    __syncthreads();

    /* ---- phase C ---- */
    const uint32_t W = a.padded ? a.mcus_x * HS * 8 : a.width;
    const uint32_t H = a.padded ? a.mcus_y * VS * 8 : a.height;
    uint8_t *outbase = a.out + im.out_off;
    const uint32_t pitch = im.out_pitch;
    const uint8_t *s_cb = s_c, *s_cr = s_c + 8 * G::CSTRIDE;
    if (HALF) {
        jd_phase_c_half<HS, VS, NC, PT, G::WCTA, G::HCTA, G::YSTRIDE, G::CSTRIDE, G::THREADS>(a, s_y, s_cb, s_cr, strip, my, tid, W, H, outbase, pitch);
    } else if ((strip + 1) * G::WCTA <= W && (my + 1) * G::HCTA <= H && ((reinterpret_cast<uintptr_t>(outbase) | pitch) & 15u) == 0u) {
        jd_phase_c_full<HS, VS, NC, PT, JPEG_ARITH_SSE2, G::WCTA, G::YSTRIDE, G::CSTRIDE, G::THREADS, true>(a, s_y, s_cb, s_cr, strip, my, tid, W, H, outbase, pitch);
    } else {
        jd_phase_c_full<HS, VS, NC, PT, JPEG_ARITH_SSE2, G::WCTA, G::YSTRIDE, G::CSTRIDE, G::THREADS, false>(a, s_y, s_cb, s_cr, strip, my, tid, W, H, outbase, pitch);
    }
}

/* ------------------------------------------------------------------------------------ */
/* 1/4 and 1/8 scale: one thread per MCU                                                   */
/* ------------------------------------------------------------------------------------ */
struct JDScaledArgs {
    const JDImageDesc *imgs;
    const jd_u64 *blk_hdr;
    const uint16_t *rec;
    const int32_t *quant;
    uint8_t *out;
    uint32_t img0;
    uint32_t pixel_type;   /* JPEGDEC.h pixel type (after LUMA_ONLY folding) */
    uint32_t eighth;       /* 1: 1/8, 0: 1/4 */
    uint32_t padded;
};

__device__ __forceinline__ void jd_scaled_block(const uint16_t *irec, jd_u64 h, const int32_t *q, bool eighth, uint32_t px[4])
{
    const int dc = JD_HDR_DC(h);
    const int q0 = q[0];
    if (eighth) { px[0] = jd_range(dc * q0); return; }
    const uint32_t ri = JD_HDR_REC(h), ncoef = JD_HDR_NCOEF(h), big = JD_HDR_BIG(h);
    int m1 = 0, m8 = 0, m9 = 0;
    bool any = false;
    /* records are in zigzag order: the ones the 1/4 path keeps (zigzag 1..4 = natural 1, 8, 16, 9 =
     * tile positions 8, 1, 2, 9; jpeg.inl:2117-2119) come first */
    for (uint32_t i = 0; i < ncoef; i++) {
        uint32_t t; int v;
        if (big) { t = irec[ri + 2 * i] & 63u; v = (int)(short)irec[ri + 2 * i + 1]; }
        else { const uint32_t r = irec[ri + i]; t = r >> 10; v = (int)(r << 22) >> 22; }
        if (t == 8u) m1 = v; else if (t == 1u) m8 = v; else if (t == 9u) m9 = v; else if (t != 2u) break;
        any = true;
    }
    if (!any) { px[0] = px[1] = px[2] = px[3] = jd_range(dc * q0); return; }
    /* 2x2 butterfly (jpeg.inl:2305-2326); q is column-major: natural 1 -> [8], natural 8 -> [1], natural 9 -> [9] */
    int t4 = dc * q0, t5 = m8 * q[1];
    const int t0 = t4 + t5, t2 = t4 - t5;
    t4 = m1 * q[8]; t5 = m9 * q[9];
    const int t1 = t4 + t5, t3 = t4 - t5;
    px[0] = jd_range(t0 + t1); px[1] = jd_range(t0 - t1); px[2] = jd_range(t2 + t3); px[3] = jd_range(t2 - t3);
}

__global__ void __launch_bounds__(128) jdk_scaled(const JDScaledArgs a)
{
    const uint32_t img_i = a.img0 + blockIdx.y;
    const JDImageDesc &im = a.imgs[img_i];
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (uint32_t)im.mcus_x * im.mcus_y) return;
    const uint32_t mx = m % im.mcus_x, my = m / im.mcus_x;
    const uint32_t hs = (im.subsample >> 4) ? (im.subsample >> 4) : 1, vs = (im.subsample & 15) ? (im.subsample & 15) : 1;
    const uint32_t nluma = hs * vs;
    const bool eighth = a.eighth != 0;
    const uint32_t bs = eighth ? 1u : 2u; /* block edge in output pixels */
    const int32_t *q = a.quant + (size_t)img_i * 192;
    const jd_u64 *hdr = a.blk_hdr + im.blk_base + (size_t)m * im.bpm;
    uint32_t ypx[4][4], cb[4], cr[4];
    const uint16_t *const irec = a.rec + im.rec_base;
    for (uint32_t b = 0; b < nluma; b++) jd_scaled_block(irec, hdr[b], q, eighth, ypx[b]);
    const bool gray_out = a.pixel_type >= EIGHT_BIT_GRAYSCALE;
    const bool colour = (im.ncomp == 3) && !gray_out;
    if (colour) {
        jd_scaled_block(irec, hdr[nluma], q + 64, eighth, cb);
        jd_scaled_block(irec, hdr[nluma + 1], q + 128, eighth, cr);
    }
    const uint32_t shift = eighth ? 3u : 2u;
    const uint32_t W = a.padded ? (uint32_t)im.mcus_x * hs * bs : (((uint32_t)im.width + (1u << shift) - 1u) >> shift);
    const uint32_t H = a.padded ? (uint32_t)im.mcus_y * vs * bs : (((uint32_t)im.height + (1u << shift) - 1u) >> shift);
    uint8_t *outbase = a.out + im.out_off;
    const uint32_t ow = hs * bs, oh = vs * bs; /* output pixels per MCU */
    for (uint32_t y = 0; y < oh; y++) {
        for (uint32_t x = 0; x < ow; x++) {
            const uint32_t gx = mx * ow + x, gy = my * oh + y;
            if (gx >= W || gy >= H) continue;
            const uint32_t bx = x / bs, by = y / bs;
            const uint32_t lb = (hs == 2 && vs == 2) ? by * 2 + bx : (hs == 2 ? bx : by);
            const uint32_t Y = ypx[lb][(y % bs) * bs + (x % bs)];
            if (gray_out) { outbase[(size_t)gy * im.out_pitch + gx] = (uint8_t)Y; continue; }
            uint8_t *dst = outbase + (size_t)gy * im.out_pitch;
            if (im.ncomp == 1) {
                uint32_t v = jd_gray565(Y);
                if (a.pixel_type != RGB565_LITTLE_ENDIAN) v = jd_bswap16(v);
                reinterpret_cast<uint16_t *>(dst)[gx] = (uint16_t)v;
                continue;
            }
            const uint32_t ci = (y / vs) * bs + (x / hs); /* nearest chroma byte of the 2x2 (or 1) chroma block */
            if (a.pixel_type == RGB8888) reinterpret_cast<uint32_t *>(dst)[gx] = jd_rgb8888_scalar((int)Y << 12, (int)cb[ci], (int)cr[ci]);
            else {
                uint32_t v = jd_rgb565_scalar((int)Y << 12, (int)cb[ci], (int)cr[ci]);
                if (a.pixel_type == RGB565_BIG_ENDIAN) v = jd_bswap16(v);
                reinterpret_cast<uint16_t *>(dst)[gx] = (uint16_t)v;
            }
        }
    }
}
