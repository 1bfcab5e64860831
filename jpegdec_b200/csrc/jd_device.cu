/*
 * jd_device.cu -- batch decode pipeline on one B200: buffers, uploads, kernel launches,
 * CUDA-event stage timings, downloads.  Exposes the JPEGB200_* C ABI (include/jpegdec_b200.h).
 * There is no CPU fallback anywhere in this file: if CUDA is unavailable every entry point fails.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <time.h>
#include <new>
#include <sched.h>
#include <unistd.h>
#include <sys/syscall.h>

#include "jd_kernels.cuh"

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(ctx_err(), 256, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 0;                                                                              \
        }                                                                                          \
    } while (0)

/* error text of the calling thread (two host threads driving two contexts never share it) */
static thread_local char g_err[256];

/* Device allocations are recycled through the context: a decode job borrows buffers and hands them
 * back on destroy, so steady-state batches do no cudaMalloc/cudaFree (both synchronise the device). */
struct JDPool {
    struct Slot { void *p; size_t bytes; };
    std::vector<Slot> free_;
    cudaError_t get(size_t bytes, void **out, size_t *got)
    {
        int best = -1;
        for (size_t i = 0; i < free_.size(); i++)
            if (free_[i].bytes >= bytes && free_[i].bytes <= 2 * bytes + (1u << 20) && (best < 0 || free_[i].bytes < free_[best].bytes)) best = (int)i;
        if (best >= 0) { *out = free_[best].p; *got = free_[best].bytes; free_.erase(free_.begin() + best); return cudaSuccess; }
        cudaError_t e = cudaMalloc(out, bytes);
        if (e != cudaSuccess) { /* give cached memory back to the driver and retry once */
            cudaGetLastError();
            drain();
            e = cudaMalloc(out, bytes);
        }
        *got = bytes;
        return e;
    }
    void put(void *p, size_t bytes) { free_.push_back(Slot{p, bytes}); }
    void drain() { for (auto &s : free_) cudaFree(s.p); free_.clear(); }
};

/* Pinned host staging blocks (status read-back) recycled the same way: a D2H copy into pageable memory would make
 * batchDownload wait for the whole job, which is what keeps several jobs from being in flight from one host thread. */
struct JDPinPool {
    struct Slot { void *p; size_t bytes; };
    std::vector<Slot> free_;
    void *get(size_t bytes, size_t *got)
    {
        for (size_t i = 0; i < free_.size(); i++)
            if (free_[i].bytes >= bytes) { void *q = free_[i].p; *got = free_[i].bytes; free_.erase(free_.begin() + i); return q; }
        void *q = nullptr;
        size_t need = (bytes + 4095) & ~(size_t)4095;
        if (cudaHostAlloc(&q, need, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        *got = need;
        return q;
    }
    void put(void *p, size_t bytes) { free_.push_back(Slot{p, bytes}); }
    void drain() { for (auto &s : free_) cudaFreeHost(s.p); free_.clear(); }
};

/* a job's stream and its timing events, recycled through the context (creating them costs more than a small decode) */
struct JDStreamSet {
    cudaStream_t stream;
    cudaEvent_t ev[JPEGB200_NUM_TIMINGS + 2];
};

struct JPEGB200_CTX {
    std::vector<JDStreamSet> free_streams;
    int device;
    int arith;
    char err[256];
    bool has_shared;
    uint64_t shared_hash, shared_hash2;
    uint16_t shared_lut[JD_LUT_ENTRIES];
    int shared_hits;
    JDPool pool;
    JDPinPool pinpool;
    int64_t last_counters[JPEGB200_NUM_COUNTERS]; /* summed over the jobs of the last JPEGB200_decodeBatch */
    float last_ms[JPEGB200_NUM_TIMINGS];          /* CUDA-event stage times summed over those jobs */
    int last_jobs;
    int numa_node;                                /* host NUMA node the GPU hangs off (-1 unknown) */
    int pipe_depth;                               /* jobs in flight inside JPEGB200_decodeBatch (0 = default) */
};

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    size_t bytes = 0;
    JDPool *pool = nullptr;
    cudaError_t alloc(JDPool *from, size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        release();
        pool = from;
        size_t need = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
        void *q = nullptr;
        cudaError_t e = pool ? pool->get(need, &q, &bytes) : cudaMalloc(&q, bytes = need);
        if (e == cudaSuccess) { p = (T *)q; n = count; }
        return e;
    }
    void release()
    {
        if (p) { if (pool) pool->put(p, bytes); else cudaFree(p); }
        p = nullptr; n = 0; bytes = 0;
    }
};

struct JPEGB200_BATCH {
    JPEGB200_CTX *ctx;
    int n;
    int pixel_type, options, sshift, ptclass, dither_bits;
    bool gray_out;
    bool padded; /* write the whole MCU-aligned frame (single-image API: callbacks deliver whole MCUs) */
    cudaStream_t stream;
    std::vector<JDInfo> infos;
    std::vector<int32_t> parse_status;
    std::vector<const uint8_t *> datas;
    std::vector<int32_t> sizes;
    std::vector<JDImageDesc> descs;
    std::vector<int32_t> quant;
    std::vector<uint16_t> luts;
    std::vector<uint32_t> work, cta_lut, seg_img;
    std::vector<uint64_t> comp_off; /* offset of each file in the device blob */
    std::vector<void *> outs;
    std::vector<int64_t> pitches;
    std::vector<uint16_t> errinit;  /* dither: initial error line per image (reference quirk), value | 0xFF00 (tag of "the band above band 0") */
    size_t comp_total, out_total, gray_total;
    uint32_t nseg, nlut;
    uint64_t nblk;
    bool contiguous_in;
    bool uploaded, out_device, arena_owned;
    DevBuf<uint8_t> d_comp, d_out, d_gray;
    DevBuf<uint16_t> d_errline;
    DevBuf<uint64_t> d_gray_off; /* [0,n): gray-stage offsets, [n,2n): packed output offsets */
    DevBuf<uint32_t> d_err_off, d_dprog;
    DevBuf<uint8_t> d_clean;       /* un-stuffed restart segments (jdk_unstuff_segs) */
    DevBuf<uint32_t> d_seg_clen;
    uint64_t rec_total;            /* coefficient records the batch may need (JD_REC_INDEX layout) */
    DevBuf<uint4> d_dbands;        /* dither: (image, band, list position of the band above, -) per warp */
    std::vector<uint4> dbands;
    JDImageDesc *descs_dl;             /* descriptors read back (status, err_mcu); pinned, from ctx->pinpool */
    size_t descs_dl_bytes;
    bool downloaded;
    DevBuf<JDImageDesc> d_descs;
    DevBuf<int32_t> d_quant;
    DevBuf<uint16_t> d_luts, d_rec;
    DevBuf<uint32_t> d_work, d_cta_lut, d_seg_img, d_seg_start, d_seg_jmap, d_seg_status, d_seg_nrec, d_seg_phase, d_counters;
    DevBuf<jd_u64> d_blk_hdr;
    DevBuf<JDEvent> d_events;
    std::vector<uint64_t> arena_off; /* per-image offset inside d_out */
    /* restart-free scans decoded chunk-parallel (jd_chunk.h) */
    std::vector<uint32_t> cimg_list;
    uint32_t nchunks, max_nch;
    DevBuf<uint8_t> d_filt;
    DevBuf<uint32_t> d_cimg_list, d_flen, d_E0, d_E1, d_Ep, d_cfirst, d_cn, d_cpre, d_cjmap, d_cstatus, d_cnown;
    DevBuf<int32_t> d_cdcs, d_cpe;
    uint32_t h_changed;
    bool chunk_iterate;            /* restart-free scans: iterate the entry states with a host check (fallback mode) */
    int decode_flags;
    cudaEvent_t ev[JPEGB200_NUM_TIMINGS + 2];
    bool have_ev;
    float ms[JPEGB200_NUM_TIMINGS];
    int64_t counters[JPEGB200_NUM_COUNTERS];
    uint32_t *h_counters;              /* 8 words at the end of the descs_dl block */
};

static char *ctx_err() { return g_err; }

#define JD_EVENT_CAP (1u << 20)
#define JD_CHUNK_PASSES 6     /* restart-free scans: entry-state passes per decode (from the third on only moved chunks are parsed; the last one verifies) */

extern "C" int JPEGB200_deviceCount(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

/* the calling thread's current CUDA device (-1 without CUDA) */
extern "C" int JPEGB200_currentDevice(void)
{
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
    return d;
}

extern "C" JPEGB200_CTX *JPEGB200_create(int device, int arith_mode)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        snprintf(g_err, sizeof(g_err), "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n) { snprintf(g_err, sizeof(g_err), "device %d out of range (%d devices)", device, n); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaSetDevice(%d) failed", device); return nullptr; }
    /* make sure the sm_100a kernel image is loadable on this GPU */
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, jdk_prescan);
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "kernel image not loadable on device %d: %s (built for sm_100a only)", device, cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    JPEGB200_CTX *c = new (std::nothrow) JPEGB200_CTX();
    if (!c) return nullptr;
    c->device = device;
    c->arith = arith_mode ? JPEG_ARITH_SCALAR : JPEG_ARITH_SSE2;
    c->err[0] = 0;
    c->has_shared = false;
    c->shared_hits = 0;
    memset(c->last_counters, 0, sizeof(c->last_counters));
    memset(c->last_ms, 0, sizeof(c->last_ms));
    c->last_jobs = 0;
    c->pipe_depth = 0;
    c->numa_node = -1;
    {   /* which host NUMA node is this GPU attached to?  (/sys/bus/pci/devices/<domain:bus:dev.fn>/numa_node) */
        char bus[32] = {0}, path[96];
        if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) == cudaSuccess) {
            for (char *q = bus; *q; q++) if (*q >= 'A' && *q <= 'Z') *q = (char)(*q - 'A' + 'a');
            snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
            FILE *f = fopen(path, "r");
            if (f) { int node = -1; if (fscanf(f, "%d", &node) == 1) c->numa_node = node; fclose(f); }
        } else cudaGetLastError();
    }
    return c;
}

extern "C" void JPEGB200_destroy(JPEGB200_CTX *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->pool.drain();
    ctx->pinpool.drain();
    for (auto &ss : ctx->free_streams) { for (auto &e : ss.ev) cudaEventDestroy(e); cudaStreamDestroy(ss.stream); }
    delete ctx;
}

extern "C" const char *JPEGB200_lastErrorString(JPEGB200_CTX *) { return g_err; }

extern "C" void *JPEGB200_hostAlloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

extern "C" void JPEGB200_hostFree(void *p) { if (p) cudaFreeHost(p); }

/* ---- host placement: the pixels of a batch leave the GPU over PCIe into pinned host memory; on a two-socket box that
 * memory (and the thread that drives the copies) should sit on the socket the GPU hangs off, or every byte crosses the
 * inter-socket link as well.  Nothing here is required for correctness. ---- */
extern "C" int JPEGB200_numaNode(JPEGB200_CTX *ctx) { return ctx ? ctx->numa_node : -1; }

/* parse "0-3,8,10-11" */
static int jd_parse_cpulist(const char *txt, cpu_set_t *set)
{
    int n = 0;
    CPU_ZERO(set);
    const char *q = txt;
    while (*q) {
        char *e;
        long a = strtol(q, &e, 10);
        if (e == q) break;
        long b2 = a;
        if (*e == '-') { q = e + 1; b2 = strtol(q, &e, 10); }
        for (long c = a; c <= b2 && c < CPU_SETSIZE; c++) { CPU_SET((int)c, set); n++; }
        q = e;
        while (*q == ',' || *q == ' ' || *q == '\n') q++;
    }
    return n;
}

/* Pins the CALLING thread (and the threads it creates later) to the CPUs of the context's NUMA node -- intersected with
 * the CPUs the process may use -- and makes that node the preferred one for its allocations; pinned memory allocated
 * afterwards (JPEGB200_hostAlloc) lands there.  Returns the number of CPUs in the new mask, 0 if nothing was changed. */
extern "C" int JPEGB200_bindHostToDevice(JPEGB200_CTX *ctx)
{
    if (!ctx || ctx->numa_node < 0) return 0;
    char path[96], txt[4096];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", ctx->numa_node);
    FILE *f = fopen(path, "r");
    if (!f) return 0;
    const size_t got = fread(txt, 1, sizeof(txt) - 1, f);
    fclose(f);
    txt[got] = 0;
    cpu_set_t node, cur, both;
    if (jd_parse_cpulist(txt, &node) == 0) return 0;
    if (sched_getaffinity(0, sizeof(cur), &cur) != 0) return 0;
    CPU_AND(&both, &node, &cur);
    const int n = CPU_COUNT(&both);
    if (n == 0) return 0;
    if (sched_setaffinity(0, sizeof(both), &both) != 0) return 0;
#ifdef SYS_set_mempolicy
    {   /* MPOL_PREFERRED = 1: fall back to other nodes rather than fail when the node is full */
        unsigned long mask[16] = {0};
        if (ctx->numa_node < (int)(8 * sizeof(mask))) {
            mask[ctx->numa_node / (8 * sizeof(unsigned long))] |= 1ul << (ctx->numa_node % (8 * sizeof(unsigned long)));
            (void)syscall(SYS_set_mempolicy, 1, mask, (unsigned long)(8 * sizeof(mask)));
        }
    }
#endif
    return n;
}

extern "C" void *JPEGB200_deviceAlloc(JPEGB200_CTX *ctx, size_t bytes)
{
    if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

extern "C" void JPEGB200_deviceFree(JPEGB200_CTX *ctx, void *p)
{
    if (ctx) cudaSetDevice(ctx->device);
    if (p) cudaFree(p);
}

extern "C" int JPEGB200_deviceRead(JPEGB200_CTX *ctx, void *host_dst, const void *dev_src, size_t bytes)
{
    if (!ctx) return 0;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost));
    return 1;
}

/* ---- digests of device-resident pixels: lets a caller (bench / tests) verify a whole device-resident batch against
 * reference digests without moving the pixels to the host.  digest = sum over 8-byte words i of mix64(word ^ i * K)
 * mod 2^64 (mix64 = the splitmix64 finaliser); order independent, so it reduces in parallel. ---- */
__device__ __forceinline__ unsigned long long jd_mix64(unsigned long long z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) jdk_digest(const uint8_t *const *ptrs, const int64_t *lens, unsigned long long *out)
{
    const uint32_t img = blockIdx.y;
    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(ptrs[img]);
    const int64_t nbytes = lens[img], nfull = nbytes >> 3;
    unsigned long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nfull; i += (int64_t)gridDim.x * blockDim.x)
        acc += jd_mix64(w[i] ^ ((unsigned long long)i * 0x9E3779B97F4A7C15ull));
    if (blockIdx.x == 0 && threadIdx.x == 0 && (nbytes & 7)) {   /* tail bytes, zero padded */
        unsigned long long t = 0;
        const uint8_t *q = ptrs[img] + (nfull << 3);
        for (int k = 0; k < (int)(nbytes & 7); k++) t |= (unsigned long long)q[k] << (8 * k);
        acc += jd_mix64(t ^ ((unsigned long long)nfull * 0x9E3779B97F4A7C15ull));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    __shared__ unsigned long long s_part[8];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < 8; k++) t += s_part[k];
        atomicAdd(out + img, t);
    }
}

/* digests of n device byte ranges (8-byte aligned starts); synchronous */
extern "C" int JPEGB200_digestDevice(JPEGB200_CTX *ctx, const void *const *dev_ptrs, const int64_t *lengths, int n, uint64_t *digests)
{
    if (!ctx || n <= 0 || !dev_ptrs || !lengths || !digests) return 0;
    for (int i = 0; i < n; i++) if (((uintptr_t)dev_ptrs[i] & 7u) || lengths[i] < 0) { snprintf(g_err, sizeof(g_err), "digest ranges must start 8-byte aligned"); return 0; }
    CK(cudaSetDevice(ctx->device));
    void *d = nullptr;
    const size_t pb = (size_t)n * 8;
    CK(cudaMalloc(&d, 3 * pb));
    uint8_t *dp = (uint8_t *)d;
    int ok = 1;
    if (cudaMemcpy(dp, dev_ptrs, pb, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dp + pb, lengths, pb, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(dp + 2 * pb, 0, pb) != cudaSuccess) ok = 0;
    if (ok) {
        for (int i0 = 0; i0 < n; i0 += 32768) {
            const int cnt = (n - i0 < 32768) ? n - i0 : 32768;
            jdk_digest<<<dim3(64, cnt), 256>>>((const uint8_t *const *)dp + i0, (const int64_t *)(dp + pb) + i0, (unsigned long long *)(dp + 2 * pb) + i0);
        }
        if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(digests, dp + 2 * pb, pb, cudaMemcpyDeviceToHost) != cudaSuccess) ok = 0;
    }
    if (!ok) snprintf(g_err, sizeof(g_err), "digest failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
    return ok;
}

extern "C" int JPEGB200_setPipelineDepth(JPEGB200_CTX *ctx, int jobs_in_flight)
{
    if (!ctx || jobs_in_flight < 0 || jobs_in_flight > 16) return 0;
    ctx->pipe_depth = jobs_in_flight;
    return 1;
}

/* ---- shared table blob ---- */
extern "C" int JPEGB200_exportTables(const uint8_t *jpeg, int size, uint8_t *blob)
{
    JDInfo *info = new JDInfo();
    int rc = jd_parse_header(jpeg, size, 0, info);
    if (rc) {
        const uint64_t h = jd_tables_hash(info), h2 = jd_tables_hash2(info);   /* 128-bit key of the DHT content */
        memcpy(blob, &h, 8);
        memcpy(blob + 8, &h2, 8);
        jd_build_lut(info, (uint16_t *)(blob + 16));
        jd_build_quant(info, (int16_t *)(blob + 16 + JD_LUT_ENTRIES * 2));
    }
    delete info;
    return rc;
}

extern "C" int JPEGB200_setSharedTables(JPEGB200_CTX *ctx, const uint8_t *blob)
{
    if (!ctx) return 0;
    memcpy(&ctx->shared_hash, blob, 8);
    memcpy(&ctx->shared_hash2, blob + 8, 8);
    memcpy(ctx->shared_lut, blob + 16, JD_LUT_ENTRIES * 2);
    ctx->has_shared = true;
    ctx->shared_hits = 0;
    return 1;
}

extern "C" int JPEGB200_sharedTableHits(JPEGB200_CTX *ctx) { return ctx ? ctx->shared_hits : 0; }

/* ---- batch ---- */
static int bytes_per_pixel_class(int ptclass) { return ptclass == JD_PT_565 ? 2 : (ptclass == JD_PT_8888 ? 4 : 1); }

extern "C" JPEGB200_BATCH *JPEGB200_batchCreate(JPEGB200_CTX *ctx, const uint8_t *const *datas, const int32_t *sizes,
                                                int n, int pixel_type, int options)
{
    if (!ctx || n <= 0 || pixel_type < 0 || pixel_type >= INVALID_PIXEL_TYPE) { snprintf(g_err, sizeof(g_err), "invalid parameter"); return nullptr; }
    JPEGB200_BATCH *b = new (std::nothrow) JPEGB200_BATCH();
    if (!b) return nullptr;
    b->ctx = ctx;
    b->n = n;
    if ((options & JPEG_LUMA_ONLY) && pixel_type < EIGHT_BIT_GRAYSCALE) pixel_type = EIGHT_BIT_GRAYSCALE; /* jpeg.inl:4991 */
    b->pixel_type = pixel_type;
    b->padded = (options & 0x10000) != 0; /* JPEGB200_OPT_PADDED (internal, jd_api.c) */
    b->options = options;
    b->sshift = (options & JPEG_SCALE_HALF) ? 1 : (options & JPEG_SCALE_QUARTER) ? 2 : (options & JPEG_SCALE_EIGHTH) ? 3 : 0;
    b->gray_out = pixel_type >= EIGHT_BIT_GRAYSCALE;
    b->ptclass = (pixel_type == RGB8888) ? JD_PT_8888 : (b->gray_out ? JD_PT_GRAY : JD_PT_565);
    b->dither_bits = (pixel_type == FOUR_BIT_DITHERED) ? 4 : (pixel_type == TWO_BIT_DITHERED) ? 2 : (pixel_type == ONE_BIT_DITHERED) ? 1 : 0;
    b->stream = nullptr;
    b->descs_dl = nullptr; b->descs_dl_bytes = 0; b->downloaded = false; b->h_counters = nullptr;
    b->nchunks = 0; b->max_nch = 0; b->chunk_iterate = false; b->decode_flags = 0;
    b->uploaded = false; b->out_device = false; b->arena_owned = false; b->have_ev = false;
    memset(b->ms, 0, sizeof(b->ms));
    memset(b->counters, 0, sizeof(b->counters));
    b->infos.resize(n);
    b->parse_status.assign(n, JPEG_SUCCESS);
    b->datas.assign(datas, datas + n);
    b->sizes.assign(sizes, sizes + n);
    b->descs.resize(n);
    b->quant.assign((size_t)n * 192, 0);
    b->outs.assign(n, nullptr);
    b->pitches.assign(n, 0);
    b->comp_off.assign(n, 0);
    b->arena_off.assign(n, 0);

    /* input layout: one span if the files already sit back to back in host memory */
    bool contig = true;
    for (int i = 1; i < n && contig; i++) {
        const uint8_t *prev_end = datas[i - 1] + sizes[i - 1];
        if (datas[i] < prev_end || (size_t)(datas[i] - prev_end) > 4096) contig = false;
    }
    b->contiguous_in = contig;
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        if (contig) off = (size_t)(datas[i] - datas[0]);
        b->comp_off[i] = off;
        if (!contig) off += ((size_t)sizes[i] + 15) & ~(size_t)15;
    }
    b->comp_total = contig ? (size_t)(datas[n - 1] + sizes[n - 1] - datas[0]) : off;
    if (b->comp_total >= (3ull << 30)) {
        /* byte offsets into the batch buffer (and into the un-stuffed copy, which adds 32 bytes per segment) are 32-bit */
        snprintf(g_err, sizeof(g_err), "batch holds %zu compressed bytes; one job takes at most 3 GiB (JPEGB200_decodeBatch splits larger batches)", b->comp_total);
        delete b;
        return nullptr;
    }

    std::vector<uint64_t> lut_hash;
    std::vector<int> lut_owner;      /* first image that defined each LUT set: a hash match is confirmed on the DHT bytes */
    uint32_t seg = 0;
    uint64_t blk = 0, rec_total = 0;
    size_t out_total = 0, gray_total = 0;
    for (int i = 0; i < n; i++) {
        JDInfo &inf = b->infos[i];
        JDImageDesc &d = b->descs[i];
        memset(&d, 0, sizeof(d));
        int ok = jd_parse_header(datas[i], sizes[i], 0, &inf);
        int st = ok ? JPEG_SUCCESS : inf.error;
        if (ok && (options & JPEG_EXIF_THUMBNAIL)) {
            if (inf.thumb_data == 0 || inf.thumb_w == 0) { ok = 0; st = JPEG_INVALID_PARAMETER; }
            else { ok = jd_parse_header(datas[i], sizes[i], inf.thumb_data, &inf); if (!ok) st = inf.error; }
        }
        bool prog = false;
        if (ok && inf.mode == 0xC2) {
            /* progressive: like the reference, only the DC coefficients of the first scan are decoded and a 1/8-size image
             * is produced (jpeg.inl:4964-4966, JPEGDecodeMCU_P :1819-1884).  That needs a first scan that is the interleaved
             * DC scan of every component (Ss = Se = 0, Ah = 0) -- what every common encoder writes -- and 1/8 scale. */
            prog = true;
            if (b->sshift != 3 || inf.p.ncomp_in_scan != inf.ncomp || inf.p.scan_start != 0 || inf.p.scan_end != 0 || (inf.approx >> 4) != 0 ||
                (inf.approx & 15) > 13) { ok = 0; st = JPEG_UNSUPPORTED_FEATURE; }
        } else if (ok && inf.mode != 0xC0) { ok = 0; st = JPEG_UNSUPPORTED_FEATURE; }
        if (ok && !inf.tables_ok) { ok = 0; st = JPEG_DECODE_ERROR; }           /* jpeg.inl:2166 */
        if (ok && inf.ncomp == 1 && pixel_type == RGB8888) { ok = 0; st = JPEG_INVALID_PARAMETER; }
        if (ok && (uint64_t)sizes[i] >= (512ull << 20)) { ok = 0; st = JPEG_UNSUPPORTED_FEATURE; }   /* image-relative record indices are 32-bit */
        b->parse_status[i] = st;
        if (!ok) { /* keep a harmless empty descriptor */
            d.nseg = 0; d.seg_base = seg; d.blk_base = (uint32_t)blk; d.status = (uint32_t)st;
            continue;
        }
        {   /* kernels read quant column-major ([c * 8 + r]) so a lane's column is one 16-byte load */
            int16_t qn[192];
            jd_build_quant(&inf, qn);
            int32_t *qt = &b->quant[(size_t)i * 192];
            for (int cc = 0; cc < 3; cc++)
                for (int nn = 0; nn < 64; nn++) qt[cc * 64 + (nn & 7) * 8 + (nn >> 3)] = qn[cc * 64 + nn];
        }
        /* Huffman LUT set: dedupe on the raw DHT content */
        const uint64_t h = jd_tables_hash(&inf);
        uint32_t li = 0;
        for (; li < lut_hash.size(); li++) if (lut_hash[li] == h && jd_tables_equal(&inf, &b->infos[lut_owner[li]])) break;
        const bool shared = ctx->has_shared && ctx->shared_hash == h && ctx->shared_hash2 == jd_tables_hash2(&inf);
        if (li == lut_hash.size()) {
            lut_hash.push_back(h); lut_owner.push_back(i);
            b->luts.resize((size_t)(li + 1) * JD_LUT_ENTRIES);
            if (shared) memcpy(&b->luts[(size_t)li * JD_LUT_ENTRIES], ctx->shared_lut, JD_LUT_ENTRIES * 2);
            else jd_build_lut(&inf, &b->luts[(size_t)li * JD_LUT_ENTRIES]);
        }
        if (shared) ctx->shared_hits++;
        const uint32_t total_mcus = (uint32_t)inf.mcus_x * inf.mcus_y;
        const uint32_t mps = inf.restart_interval ? (uint32_t)inf.restart_interval : total_mcus;
        d.scan_off = (uint32_t)(b->comp_off[i] + inf.scan_offset);
        d.scan_end = (uint32_t)(b->comp_off[i] + sizes[i]);
        d.width = (uint16_t)inf.width; d.height = (uint16_t)inf.height;
        d.mcus_x = (uint16_t)inf.mcus_x; d.mcus_y = (uint16_t)inf.mcus_y;
        d.subsample = (uint8_t)inf.subsample; d.ncomp = (uint8_t)inf.ncomp; d.bpm = (uint8_t)inf.bpm; d.tsel = (uint8_t)inf.tsel;
        d.mcus_per_seg = mps;
        d.nseg = (total_mcus + mps - 1) / mps;
        d.chunk_base = 0; d.nch = 0;
        d.prog = prog ? (1u | ((uint32_t)(inf.approx & 15) << 8)) : 0u;
        if (!prog && inf.restart_interval == 0 && d.nseg == 1 && sizes[i] - inf.scan_offset >= 4096) {
            /* no restart markers: one long dependent stream -> chunk-parallel decode */
            d.chunk_base = b->nchunks;
            d.nch = ((uint32_t)(sizes[i] - inf.scan_offset) + JD_CHUNK_BYTES - 1) / JD_CHUNK_BYTES + 1;
            b->nchunks += d.nch;
            if (d.nch > b->max_nch) b->max_nch = d.nch;
            b->cimg_list.push_back((uint32_t)i);
        }
        d.seg_base = seg;
        d.blk_base = (uint32_t)blk;
        d.lutset = li;
        /* coefficient records: image-relative indices (jd_core.h JD_REC_INDEX), one slot per restart segment and per chunk */
        d.comp_off = (uint32_t)b->comp_off[i];
        d.rec_base = rec_total;
        rec_total += (uint64_t)JD_REC_PER_BYTE * (uint64_t)(((size_t)sizes[i] + 15) & ~(size_t)15) + (uint64_t)JD_REC_SLOT_SLACK * (d.nseg + d.nch + 1u);

        const int s = b->sshift;
        d.out_w = (uint32_t)((inf.width + (1 << s) - 1) >> s);
        d.out_h = (uint32_t)((inf.height + (1 << s) - 1) >> s);
        if (b->padded) {
            d.out_w = (uint32_t)inf.mcus_x * (uint32_t)(inf.mcu_w >> s);
            d.out_h = (uint32_t)inf.mcus_y * (uint32_t)(inf.mcu_h >> s);
        }
        size_t pitch;
        if (b->dither_bits) {
            const uint32_t pw = (uint32_t)inf.mcus_x * (uint32_t)(inf.mcu_w >> s);
            pitch = ((size_t)pw * b->dither_bits + 7) / 8;
            gray_total += (((size_t)pw * (size_t)inf.mcus_y * (size_t)(inf.mcu_h >> s)) + 255) & ~(size_t)255;
        } else pitch = (size_t)d.out_w * bytes_per_pixel_class(b->ptclass);
        d.out_pitch = (uint32_t)pitch;
        b->pitches[i] = (int64_t)pitch;
        b->arena_off[i] = out_total;
        out_total += (pitch * d.out_h + 255) & ~(size_t)255;
        seg += d.nseg;
        blk += (uint64_t)total_mcus * inf.bpm;
        if (blk >= (1ull << 32)) { snprintf(g_err, sizeof(g_err), "batch too large (block count)"); delete b; return nullptr; }
    }
    b->nseg = seg; b->nblk = blk; b->nlut = (uint32_t)lut_hash.size();
    b->rec_total = rec_total;
    if ((uint64_t)b->comp_total + 32ull * seg + 4096ull >= (1ull << 32)) {
        snprintf(g_err, sizeof(g_err), "batch too large (%zu compressed bytes in %u restart segments)", b->comp_total, seg);
        delete b;
        return nullptr;
    }
    b->out_total = out_total; b->gray_total = gray_total;
    /* work list: CTAs of 128 segments sharing one LUT set */
    b->seg_img.resize(seg ? seg : 1);
    for (uint32_t li = 0; li < (b->nlut ? b->nlut : 1); li++) {
        for (int i = 0; i < n; i++) {
            const JDImageDesc &d = b->descs[i];
            if (d.nseg == 0 || b->parse_status[i] != JPEG_SUCCESS || d.lutset != li) continue;
            for (uint32_t s2 = 0; s2 < d.nseg; s2++) { b->seg_img[d.seg_base + s2] = (uint32_t)i; if (d.nch == 0) b->work.push_back(d.seg_base + s2); }
        }
        while (b->work.size() % JD_ENTROPY_THREADS) b->work.push_back(JD_NONE);
        while (b->cta_lut.size() < b->work.size() / JD_ENTROPY_THREADS) b->cta_lut.push_back(li);
    }
    return b;
}

extern "C" void JPEGB200_batchDestroy(JPEGB200_BATCH *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    b->d_comp.release(); b->d_out.release(); b->d_gray.release(); b->d_errline.release();
    b->d_gray_off.release(); b->d_err_off.release(); b->d_dprog.release(); b->d_dbands.release();
    b->d_clean.release(); b->d_seg_clen.release();
    b->d_filt.release(); b->d_cimg_list.release(); b->d_flen.release(); b->d_E0.release(); b->d_E1.release(); b->d_Ep.release(); b->d_cfirst.release();
    b->d_cn.release(); b->d_cpre.release(); b->d_cjmap.release(); b->d_cstatus.release(); b->d_cnown.release(); b->d_cdcs.release(); b->d_cpe.release();
    b->d_descs.release(); b->d_quant.release(); b->d_luts.release(); b->d_rec.release();
    b->d_work.release(); b->d_cta_lut.release(); b->d_seg_img.release(); b->d_seg_start.release();
    b->d_seg_jmap.release(); b->d_seg_status.release(); b->d_seg_nrec.release(); b->d_seg_phase.release();
    b->d_counters.release(); b->d_blk_hdr.release(); b->d_events.release();
    if (b->stream && b->have_ev) {   /* back to the context for the next job */
        JDStreamSet ss;
        ss.stream = b->stream;
        for (int i = 0; i < JPEGB200_NUM_TIMINGS + 2; i++) ss.ev[i] = b->ev[i];
        b->ctx->free_streams.push_back(ss);
    } else {
        if (b->have_ev) for (auto &e : b->ev) cudaEventDestroy(e);
        if (b->stream) cudaStreamDestroy(b->stream);
    }
    if (b->descs_dl) b->ctx->pinpool.put(b->descs_dl, b->descs_dl_bytes);
    delete b;
}

extern "C" int JPEGB200_batchCount(JPEGB200_BATCH *b) { return b ? b->n : 0; }

extern "C" int JPEGB200_batchImageInfo(JPEGB200_BATCH *b, int i, int32_t *width, int32_t *height, int32_t *subsample,
                                       int32_t *out_w, int32_t *out_h, int32_t *status)
{
    if (!b || i < 0 || i >= b->n) return 0;
    const JDInfo &inf = b->infos[i];
    if (width) *width = inf.width;
    if (height) *height = inf.height;
    if (subsample) *subsample = inf.subsample;
    if (out_w) *out_w = (int32_t)b->descs[i].out_w;
    if (out_h) *out_h = (int32_t)b->descs[i].out_h;
    if (status) *status = b->parse_status[i];
    return 1;
}

extern "C" int64_t JPEGB200_batchOutputBytes(JPEGB200_BATCH *b, int i, int64_t *pitch_bytes)
{
    if (!b || i < 0 || i >= b->n) return 0;
    if (pitch_bytes) *pitch_bytes = (int64_t)b->descs[i].out_pitch;
    return (int64_t)b->descs[i].out_pitch * b->descs[i].out_h;
}

extern "C" int JPEGB200_batchSetOutput(JPEGB200_BATCH *b, int i, void *out, int64_t pitch_bytes)
{
    if (!b || i < 0 || i >= b->n) return 0;
    b->outs[i] = out;
    if (pitch_bytes > 0) b->pitches[i] = pitch_bytes;
    return 1;
}

static int batch_stream(JPEGB200_BATCH *b)
{
    CK(cudaSetDevice(b->ctx->device));
    if (!b->stream && !b->have_ev && !b->ctx->free_streams.empty()) {
        const JDStreamSet ss = b->ctx->free_streams.back();
        b->ctx->free_streams.pop_back();
        b->stream = ss.stream;
        for (int i = 0; i < JPEGB200_NUM_TIMINGS + 2; i++) b->ev[i] = ss.ev[i];
        b->have_ev = true;
    }
    if (!b->stream) CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    if (!b->have_ev) {
        for (auto &e : b->ev) CK(cudaEventCreate(&e));
        b->have_ev = true;
    }
    return 1;
}

extern "C" void *JPEGB200_batchStream(JPEGB200_BATCH *b) { if (!b || !batch_stream(b)) return nullptr; return (void *)b->stream; }

extern "C" int JPEGB200_batchAllocDeviceOutput(JPEGB200_BATCH *b)
{
    if (!b) return 0;
    CK(cudaSetDevice(b->ctx->device));
    CK(b->d_out.alloc(&b->ctx->pool, b->out_total + 256));
    b->arena_owned = true;
    return 1;
}

extern "C" int JPEGB200_batchGetDeviceOutput(JPEGB200_BATCH *b, int i, void **devptr, int64_t *pitch_bytes)
{
    if (!b || i < 0 || i >= b->n || !b->d_out.p) return 0;
    if (devptr) *devptr = b->d_out.p + b->arena_off[i];
    if (pitch_bytes) *pitch_bytes = (int64_t)b->descs[i].out_pitch;
    return 1;
}

/* synchronous copy of image i's pixels out of the device arena (tests / spot checks) */
extern "C" int JPEGB200_batchReadOutput(JPEGB200_BATCH *b, int i, void *host_dst)
{
    if (!b || i < 0 || i >= b->n || !b->d_out.p || !host_dst) return 0;
    CK(cudaSetDevice(b->ctx->device));
    if (b->stream) CK(cudaStreamSynchronize(b->stream));
    CK(cudaMemcpy(host_dst, b->d_out.p + b->arena_off[i], (size_t)b->descs[i].out_pitch * b->descs[i].out_h, cudaMemcpyDeviceToHost));
    return 1;
}

extern "C" int JPEGB200_batchUpload(JPEGB200_BATCH *b)
{
    if (!b) return 0;
    if (!batch_stream(b)) return 0;
    const int n = b->n;
    CK(b->d_comp.alloc(&b->ctx->pool, b->comp_total + 256));
    CK(b->d_descs.alloc(&b->ctx->pool, n));
    CK(b->d_quant.alloc(&b->ctx->pool, (size_t)n * 192));
    CK(b->d_luts.alloc(&b->ctx->pool, b->luts.size() ? b->luts.size() : 1));
    CK(b->d_work.alloc(&b->ctx->pool, b->work.size() ? b->work.size() : 1));
    CK(b->d_cta_lut.alloc(&b->ctx->pool, b->cta_lut.size() ? b->cta_lut.size() : 1));
    CK(b->d_seg_img.alloc(&b->ctx->pool, b->seg_img.size()));
    const size_t ns = b->nseg ? b->nseg : 1;
    CK(b->d_seg_start.alloc(&b->ctx->pool, ns + 1)); CK(b->d_seg_jmap.alloc(&b->ctx->pool, ns)); CK(b->d_seg_status.alloc(&b->ctx->pool, ns));
    CK(b->d_seg_nrec.alloc(&b->ctx->pool, ns)); CK(b->d_seg_phase.alloc(&b->ctx->pool, ns + b->nchunks));
    if (b->nchunks) {
        const size_t nc = b->nchunks;
        CK(b->d_filt.alloc(&b->ctx->pool, b->comp_total + 512));
        CK(b->d_cimg_list.alloc(&b->ctx->pool, b->cimg_list.size())); CK(b->d_flen.alloc(&b->ctx->pool, n));
        CK(b->d_E0.alloc(&b->ctx->pool, nc + 1)); CK(b->d_E1.alloc(&b->ctx->pool, nc + 1)); CK(b->d_Ep.alloc(&b->ctx->pool, nc)); CK(b->d_cfirst.alloc(&b->ctx->pool, nc)); CK(b->d_cn.alloc(&b->ctx->pool, nc)); CK(b->d_cpre.alloc(&b->ctx->pool, nc)); CK(b->d_cjmap.alloc(&b->ctx->pool, nc));
        CK(b->d_cstatus.alloc(&b->ctx->pool, nc)); CK(b->d_cnown.alloc(&b->ctx->pool, nc)); CK(b->d_cdcs.alloc(&b->ctx->pool, 3 * nc)); CK(b->d_cpe.alloc(&b->ctx->pool, 3 * nc));
    }
    CK(b->d_counters.alloc(&b->ctx->pool, 8));
    CK(b->d_blk_hdr.alloc(&b->ctx->pool, b->nblk ? b->nblk : 1));
    CK(b->d_rec.alloc(&b->ctx->pool, b->rec_total + 1024));
    CK(b->d_events.alloc(&b->ctx->pool, JD_EVENT_CAP));
    cudaStream_t st = b->stream;
    CK(cudaEventRecord(b->ev[0], st));
    /* zero the tail padding so word loads past the last file read zeros */
    CK(cudaMemsetAsync(b->d_comp.p + b->comp_total, 0, 256, st));
    if (b->contiguous_in) {
        CK(cudaMemcpyAsync(b->d_comp.p, b->datas[0], b->comp_total, cudaMemcpyHostToDevice, st));
    } else {
        for (int i = 0; i < n; i++)
            CK(cudaMemcpyAsync(b->d_comp.p + b->comp_off[i], b->datas[i], (size_t)b->sizes[i], cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(b->d_descs.p, b->descs.data(), sizeof(JDImageDesc) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b->d_quant.p, b->quant.data(), sizeof(int32_t) * 192 * n, cudaMemcpyHostToDevice, st));
    if (b->luts.size()) CK(cudaMemcpyAsync(b->d_luts.p, b->luts.data(), b->luts.size() * 2, cudaMemcpyHostToDevice, st));
    if (b->work.size()) CK(cudaMemcpyAsync(b->d_work.p, b->work.data(), b->work.size() * 4, cudaMemcpyHostToDevice, st));
    if (b->cta_lut.size()) CK(cudaMemcpyAsync(b->d_cta_lut.p, b->cta_lut.data(), b->cta_lut.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b->d_seg_img.p, b->seg_img.data(), b->seg_img.size() * 4, cudaMemcpyHostToDevice, st));
    if (b->nchunks) {
        CK(cudaMemcpyAsync(b->d_cimg_list.p, b->cimg_list.data(), b->cimg_list.size() * 4, cudaMemcpyHostToDevice, st));
    }
    CK(cudaEventRecord(b->ev[1], st));
    b->uploaded = true;
    b->counters[JPEGB200_C_H2D_BYTES] = (int64_t)(b->comp_total + sizeof(JDImageDesc) * n + 768 * (size_t)n + b->luts.size() * 2 +
                                                  b->work.size() * 4 + b->cta_lut.size() * 4 + b->seg_img.size() * 4);
    return 1;
}

/* ---- IDCT kernel dispatch ---- */
template <int HS, int VS, int NC, int MPB, int PT>
static void launch_idct_pt(const JDIdctArgs &a, dim3 grid, int arith, bool half, cudaStream_t st)
{
    using G = JDGeo<HS, VS, NC, MPB>;
    if (arith == JPEG_ARITH_SSE2) {
        if (half) jdk_idct_color<HS, VS, NC, MPB, PT, JPEG_ARITH_SSE2, true><<<grid, G::THREADS, 0, st>>>(a);
        else jdk_idct_color<HS, VS, NC, MPB, PT, JPEG_ARITH_SSE2, false><<<grid, G::THREADS, 0, st>>>(a);
    } else {
        if (half) jdk_idct_color<HS, VS, NC, MPB, PT, JPEG_ARITH_SCALAR, true><<<grid, G::THREADS, 0, st>>>(a);
        else jdk_idct_color<HS, VS, NC, MPB, PT, JPEG_ARITH_SCALAR, false><<<grid, G::THREADS, 0, st>>>(a);
    }
}

static int g_use_tb = -1; /* thread-per-block IDCT kernel (default) unless JPEGDEC_B200_IDCT=lanes */
static int g_tb_mpb = 0;  /* development switch JPEGDEC_B200_TB_MPB=16|20: force the strip width */

template <int HS, int VS, int NC, int MPB, int PT>
static void launch_idct_tb(const JDIdctArgs &a, uint32_t mcus_x, uint32_t mcus_y, uint32_t nimg, int arith, cudaStream_t st)
{
    using G = JDGeoTB<HS, VS, NC, MPB>;
    dim3 grid((mcus_x + MPB - 1) / MPB, mcus_y, nimg);
    static bool carveout_set = false;   /* 10 CTAs of ~17-21 KB static shared memory per SM need the large carveout */
    if (!carveout_set) {
        cudaFuncSetAttribute(jdk_idct_tb<HS, VS, NC, MPB, PT, JPEG_ARITH_SSE2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(jdk_idct_tb<HS, VS, NC, MPB, PT, JPEG_ARITH_SCALAR>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carveout_set = true;
    }
    if (arith == JPEG_ARITH_SSE2) jdk_idct_tb<HS, VS, NC, MPB, PT, JPEG_ARITH_SSE2><<<grid, G::THREADS, 0, st>>>(a);
    else jdk_idct_tb<HS, VS, NC, MPB, PT, JPEG_ARITH_SCALAR><<<grid, G::THREADS, 0, st>>>(a);
}

/* SSE2-build arithmetic: the packed thread-per-block kernel for every sampling / pixel type, full and half size */
template <int HS, int VS, int NC, int MPB, int PT>
static void launch_idct_p(const JDIdctArgs &a, uint32_t mcus_x, uint32_t mcus_y, uint32_t nimg, bool half, cudaStream_t st)
{
    dim3 grid((mcus_x + MPB - 1) / MPB, mcus_y, nimg);
    static bool carveout_set = false;   /* 7-8 CTAs of ~27 KB static shared memory per SM need the large carveout */
    if (!carveout_set) {
        cudaFuncSetAttribute(jdk_idct_p<HS, VS, NC, MPB, PT, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(jdk_idct_p<HS, VS, NC, MPB, PT, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carveout_set = true;
    }
    if (half) jdk_idct_p<HS, VS, NC, MPB, PT, true><<<grid, 128, 0, st>>>(a);
    else jdk_idct_p<HS, VS, NC, MPB, PT, false><<<grid, 128, 0, st>>>(a);
}

template <int HS, int VS>
static int launch_idct_packed(const JDIdctArgs &a, uint32_t mcus_x, uint32_t mcus_y, uint32_t nimg, int ncomp, int ptclass, bool half, cudaStream_t st)
{
    constexpr int MPB1 = 128 / (HS * VS);            /* luma only: HS * VS blocks per MCU */
    constexpr int MPB3 = (HS * VS == 4) ? 20 : (HS * VS == 2) ? 30 : 40;
    if (ptclass == JD_PT_GRAY) { launch_idct_p<HS, VS, 1, MPB1, JD_PT_GRAY>(a, mcus_x, mcus_y, nimg, half, st); return 1; }
    if (ncomp == 1) {
        if (HS != 1 || VS != 1 || ptclass != JD_PT_565) return 0;
        launch_idct_p<1, 1, 1, 128, JD_PT_565>(a, mcus_x, mcus_y, nimg, half, st);
        return 1;
    }
    if (HS == 2 && VS == 2) {
        /* strips of 20 MCUs (320 px) unless strips of 16 waste fewer MCU slots (1920 and 3840 px divide evenly by 320) */
        const uint32_t pad16 = (mcus_x + 15) / 16 * 16 - mcus_x, pad20 = (mcus_x + 19) / 20 * 20 - mcus_x;
        if (pad20 <= pad16) {
            if (ptclass == JD_PT_565) launch_idct_p<2, 2, 3, 20, JD_PT_565>(a, mcus_x, mcus_y, nimg, half, st);
            else launch_idct_p<2, 2, 3, 20, JD_PT_8888>(a, mcus_x, mcus_y, nimg, half, st);
        } else {
            if (ptclass == JD_PT_565) launch_idct_p<2, 2, 3, 16, JD_PT_565>(a, mcus_x, mcus_y, nimg, half, st);
            else launch_idct_p<2, 2, 3, 16, JD_PT_8888>(a, mcus_x, mcus_y, nimg, half, st);
        }
        return 1;
    }
    if (ptclass == JD_PT_565) launch_idct_p<HS, VS, 3, MPB3, JD_PT_565>(a, mcus_x, mcus_y, nimg, half, st);
    else launch_idct_p<HS, VS, 3, MPB3, JD_PT_8888>(a, mcus_x, mcus_y, nimg, half, st);
    return 1;
}

template <int HS, int VS, int MPB3, int MPB1>
static int launch_idct_geo(const JDIdctArgs &a, uint32_t mcus_x, uint32_t mcus_y, uint32_t nimg, int ncomp, int ptclass,
                           int arith, bool half, cudaStream_t st)
{
    if (g_use_tb < 0) {
        const char *e = getenv("JPEGDEC_B200_IDCT"); g_use_tb = (e && strcmp(e, "lanes") == 0) ? 0 : 1;
        const char *m = getenv("JPEGDEC_B200_TB_MPB"); g_tb_mpb = m ? atoi(m) : 0;
    }
    if (g_use_tb && !half && HS == 2 && VS == 2 && ncomp == 3 && ptclass != JD_PT_GRAY) {
        /* 4:2:0 colour, full size: the throughput configuration */
        /* strips of 20 MCUs (320 px) unless that wastes more MCU slots than strips of 16 (1920 and 3840 px divide evenly by 320);
         * measured: HD q75 3.47 vs 3.26 ms, UHD q85 8.00 vs 7.79 ms */
        const uint32_t pad16 = (mcus_x + 15) / 16 * 16 - mcus_x, pad20 = (mcus_x + 19) / 20 * 20 - mcus_x;
        const bool wide = g_tb_mpb == 20 || (g_tb_mpb == 0 && pad20 <= pad16);
        if (wide) {
            if (ptclass == JD_PT_565) launch_idct_tb<2, 2, 3, 20, JD_PT_565>(a, mcus_x, mcus_y, nimg, arith, st);
            else launch_idct_tb<2, 2, 3, 20, JD_PT_8888>(a, mcus_x, mcus_y, nimg, arith, st);
        } else {
            if (ptclass == JD_PT_565) launch_idct_tb<2, 2, 3, 16, JD_PT_565>(a, mcus_x, mcus_y, nimg, arith, st);
            else launch_idct_tb<2, 2, 3, 16, JD_PT_8888>(a, mcus_x, mcus_y, nimg, arith, st);
        }
        return 1;
    }
    if (ptclass == JD_PT_GRAY) {
        dim3 grid((mcus_x + MPB1 - 1) / MPB1, mcus_y, nimg);
        launch_idct_pt<HS, VS, 1, MPB1, JD_PT_GRAY>(a, grid, arith, half, st);
    } else if (ncomp == 1) {
        if (HS != 1 || VS != 1 || ptclass != JD_PT_565) return 0;
        dim3 grid((mcus_x + MPB1 - 1) / MPB1, mcus_y, nimg);
        launch_idct_pt<1, 1, 1, MPB1, JD_PT_565>(a, grid, arith, half, st);
    } else {
        dim3 grid((mcus_x + MPB3 - 1) / MPB3, mcus_y, nimg);
        if (ptclass == JD_PT_565) launch_idct_pt<HS, VS, 3, MPB3, JD_PT_565>(a, grid, arith, half, st);
        else launch_idct_pt<HS, VS, 3, MPB3, JD_PT_8888>(a, grid, arith, half, st);
    }
    return 1;
}

#ifndef JD_DITHER_MINB
#define JD_DITHER_MINB 9
#endif
#ifndef JD_DITHER_SKEW
#define JD_DITHER_SKEW 2   /* pixels by which a row trails the row above: 3 = the error from above is folded into the NEXT pixel's
                             forward error (one step of slack for the shuffle); 2 = it is added to the current pixel */
#endif
template <int BITS>
__global__ void jdk_dither(const JDImageDesc *imgs, uint32_t nimg, const uint8_t *gray, const uint64_t *gray_off,
                           uint16_t *errlines, const uint32_t *err_off, uint8_t *out, uint32_t sshift,
                           const uint4 *bands, uint32_t nbands, uint32_t *progress);

extern "C" int JPEGB200_batchDecode(JPEGB200_BATCH *b, int flags)
{
    if (!b) return 0;
    if (!b->uploaded) { snprintf(g_err, sizeof(g_err), "batchDecode before batchUpload"); return 0; }
    if (!batch_stream(b)) return 0;
    cudaStream_t st = b->stream;
    const int n = b->n;
    b->out_device = (flags & JPEGB200_OUT_DEVICE) != 0;
    b->decode_flags = flags;
    int launches = 0;
    /* output placement */
    bool user_dev_out = false;
    if (b->out_device && !b->arena_owned) {
        user_dev_out = true;
        for (int i = 0; i < n; i++) if (!b->outs[i] && b->parse_status[i] == JPEG_SUCCESS) user_dev_out = false;
        bool any_ptr = false;
        for (int i = 0; i < n; i++) if (b->outs[i]) any_ptr = true;
        if (!user_dev_out && any_ptr) { snprintf(g_err, sizeof(g_err), "device output pointers given for some images only"); return 0; }
    }
    uint8_t *out_base = nullptr;
    if (user_dev_out) {
        /* user device pointers: offsets relative to the lowest pointer */
        uintptr_t lo = ~(uintptr_t)0;
        for (int i = 0; i < n; i++) if (b->outs[i] && (uintptr_t)b->outs[i] < lo) lo = (uintptr_t)b->outs[i];
        out_base = (uint8_t *)lo;
        for (int i = 0; i < n; i++) {
            b->descs[i].out_off = b->outs[i] ? (uint64_t)((uintptr_t)b->outs[i] - lo) : 0;
            b->descs[i].out_pitch = (uint32_t)b->pitches[i];
        }
    } else {
        if (!b->d_out.p) { CK(b->d_out.alloc(&b->ctx->pool, b->out_total + 256)); b->arena_owned = true; }
        out_base = b->d_out.p;
        for (int i = 0; i < n; i++) b->descs[i].out_off = b->arena_off[i];
    }
    /* dither: the IDCT stage writes an MCU-aligned 8-bit image first */
    std::vector<uint64_t> gray_off;
    std::vector<uint32_t> err_off;
    std::vector<JDImageDesc> descs_stage = b->descs;
    if (b->dither_bits) {
        CK(b->d_gray.alloc(&b->ctx->pool, b->gray_total + 256));
        size_t go = 0, eo = 0;
        gray_off.resize(2 * (size_t)n); err_off.resize(n);
        b->errinit.clear();
        for (int i = 0; i < n; i++) {
            const JDInfo &inf = b->infos[i];
            gray_off[i] = go; err_off[i] = (uint32_t)eo;
            gray_off[(size_t)n + i] = b->descs[i].out_off;
            if (b->parse_status[i] != JPEG_SUCCESS) continue;
            const uint32_t pw = (uint32_t)inf.mcus_x * (uint32_t)(inf.mcu_w >> b->sshift);
            const uint32_t ph = (uint32_t)inf.mcus_y * (uint32_t)(inf.mcu_h >> b->sshift);
            descs_stage[i].out_off = go;
            descs_stage[i].out_pitch = pw;
            go += (((size_t)pw * ph) + 255) & ~(size_t)255;
            /* initial error line = the reference's DHT scratch bytes (they share usPixels, jpeg.inl:843 / :4881) */
            const size_t el = ((size_t)pw + 16 + 15) & ~(size_t)15;
            b->errinit.resize(eo + el, (uint16_t)0xFF00u);
            /* device line S[x] = errors[x + 2] */
            const size_t cp = (el + 2 < JD_HUFFVALS_BYTES) ? el : JD_HUFFVALS_BYTES - 2;
            for (size_t q = 0; q < cp; q++) b->errinit[eo + q] = (uint16_t)(0xFF00u | inf.p.huffvals[q + 2]);
            eo += el;
        }
        CK(b->d_errline.alloc(&b->ctx->pool, eo + 16));
        CK(cudaMemcpyAsync(b->d_errline.p, b->errinit.data(), eo * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
        CK(b->d_gray_off.alloc(&b->ctx->pool, 2 * (size_t)n)); CK(b->d_err_off.alloc(&b->ctx->pool, n));
        /* pageable sources: the runtime stages them before returning, so the vectors may go out of scope */
        CK(cudaMemcpyAsync(b->d_gray_off.p, gray_off.data(), (size_t)n * 16, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b->d_err_off.p, err_off.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        /* one warp per band of 32 rows, band-major (band k of every image, then band k + 1): a band's producer is always
         * launched before it, and the warps resident at any time are bands that can actually run (a band may start ~113
         * steps after the one above it, so only ~W/113 bands of an image are ever active together) */
        b->dbands.clear();
        {
            uint32_t maxb = 0;
            std::vector<uint32_t> nb(n, 0), prevpos(n, 0);
            for (int i = 0; i < n; i++) if (b->parse_status[i] == JPEG_SUCCESS) { nb[i] = (b->descs[i].out_h + 31) / 32; if (nb[i] > maxb) maxb = nb[i]; }
            for (uint32_t k = 0; k < maxb; k++)
                for (int i = 0; i < n; i++) {
                    if (k >= nb[i]) continue;
                    const uint32_t pos = (uint32_t)b->dbands.size();
                    b->dbands.push_back(make_uint4((uint32_t)i, k, prevpos[i], 0u));
                    prevpos[i] = pos;
                }
        }
        CK(b->d_dbands.alloc(&b->ctx->pool, b->dbands.size() ? b->dbands.size() : 1)); CK(b->d_dprog.alloc(&b->ctx->pool, b->dbands.size() + 1));
        if (!b->dbands.empty()) CK(cudaMemcpyAsync(b->d_dbands.p, b->dbands.data(), b->dbands.size() * sizeof(uint4), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(b->d_dprog.p, 0, (b->dbands.size() + 1) * 4, st));
    }
    CK(cudaMemcpyAsync(b->d_descs.p, descs_stage.data(), sizeof(JDImageDesc) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(b->d_counters.p, 0, 32, st));
    if (b->nchunks) CK(cudaMemsetAsync(b->d_blk_hdr.p, 0, (size_t)b->nblk * 8, st)); /* blocks a truncated restart-free scan never reaches stay empty */

    CK(cudaEventRecord(b->ev[2], st));
    jdk_prescan<<<n, 256, 0, st>>>(b->d_comp.p, b->d_descs.p, b->d_seg_start.p);
    launches++;
    CK(cudaEventRecord(b->ev[3], st));
    if (!b->work.empty()) {
        JDEntropyArgs ea;
        ea.data = b->d_comp.p; ea.imgs = b->d_descs.p; ea.luts = b->d_luts.p; ea.work = b->d_work.p; ea.cta_lut = b->d_cta_lut.p;
        ea.seg_img = b->d_seg_img.p; ea.seg_start = b->d_seg_start.p; ea.blk_hdr = b->d_blk_hdr.p; ea.rec = b->d_rec.p;
        ea.seg_jmap = b->d_seg_jmap.p; ea.seg_status = b->d_seg_status.p; ea.seg_nrec = b->d_seg_nrec.p;
        ea.events = b->d_events.p; ea.event_count = b->d_counters.p; ea.event_cap = JD_EVENT_CAP;
        ea.nwork = (uint32_t)b->work.size(); ea.dc_output = (b->sshift == 3) ? 1u : (b->sshift == 2) ? 2u : 0u;
        /* JPEGDEC_B200_ENTROPY=raw: the entropy kernel un-stuffs in its bit reader; default ("clean"): jdk_unstuff_segs
         * first, so that the reader is a plain word stream */
        static int use_clean = -1;
        if (use_clean < 0) { const char *e = getenv("JPEGDEC_B200_ENTROPY"); use_clean = (e && strcmp(e, "raw") == 0) ? 0 : 1; }
        const unsigned egrid = (unsigned)(b->work.size() / JD_ENTROPY_THREADS);
        if (use_clean) {
            CK(b->d_clean.alloc(&b->ctx->pool, b->comp_total + 32 * (size_t)b->nseg + 4096));
            CK(b->d_seg_clen.alloc(&b->ctx->pool, b->nseg ? b->nseg : 1));
            jdk_unstuff_segs<<<(b->nseg * 32u + JD_UNSTUFF_WARPS * 32u - 1u) / (JD_UNSTUFF_WARPS * 32u), JD_UNSTUFF_WARPS * 32, 0, st>>>(
                b->d_comp.p, b->d_descs.p, b->d_seg_img.p, b->d_seg_start.p, b->nseg, b->d_clean.p, b->d_seg_clen.p);
            ea.clean = b->d_clean.p; ea.seg_clen = b->d_seg_clen.p;
            jdk_entropy<true><<<egrid, JD_ENTROPY_THREADS, 0, st>>>(ea);
            launches += 2;
        } else {
            ea.clean = nullptr; ea.seg_clen = nullptr;
            jdk_entropy<false><<<egrid, JD_ENTROPY_THREADS, 0, st>>>(ea);
            launches++;
        }
    }
    if (b->nchunks) {
        /* restart-free scans: un-stuff, iterate the chunk entry states to their fix point, then emit */
        JDChunkArgs ca;
        ca.comp = b->d_comp.p; ca.filt = b->d_filt.p; ca.imgs = b->d_descs.p; ca.luts = b->d_luts.p;
        ca.cimg_list = b->d_cimg_list.p; ca.ncimg = (uint32_t)b->cimg_list.size(); ca.flen = b->d_flen.p;
        ca.nchunks = b->nchunks;
        ca.cn = b->d_cn.p; ca.cpre = b->d_cpre.p; ca.cjmap = b->d_cjmap.p; ca.cstatus = b->d_cstatus.p; ca.cnown = b->d_cnown.p;
        ca.cdcs = b->d_cdcs.p; ca.cpe = b->d_cpe.p; ca.changed = b->d_counters.p + 2;
        ca.blk_hdr = b->d_blk_hdr.p; ca.rec = b->d_rec.p;
        ca.events = b->d_events.p; ca.event_count = b->d_counters.p; ca.event_cap = JD_EVENT_CAP;
        ca.seg_phase = b->d_seg_phase.p; ca.seg_jmap = b->d_seg_jmap.p; ca.seg_status = b->d_seg_status.p; ca.nseg_total = b->nseg;
        const unsigned gi = ((unsigned)b->cimg_list.size() * 32 + 127) / 128;
        ca.max_nch = b->max_nch; ca.Ep = b->d_Ep.p; ca.cfirst = b->d_cfirst.p;
        const dim3 gchunks((b->max_nch + 127) / 128, (unsigned)b->cimg_list.size());
        /* guess: every chunk starts a block at its first bit (exit state of every left neighbour = (0, 0, 0)); no chunk parsed yet */
        CK(cudaMemsetAsync(b->d_E0.p, 0, (size_t)(b->nchunks + 1) * 4, st));
        CK(cudaMemsetAsync(b->d_Ep.p, 0xFE, (size_t)b->nchunks * 4, st));
        {
            const dim3 gu(((b->max_nch * JD_CHUNK_BYTES + JD_UNSTUFF_PIECE - 1) / JD_UNSTUFF_PIECE + 3) / 4, (unsigned)b->cimg_list.size());
            jdk_unstuff<false><<<gu, 128, 0, st>>>(ca);
            jdk_unstuff<true><<<gu, 128, 0, st>>>(ca);
        }
        launches += 2;
        uint32_t *Xin = b->d_E0.p, *Xout = b->d_E1.p;
        int passes = 0;
        /* The entry states reach their fix point in 2-4 passes on real streams (a chunk re-synchronises well inside its 512
         * bytes), and from the third pass on only the chunks whose entry state moved are parsed again.  Normal mode:
         * JD_CHUNK_PASSES passes back to back, the last one verifying (it raises a flag if an exit state still moved) -- no
         * host round trip, so jobs of JPEGB200_decodeBatch stay in flight; batchWait re-runs the job in the iterating mode
         * below if the flag came back set. */
        static int fixed_passes = -1;   /* JPEGDEC_B200_CHUNK_PASSES=n: test hook (n = 1 forces the fallback) */
        if (fixed_passes < 0) { const char *e = getenv("JPEGDEC_B200_CHUNK_PASSES"); fixed_passes = (e && atoi(e) > 0) ? atoi(e) : JD_CHUNK_PASSES; }
        const int fixed = b->chunk_iterate ? 0 : fixed_passes;
        for (;;) {
            const int burst = fixed ? fixed : 3;
            for (int k = 0; k < burst; k++) {
                if (k == burst - 1) CK(cudaMemsetAsync(b->d_counters.p + 2, 0, 4, st));
                ca.X_in = Xin; ca.X_out = Xout;
                jdk_chunk_parse<<<gchunks, 128, 0, st>>>(ca);
                launches++; passes++;
                uint32_t *tmp = Xin; Xin = Xout; Xout = tmp;
            }
            if (fixed) break;
            CK(cudaMemcpyAsync(&b->h_changed, b->d_counters.p + 2, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (!b->h_changed || passes > (int)b->nchunks + 8) break;
        }
        ca.X_in = Xin; ca.X_out = Xout;
        jdk_chunk_prefix<<<gi, 128, 0, st>>>(ca);
        jdk_chunk_emit<<<gchunks, 128, 0, st>>>(ca);
        jdk_chunk_stitch<<<gi, 128, 0, st>>>(ca);
        launches += 3;
    }
    CK(cudaEventRecord(b->ev[4], st));
    jdk_stitch<<<(n + 127) / 128, 128, 0, st>>>(b->d_descs.p, (uint32_t)n, b->d_seg_jmap.p, b->d_seg_status.p, b->d_seg_phase.p, b->d_seg_nrec.p,
                                               reinterpret_cast<unsigned long long *>(b->d_counters.p + 4));
    jdk_patch<<<32, 256, 0, st>>>(b->d_descs.p, b->d_events.p, b->d_counters.p, JD_EVENT_CAP, b->d_seg_phase.p, b->d_blk_hdr.p, b->d_rec.p, b->d_counters.p + 1);
    launches += 2;
    CK(cudaEventRecord(b->ev[5], st));
    /* IDCT + colour: one launch per run of images with the same geometry class */
    const bool half = b->sshift == 1;
    uint8_t *stage_out = b->dither_bits ? b->d_gray.p : out_base;
    for (int i0 = 0; i0 < n;) {
        if (b->parse_status[i0] != JPEG_SUCCESS) { i0++; continue; }
        const JDInfo &f = b->infos[i0];
        int i1 = i0 + 1;
        uint32_t max_mx = f.mcus_x, max_my = f.mcus_y;
        while (i1 < n && i1 - i0 < 65535 && b->parse_status[i1] == JPEG_SUCCESS && b->infos[i1].subsample == f.subsample &&
               b->infos[i1].ncomp == f.ncomp && (b->sshift >= 2 || (b->infos[i1].width == f.width && b->infos[i1].height == f.height))) {
            if ((uint32_t)b->infos[i1].mcus_x > max_mx) max_mx = b->infos[i1].mcus_x;
            if ((uint32_t)b->infos[i1].mcus_y > max_my) max_my = b->infos[i1].mcus_y;
            i1++;
        }
        const uint32_t nimg = (uint32_t)(i1 - i0);
        if (b->sshift >= 2) {
            JDScaledArgs sa;
            sa.imgs = b->d_descs.p; sa.blk_hdr = b->d_blk_hdr.p; sa.rec = b->d_rec.p; sa.quant = b->d_quant.p;
            sa.out = stage_out; sa.img0 = (uint32_t)i0; sa.pixel_type = (uint32_t)b->pixel_type; sa.eighth = (b->sshift == 3);
            sa.padded = (b->dither_bits || b->padded) ? 1u : 0u;
            dim3 grid((max_mx * max_my + 127) / 128, nimg);
            jdk_scaled<<<grid, 128, 0, st>>>(sa);
        } else {
            JDIdctArgs ia;
            ia.imgs = b->d_descs.p; ia.blk_hdr = b->d_blk_hdr.p; ia.rec = b->d_rec.p; ia.quant = b->d_quant.p;
            ia.out = stage_out; ia.img0 = (uint32_t)i0;
            ia.big_endian = (f.ncomp == 1) ? (b->pixel_type != RGB565_LITTLE_ENDIAN) : (b->pixel_type == RGB565_BIG_ENDIAN);
            ia.padded = (b->dither_bits || b->padded) ? 1u : 0u;
            ia.mcus_x = (uint32_t)f.mcus_x; ia.mcus_y = (uint32_t)f.mcus_y; ia.width = (uint32_t)f.width; ia.height = (uint32_t)f.height;
            ia.bpm = (uint32_t)f.bpm;
            int ok = 0;
            const int ar = b->ctx->arith;
            static int use_packed = -1;   /* JPEGDEC_B200_IDCT=lanes|tb: the round-1 kernels also for the SSE2-build arithmetic (A/B) */
            if (use_packed < 0) { const char *e = getenv("JPEGDEC_B200_IDCT"); use_packed = (e && (strcmp(e, "lanes") == 0 || strcmp(e, "tb") == 0)) ? 0 : 1; }
            /* 4:2:0 colour at full size keeps jdk_idct_tb (measured faster there: 3.25 vs 3.95 ms on 1024 x HD -- packed
             * 16-bit subtraction costs three instructions on this part, which eats what the packed adds save); every other
             * sampling / pixel type / half scale takes the packed thread-per-block kernel instead of the 8-lanes-per-block one */
            const bool tb_case = f.subsample == 0x22 && f.ncomp == 3 && b->ptclass != JD_PT_GRAY && !half;
            static int force_packed = -1;
            if (force_packed < 0) { const char *e = getenv("JPEGDEC_B200_IDCT"); force_packed = (e && strcmp(e, "packed") == 0) ? 1 : 0; }
            if (ar == JPEG_ARITH_SSE2 && use_packed && (!tb_case || force_packed)) {
                switch (f.subsample) {
                    case 0x00: case 0x11: ok = launch_idct_packed<1, 1>(ia, max_mx, max_my, nimg, f.ncomp, b->ptclass, half, st); break;
                    case 0x21: ok = launch_idct_packed<2, 1>(ia, max_mx, max_my, nimg, f.ncomp, b->ptclass, half, st); break;
                    case 0x12: ok = launch_idct_packed<1, 2>(ia, max_mx, max_my, nimg, f.ncomp, b->ptclass, half, st); break;
                    case 0x22: ok = launch_idct_packed<2, 2>(ia, max_mx, max_my, nimg, f.ncomp, b->ptclass, half, st); break;
                }
            } else
            switch (f.subsample) {
                case 0x00: case 0x11: ok = launch_idct_geo<1, 1, 16, 32>(ia, max_mx, max_my, nimg, f.ncomp, b->ptclass, ar, half, st); break;
                case 0x21: ok = launch_idct_geo<2, 1, 8, 16>(ia, max_mx, max_my, nimg, f.ncomp, b->ptclass, ar, half, st); break;
                case 0x12: ok = launch_idct_geo<1, 2, 8, 16>(ia, max_mx, max_my, nimg, f.ncomp, b->ptclass, ar, half, st); break;
                case 0x22: ok = launch_idct_geo<2, 2, 8, 8>(ia, max_mx, max_my, nimg, f.ncomp, b->ptclass, ar, half, st); break;
            }
            if (!ok) { snprintf(g_err, sizeof(g_err), "no kernel for subsample 0x%02x / pixel type %d", f.subsample, b->pixel_type); return 0; }
        }
        launches++;
        i0 = i1;
    }
    CK(cudaEventRecord(b->ev[6], st));
    if (b->dither_bits) {
        if (!b->dbands.empty()) {
            const unsigned dgrid = ((unsigned)b->dbands.size() * 32 + 127) / 128;
#define JD_DITHER_ARGS b->d_descs.p, (uint32_t)n, b->d_gray.p, b->d_gray_off.p, b->d_errline.p, b->d_err_off.p, out_base, (uint32_t)b->sshift, \
                       b->d_dbands.p, (uint32_t)b->dbands.size(), b->d_dprog.p
            if (b->dither_bits == 1) jdk_dither<1><<<dgrid, 128, 0, st>>>(JD_DITHER_ARGS);
            else if (b->dither_bits == 2) jdk_dither<2><<<dgrid, 128, 0, st>>>(JD_DITHER_ARGS);
            else jdk_dither<4><<<dgrid, 128, 0, st>>>(JD_DITHER_ARGS);
#undef JD_DITHER_ARGS
            launches++;
        }
    }
    CK(cudaEventRecord(b->ev[7], st));
    CK(cudaGetLastError());
    b->counters[JPEGB200_C_LAUNCHES] = launches;
    b->counters[JPEGB200_C_SEGMENTS] = b->nseg;
    b->counters[JPEGB200_C_BLOCKS] = (int64_t)b->nblk;
    b->counters[JPEGB200_C_COMPRESSED_BYTES] = (int64_t)b->comp_total;
    int64_t ob = 0;
    for (int i = 0; i < n; i++) if (b->parse_status[i] == JPEG_SUCCESS) ob += (int64_t)b->pitches[i] * b->descs[i].out_h;
    b->counters[JPEGB200_C_OUTPUT_BYTES] = ob;
    return 1;
}

extern "C" int JPEGB200_batchDownload(JPEGB200_BATCH *b)
{
    if (!b || !b->stream) return 0;
    cudaStream_t st = b->stream;
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaEventRecord(b->ev[8], st));
    int64_t bytes = 0;
    if (!b->out_device) {
        const int n = b->n;
        /* one copy when the user's buffers mirror the arena layout, else one 2-D copy per image */
        bool mirror = true;
        for (int i = 0; i < n && mirror; i++) {
            if (b->parse_status[i] != JPEG_SUCCESS) continue;
            if (!b->outs[i] || !b->outs[0]) { mirror = false; break; }
            if ((uint8_t *)b->outs[i] - (uint8_t *)b->outs[0] != (ptrdiff_t)b->arena_off[i]) mirror = false;
            if (b->pitches[i] != (int64_t)b->descs[i].out_pitch) mirror = false;
        }
        int last_ok = -1;
        for (int i = n - 1; i >= 0 && last_ok < 0; i--) if (b->parse_status[i] == JPEG_SUCCESS) last_ok = i;
        if (mirror && last_ok >= 0 && b->parse_status[0] == JPEG_SUCCESS) {
            /* up to the end of the last image that has pixels (a rejected file owns no arena space) */
            size_t span = b->arena_off[last_ok] + (size_t)b->descs[last_ok].out_pitch * b->descs[last_ok].out_h;
            CK(cudaMemcpyAsync(b->outs[0], b->d_out.p, span, cudaMemcpyDeviceToHost, st));
            bytes = (int64_t)span;
        } else {
            for (int i = 0; i < n; i++) {
                if (b->parse_status[i] != JPEG_SUCCESS || !b->outs[i]) continue;
                const JDImageDesc &d = b->descs[i];
                CK(cudaMemcpy2DAsync(b->outs[i], (size_t)b->pitches[i], b->d_out.p + b->arena_off[i], d.out_pitch, d.out_pitch, d.out_h,
                                     cudaMemcpyDeviceToHost, st));
                bytes += (int64_t)d.out_pitch * d.out_h;
            }
        }
    }
    if (!b->descs_dl) {
        b->descs_dl = (JDImageDesc *)b->ctx->pinpool.get(sizeof(JDImageDesc) * b->n + 32, &b->descs_dl_bytes);
        if (!b->descs_dl) { snprintf(g_err, sizeof(g_err), "pinned status buffer allocation failed"); return 0; }
        b->h_counters = (uint32_t *)(b->descs_dl + b->n);
    }
    CK(cudaMemcpyAsync(b->descs_dl, b->d_descs.p, sizeof(JDImageDesc) * b->n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(b->h_counters, b->d_counters.p, 32, cudaMemcpyDeviceToHost, st));
    b->downloaded = true;
    bytes += (int64_t)sizeof(JDImageDesc) * b->n + 32;
    CK(cudaEventRecord(b->ev[9], st));
    b->counters[JPEGB200_C_D2H_BYTES] = bytes;
    return 1;
}

extern "C" int JPEGB200_batchWait(JPEGB200_BATCH *b, int32_t *status)
{
    if (!b || !b->stream) return 0;
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaGetLastError());
    if (b->nchunks && b->downloaded && !b->chunk_iterate && b->h_counters[2] != 0u) {
        /* a restart-free scan whose chunk entry states had not settled after the fixed passes: decode the job again,
         * iterating to the fix point */
        b->chunk_iterate = true;
        const int again = JPEGB200_batchDecode(b, b->decode_flags) && JPEGB200_batchDownload(b);
        b->chunk_iterate = false;
        if (!again) return 0;
        CK(cudaStreamSynchronize(b->stream));
        CK(cudaGetLastError());
    }
    int all_ok = 1;
    /* more window-truncation events than the event buffer holds: some coefficients of this job were not patched, so its
     * pixels may differ from the reference's -- report that instead of returning them as good */
    const bool ev_overflow = b->downloaded && b->h_counters[0] > JD_EVENT_CAP;
    if (ev_overflow) snprintf(g_err, sizeof(g_err), "%u window-truncation events exceed the event buffer (%u): job rejected", b->h_counters[0], JD_EVENT_CAP);
    for (int i = 0; i < b->n; i++) {
        int st = b->parse_status[i];
        if (st == JPEG_SUCCESS && b->downloaded && b->descs_dl[i].status != 0) st = JPEG_DECODE_ERROR; /* jpeg.inl:5354 */
        if (st == JPEG_SUCCESS && ev_overflow) st = JPEG_DECODE_ERROR;
        if (status) status[i] = st;
        if (st != JPEG_SUCCESS) all_ok = 0;
    }
    if (b->downloaded) {
        b->counters[JPEGB200_C_EVENTS] = b->h_counters[1];            /* truncated reads the reference would have made */
        b->counters[JPEGB200_C_EVENT_CANDIDATES] = b->h_counters[0];  /* reads that are truncated for SOME start phase */
        b->counters[JPEGB200_C_RECORD_BYTES] = 2 * (int64_t)(((uint64_t)b->h_counters[5] << 32) | b->h_counters[4]);
    }
    float t;
    auto el = [&](int a, int c) { t = 0; cudaEventElapsedTime(&t, b->ev[a], b->ev[c]); return t; };
    b->ms[JPEGB200_T_H2D] = el(0, 1);
    b->ms[JPEGB200_T_PRESCAN] = el(2, 3);
    b->ms[JPEGB200_T_ENTROPY] = el(3, 4);
    b->ms[JPEGB200_T_STITCH] = el(4, 5);
    b->ms[JPEGB200_T_IDCT] = el(5, 6);
    b->ms[JPEGB200_T_DITHER] = el(6, 7);
    b->ms[JPEGB200_T_D2H] = el(8, 9);
    b->ms[JPEGB200_T_TOTAL] = el(2, 7);
    cudaGetLastError();
    return all_ok ? 1 : 2;
}

extern "C" int JPEGB200_batchErrMcu(JPEGB200_BATCH *b, int i)
{
    if (!b || i < 0 || i >= b->n) return -1;
    if (!b->downloaded) return -1;
    return b->descs_dl[i].status ? (int)b->descs_dl[i].err_mcu : -1;
}

extern "C" int JPEGB200_batchGetTimings(JPEGB200_BATCH *b, float *ms)
{
    if (!b) return 0;
    memcpy(ms, b->ms, sizeof(b->ms));
    return 1;
}

extern "C" int JPEGB200_batchGetCounters(JPEGB200_BATCH *b, int64_t *counters)
{
    if (!b) return 0;
    memcpy(counters, b->counters, sizeof(b->counters));
    return 1;
}

/* One call for a whole batch of any size.  The batch is cut into jobs, each on its own stream, all enqueued before the
 * first wait, so that job k's pixels cross PCIe (host outputs) or its IDCT runs (device outputs) while job k+1's entropy
 * kernel runs and job k+2's compressed bytes go up.  Host outputs: jobs of JD_PIPE_IMAGES images (more when the images are
 * small), so the call costs about one D2H of the pixels instead of H2D + kernels + D2H.  Device outputs: jobs of up to
 * JD_JOB_COMP_BYTES compressed bytes, which bounds the transient coefficient records (12 B per compressed byte) however
 * large the batch is; the pixels go straight to the caller's device pointers. */
#define JD_PIPE_IMAGES 64
#define JD_PIPE_MIN_BYTES ((int64_t)64 << 20)
#define JD_PIPE_INFLIGHT 6
#define JD_JOB_COMP_BYTES ((int64_t)192 << 20)
#define JD_JOB_MAX_IMAGES 4096
#define JD_PIPE_INFLIGHT_DEVICE 3
extern "C" int JPEGB200_decodeBatch(JPEGB200_CTX *ctx, const uint8_t *const *datas, const int32_t *sizes, int n,
                                    int pixel_type, int options, void *const *outs, const int64_t *pitches,
                                    int flags, int32_t *status)
{
    if (!ctx || n <= 0) return 0;
    const bool dev_out = (flags & JPEGB200_OUT_DEVICE) != 0;
    if (dev_out && !outs) { snprintf(g_err, sizeof(g_err), "JPEGB200_decodeBatch with JPEGB200_OUT_DEVICE needs the caller's device pointers"); return 0; }
    memset(ctx->last_counters, 0, sizeof(ctx->last_counters));
    memset(ctx->last_ms, 0, sizeof(ctx->last_ms));
    ctx->last_jobs = 0;
    std::vector<JPEGB200_BATCH *> jobs;
    std::vector<int> first;
    int rc = 1, all = 1;
    size_t retired = 0;
    const size_t depth = ctx->pipe_depth ? (size_t)ctx->pipe_depth : (size_t)(dev_out ? JD_PIPE_INFLIGHT_DEVICE : JD_PIPE_INFLIGHT);
    static int64_t job_bytes = 0;   /* JPEGDEC_B200_JOB_MB=n: tuning hook for the compressed bytes per job */
    if (job_bytes == 0) { const char *e = getenv("JPEGDEC_B200_JOB_MB"); job_bytes = (e && atoi(e) > 0) ? ((int64_t)atoi(e) << 20) : JD_JOB_COMP_BYTES; }
    /* jobs complete in order; retiring one = wait + per-image status + counters + buffers back to the context's pools */
    auto retire = [&](size_t k) {
        if (rc) {
            const int r = JPEGB200_batchWait(jobs[k], status ? status + first[k] : nullptr);
            if (r == 0) all = 0; else if (r == 2 && all == 1) all = 2;
            for (int c = 0; c < JPEGB200_NUM_COUNTERS; c++) ctx->last_counters[c] += jobs[k]->counters[c];
            for (int c = 0; c < JPEGB200_NUM_TIMINGS; c++) ctx->last_ms[c] += jobs[k]->ms[c];
            ctx->last_jobs++;
        }
        JPEGB200_batchDestroy(jobs[k]);
        jobs[k] = nullptr;
    };
    static int trace = -1;          /* JPEGDEC_B200_TRACE=1: host wall clock per job on stderr (development aid) */
    if (trace < 0) { const char *e = getenv("JPEGDEC_B200_TRACE"); trace = (e && atoi(e) > 0) ? 1 : 0; }
    auto now_ms = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_call = trace ? now_ms() : 0.0;
    for (int i0 = 0; i0 < n && rc;) {
        const double t0 = trace ? now_ms() : 0.0;
        /* how many images the next job takes */
        int cnt = 0;
        int64_t cb = 0;
        const int maxcnt = dev_out ? JD_JOB_MAX_IMAGES : JD_PIPE_IMAGES;
        /* (measured and dropped: ramping the job size up from a small first job and down towards the end of a device-output
         * call -- 625 UHD files: 35.6 ms against 32.0 ms with equal jobs; every job pays the full latency of an entropy walk,
         * so fewer, larger jobs win.) */
        const int64_t limit = job_bytes;
        while (i0 + cnt < n && cnt < maxcnt) {
            const int64_t sz = sizes[i0 + cnt] > 0 ? sizes[i0 + cnt] : 0;
            if (cnt > 0 && cb + sz > limit) break;
            cb += sz; cnt++;
        }
        JPEGB200_BATCH *b = JPEGB200_batchCreate(ctx, datas + i0, sizes + i0, cnt, pixel_type, options);
        if (!b) { rc = 0; break; }
        if (!dev_out && i0 + cnt < n && cnt == JD_PIPE_IMAGES) {
            int64_t ob = 0;
            for (int i = 0; i < cnt; i++) { int64_t pb = 0; ob += JPEGB200_batchOutputBytes(b, i, &pb); }
            if (ob < JD_PIPE_MIN_BYTES) { /* small images: redo with a job big enough to keep the kernels efficient */
                int64_t per = ob > 0 ? (ob + cnt - 1) / cnt : 1;
                int64_t want = (JD_PIPE_MIN_BYTES + per - 1) / per;
                int cnt2 = (int)(want < (int64_t)(n - i0) ? want : (int64_t)(n - i0));
                if (cnt2 > JD_JOB_MAX_IMAGES) cnt2 = JD_JOB_MAX_IMAGES;
                int64_t cb2 = 0; int c3 = 0;
                while (c3 < cnt2 && (c3 == 0 || cb2 + sizes[i0 + c3] <= job_bytes)) { cb2 += sizes[i0 + c3] > 0 ? sizes[i0 + c3] : 0; c3++; }
                cnt2 = c3;
                if (cnt2 > cnt) {
                    JPEGB200_batchDestroy(b);
                    cnt = cnt2;
                    b = JPEGB200_batchCreate(ctx, datas + i0, sizes + i0, cnt, pixel_type, options);
                    if (!b) { rc = 0; break; }
                }
            }
        }
        jobs.push_back(b); first.push_back(i0);
        const double t1 = trace ? now_ms() : 0.0;
        for (int i = 0; i < cnt; i++) JPEGB200_batchSetOutput(b, i, outs ? outs[i0 + i] : nullptr, pitches ? pitches[i0 + i] : 0);
        rc = JPEGB200_batchUpload(b);
        const double t2 = trace ? now_ms() : 0.0;
        rc = rc && JPEGB200_batchDecode(b, flags);
        const double t3 = trace ? now_ms() : 0.0;
        rc = rc && JPEGB200_batchDownload(b);
        const double t4 = trace ? now_ms() : 0.0;
        i0 += cnt;
        /* bound the device memory of a very large batch: at most `depth` jobs hold buffers at a time */
        while (rc && jobs.size() - retired > depth) retire(retired++);
        if (trace) fprintf(stderr, "[jpegdec_b200] job %zu (%d images) at %.2f ms: create %.2f upload %.2f decode %.2f download %.2f retire %.2f\n",
                           jobs.size() - 1, cnt, t0 - t_call, t1 - t0, t2 - t1, t3 - t2, t4 - t3, now_ms() - t4);
    }
    { const double t5 = trace ? now_ms() : 0.0;
      while (retired < jobs.size()) retire(retired++);
      if (trace) fprintf(stderr, "[jpegdec_b200] drain %.2f ms, call %.2f ms\n", now_ms() - t5, now_ms() - t_call); }
    return rc ? all : 0;
}

extern "C" int JPEGB200_lastCallTimings(JPEGB200_CTX *ctx, float *ms, int *jobs)
{
    if (!ctx || !ms) return 0;
    memcpy(ms, ctx->last_ms, sizeof(ctx->last_ms));
    if (jobs) *jobs = ctx->last_jobs;
    return 1;
}

extern "C" int JPEGB200_lastCallCounters(JPEGB200_CTX *ctx, int64_t *counters)
{
    if (!ctx || !counters) return 0;
    memcpy(counters, ctx->last_counters, sizeof(ctx->last_counters));
    return 1;
}

/* ------------------------------------------------------------------------------------ */
/* Floyd-Steinberg dither (reference JPEGDither src/jpeg.inl:4871-4940).                    */
/*                                                                                          */
/* One warp per band of 32 rows, a wavefront inside the warp and a second one across the warps  */
/* of an image.  Inside: lane l works on row (band*32 + l) and trails lane l-1 by two pixels:    */
/* the error row l-1 sends down to a pixel (e2 of its left neighbour + e3 + e4 of its right     */
/* neighbour, summed in uint8 like the reference's error line) is complete one step before the  */
/* pixel is due and travels to the next lane with one shuffle per step.  Across: the last       */
/* lane's outgoing errors go through an error line in global memory to the next band -- the     */
/* same line the reference keeps in usPixels: it persists across MCU rows, only entries 0..2    */
/* are cleared per MCU row (:4881), and before the first row it holds the DHT scratch bytes     */
/* (the host uploads them, see batchDecode).  Band b+1 runs concurrently about 80 steps behind  */
/* band b.  Every entry of the line is 16 bits: the error and the number (mod 256) of the band  */
/* that wrote it; band b+1 reads 16 entries at a time and asks again until all of them carry    */
/* band b's number, so value and "ready" arrive in one store and the bands need no counters or  */
/* fences between them.  Each entry is read by band b+1 before band b+1 overwrites it (64 steps */
/* later), so one line per image serves all bands, as in the reference.  An image is a chain of  */
/* ~(bands x 80 + width) dependent steps whatever the batch size; with fewer than ~500 images    */
/* that chain, not throughput, sets the kernel's time.                     */
/* ------------------------------------------------------------------------------------ */
/* 16 bytes starting at byte offset `mo` (0..15) of the 32-byte pair (a, b) */
__device__ __forceinline__ uint4 jd_window16(const uint4 a, const uint4 b, uint32_t mo)
{
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const uint32_t ws = mo >> 2, bs = (mo & 3u) * 8u;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        /* v[i] = w[ws + i] without dynamic register indexing */
        uint32_t x = w[i];
        if (ws == 1) x = w[i + 1]; else if (ws == 2) x = w[i + 2]; else if (ws == 3) x = (i + 3 < 8) ? w[i + 3] : 0u;
        v[i] = x;
    }
    return make_uint4(__funnelshift_r(v[0], v[1], bs), __funnelshift_r(v[1], v[2], bs), __funnelshift_r(v[2], v[3], bs),
                      __funnelshift_r(v[3], v[4], bs));
}

template <int BITS /* output bits per pixel: 1, 2, 4 */>
__global__ void __launch_bounds__(128, JD_DITHER_MINB)
jdk_dither(const JDImageDesc *imgs, uint32_t nimg, const uint8_t *gray, const uint64_t *gray_off,
           uint16_t *errlines, const uint32_t *err_off, uint8_t *out, uint32_t sshift,
           const uint4 *bands, uint32_t nbands, uint32_t *progress)
{
    constexpr uint32_t bits = BITS;
    /* Bands are handed out through a ticket counter (progress[nbands]; the words before it are unused since the bands signal each
     * other through the error line -- but shrinking this buffer to the one counter measured 10 % slower, 2.74 vs 2.49 ms per 256
     * images on the same GPU, with identical SASS: kept as it was) in the order in which warps START, not by warp index:
     * the list is band-major (band k of every image before band k + 1), so the band a warp waits on was always claimed by a
     * warp that is already running -- forward progress does not depend on the order in which the hardware schedules CTAs. */
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t wg = 0;
    if (lane == 0) wg = atomicAdd(progress + nbands, 1u);
    wg = __shfl_sync(0xffffffffu, wg, 0);
    if (wg >= nbands) return;
    const uint4 bd = bands[wg];
    const uint32_t i = bd.x, bi = bd.y;
    const JDImageDesc &im = imgs[i];
    const uint32_t hs = (im.subsample >> 4) ? (im.subsample >> 4) : 1, vs = (im.subsample & 15) ? (im.subsample & 15) : 1;
    const uint32_t mcu_h = (vs * 8) >> sshift;
    const int W = (int)((uint32_t)im.mcus_x * ((hs * 8) >> sshift)); /* padded width = pitch of the gray stage (multiple of 8) */
    const uint32_t rows = im.out_h;
    const uint8_t *src = gray + gray_off[i];
    /* S[x] = error flowing from the row above into pixel x+1 (the reference's errors[x + 2]) in the low byte, and in the high
     * byte the number (mod 256) of the band that wrote it: the band below polls the entries themselves until they carry the
     * tag of the band above it.  Value and tag travel in one 16-bit store, so no fence and no progress counter is needed (a
     * release store per 16-32 steps cost 2.2 us each on the critical path of an image). */
    uint16_t *S = errlines + err_off[i];
    const uint32_t tag_mine = (bi & 0xFFu) << 8, tag_above = ((bi - 1u) & 0xFFu) * 0x01000100u;
    uint8_t *o = out + gray_off[nimg + i];
    const uint32_t dpitch = ((uint32_t)W * bits + 7) / 8;
    const int mask = (bits == 4) ? 0xF0 : (bits == 2 ? 0xC0 : 0x80);
    const uint32_t xmask = (bits == 4) ? 1u : (bits == 2 ? 3u : 7u);
    const bool vec = ((W & 15) == 0);
    /* Lane l works on pixel x = t - JD_DITHER_SKEW * l at step t.  To keep every global load at a warp-uniform step (a load into a
     * register that other lanes are still consuming would serialise the whole warp on the scoreboard), each lane reads
     * its row through a pointer skewed by that many bytes: at step t every lane needs byte t of its skewed row, so all lanes
     * cross 16-byte boundaries together.  The skewed 16 bytes are cut out of two aligned chunks (jd_window16). */
    const int skew = JD_DITHER_SKEW * (int)lane;
    const uint32_t mo = (uint32_t)((16 - (skew & 15)) & 15);   /* byte offset of the window inside the aligned pair */
    const int jsh = (skew + 15) >> 4;                           /* aligned chunk index of window m = m - jsh */
    const int nchunks = W >> 4;
    /* 16 line entries starting at entry 16 * m (two 16-byte loads that bypass L1 and are never hoisted) */
    auto line_load = [&](int m, uint4 &lo, uint4 &hi) {
        const uint16_t *q = S + 16 * m;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "l"(q) : "memory");
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(q + 8) : "memory");
    };
    auto line_ok = [&](const uint4 &lo, const uint4 &hi) {
        const uint32_t bad = ((lo.x ^ tag_above) | (lo.y ^ tag_above) | (lo.z ^ tag_above) | (lo.w ^ tag_above) |
                              (hi.x ^ tag_above) | (hi.y ^ tag_above) | (hi.z ^ tag_above) | (hi.w ^ tag_above)) & 0xFF00FF00u;
        return bad == 0u;
    };
    /* lane 0: make (lo, hi) the entries of window m as the band above left them */
    auto line_settle = [&](int m, uint4 &lo, uint4 &hi) {
        uint32_t ns = 128;
        while (!line_ok(lo, hi)) { __nanosleep(ns); if (ns < 1024u) ns *= 2u; line_load(m, lo, hi); }
    };
    {
        const uint32_t band = bi * 32u;
        const uint32_t y = band + lane;
        const bool live = y < rows;
        const bool mcu_first = (y % mcu_h) == 0;        /* errors[0..2] are cleared at each JPEGDither call */
        const uint8_t *p = src + (size_t)(live ? y : 0) * W;
        uint8_t *d = o + (size_t)(live ? y : 0) * dpitch;
        int fwd = 0;                 /* lFErr: e1 of the previous pixel + error arriving from above */
        int e2_prev = 0;             /* e2(x-1) */
        int down_m1 = 0;             /* partial outgoing error for pixel x-1: e2(x-2) + e3(x-1) */
        uint32_t acc = 0;
        uint32_t from_above = 0;     /* D[x+1] of the row above, delivered by the previous step's shuffle */
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        /* aligned chunks A0 = chunk(m - jsh), A1 = chunk(m - jsh + 1), A2 = prefetch of chunk(m - jsh + 2) */
        uint4 A0 = zero4, A1 = zero4, win = zero4;
        uint4 ewin = zero4;                    /* lane 0: the 16 error values of this window (the entries' low bytes) */
        auto line_values = [](const uint4 &lo, const uint4 &hi) {
            return make_uint4(__byte_perm(lo.x, lo.y, 0x6420), __byte_perm(lo.z, lo.w, 0x6420), __byte_perm(hi.x, hi.y, 0x6420), __byte_perm(hi.z, hi.w, 0x6420));
        };
        auto chunk = [&](int j) -> uint4 {
            return (live && j >= 0 && j < nchunks) ? *reinterpret_cast<const uint4 *>(p + 16 * j) : zero4;
        };
        /* what the next window switch will load is requested one window ahead with prefetches (no registers held across the
         * 16 unrolled steps): the pixel chunk into L1, the line entries -- written by another SM -- into L2 */
        auto prefetch_next = [&](int m) {
            const int j = m - jsh + 1;
            if (live && j >= 0 && j < nchunks) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 16 * j));
            if (lane == 0 && m < nchunks) asm volatile("prefetch.global.L2 [%0];" ::"l"(S + 16 * m));
        };
        if (vec) {
            A0 = chunk(-jsh); A1 = chunk(1 - jsh);
            if (lane == 0) {
                uint4 lo, hi;
                line_load(0, lo, hi);
                line_settle(0, lo, hi);
                ewin = line_values(lo, hi);
            }
            win = jd_window16(A0, A1, mo);
            prefetch_next(1);
        }
        const bool parks = live && (lane == 31 || y + 1 == rows);
        const int nsteps = W + JD_DITHER_SKEW * 31 + 2;
        uint32_t line_prev = 0;     /* lane 0, skew 2: the line entry of the previous step (= error into the current pixel) */
        for (int tb = 0; tb < nsteps; tb += 16) {
#pragma unroll
        for (int k = 0; k < 16; k++) {                    /* unrolled: byte k of the 16-byte windows is a constant extract */
            const int t = tb + k;
            const int x = t - skew;
            const bool inrow = live && x >= 0 && x < W;
            uint32_t pix, inc = from_above;
            if (vec) {
                /* warp-uniform: byte t of every lane's skewed row, and (lane 0) entry t of the error line */
                const uint32_t ww = (k < 4) ? win.x : (k < 8) ? win.y : (k < 12) ? win.z : win.w;
                const uint32_t ee = (k < 4) ? ewin.x : (k < 8) ? ewin.y : (k < 12) ? ewin.z : ewin.w;
                pix = (ww >> (8 * (k & 3))) & 0xFFu;
                if (lane == 0) {
                    const uint32_t line_now = (ee >> (8 * (k & 3))) & 0xFFu;   /* S[t] = error into pixel t + 1 */
                    inc = (JD_DITHER_SKEW == 3) ? line_now : line_prev;
                    line_prev = line_now;
                }
            } else {
                pix = inrow ? p[x] : 0u;
                if (lane == 0 && inrow && (JD_DITHER_SKEW == 3 || x >= 1)) {
                    /* unusual widths: entry by entry */
                    uint32_t v, ns = 128;
                    for (;;) {
                        asm volatile("ld.volatile.global.u16 %0, [%1];" : "=r"(v) : "l"(S + (JD_DITHER_SKEW == 3 ? x : x - 1)) : "memory");
                        if (((v ^ tag_above) & 0xFF00u) == 0u) break;
                        __nanosleep(ns); if (ns < 1024u) ns *= 2u;
                    }
                    inc = v & 0xFFu;
                }
            }
            uint32_t dcomplete = 0;   /* outgoing error for pixel x-1, complete after this step */
            if (inrow) {
                int c = (int)pix + fwd;
                if (JD_DITHER_SKEW == 2) {
                    /* error arriving at THIS pixel from the row above: none at pixel 0, and pixel 1's slot (errors[2]) is cleared at
                     * the first row of every MCU row */
                    uint32_t upc = inc & 0xFFu;
                    if (x == 0 || (mcu_first && x == 1)) upc = 0;
                    c += (int)upc;
                }
                if (c > 255) c = 255;
                acc = ((acc << bits) | ((uint32_t)c >> (8 - bits))) & 0xFFu;
                if (((uint32_t)x & xmask) == xmask) { *d++ = (uint8_t)acc; acc = 0; }
                const int v = c - (c & mask);
                const int h = v >> 1;
                const int e1 = (7 * h) >> 3, e2 = h - e1, e3 = (5 * h) >> 3, e4 = h - e3;
                /* error arriving at pixel x+1 from the row above; pixel 1's slot (errors[2]) is cleared at the first row of
                 * every MCU row, and nothing ever reaches pixel 0 from above (lFErr starts at 0) */
                uint32_t up = inc & 0xFFu;
                if (mcu_first && x == 0) up = 0;
                fwd = (JD_DITHER_SKEW == 3) ? e1 + (int)up : e1;
                dcomplete = (uint32_t)(down_m1 + e4) & 0xFFu;   /* D[x-1] = e2(x-2) + e3(x-1) + e4(x) */
                down_m1 = e2_prev + e3;                          /* becomes D[x] once e4(x+1) arrives */
                e2_prev = e2;
            } else if (live && x == W) {
                dcomplete = (uint32_t)down_m1 & 0xFFu;           /* D[W-1] = e2(W-2) + e3(W-1) (no right neighbour) */
            }
            /* the last row of the band parks what it sends down: D[x-1] feeds pixel x-1 of the next band's first row = S[x-2].
             * One entry more than anybody consumes (x = W + 1 -> S[W-1], value 0): the band below waits for the tags of whole
             * windows. */
            if (parks && x >= 2 && x <= W + 1) {
                const uint16_t ev = (uint16_t)(dcomplete | tag_mine);
                asm volatile("st.relaxed.gpu.global.u16 [%0], %1;" ::"l"(S + (x - 2)), "h"(ev) : "memory");
            }
            /* next step lane l+1 handles pixel x-2 and needs D[x-1] of this row */
            from_above = __shfl_up_sync(0xffffffffu, dcomplete, 1);
        }
            /* ---- every 16 steps: next windows ---- */
            if (vec) {
                const int m = (tb >> 4) + 1;               /* next window index */
                A0 = A1; A1 = chunk(m - jsh + 1);
                win = jd_window16(A0, A1, mo);
                if (lane == 0 && m < nchunks) {
                    /* usually the band above wrote these entries long before (it runs >= 95 + 16 steps ahead), else ask again until
                     * they carry its tag */
                    /* read now, not a window ahead: a band that follows the one above in lock step would mostly have read
                     * entries that were not written yet (measured: 2.92 vs 2.49 ms per 256 images) */
                    uint4 lo, hi;
                    line_load(m, lo, hi);
                    line_settle(m, lo, hi);
                    ewin = line_values(lo, hi);
                }
                prefetch_next(m + 1);
            }
        }
    }
}
