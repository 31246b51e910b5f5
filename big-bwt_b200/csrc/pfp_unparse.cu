// pfp_unparse.cu -- the inverse of the parse: the text back from the dictionary and the parse.
// SURVEY.md section 8(f) row 4.
//
// Replaces unparse.c of the reference (main(), unparse.c:76-137): it maps the .dicz file (the
// dictionary of `newscan -c`: every word without its last w bytes, the first one without its
// leading 0x02, newscan.cpp:410-413), notes where the words start (:93-107) and writes, for every
// symbol of .parse, the word it names (:115-124).  Here, for both streams resident in HBM:
//   1. terminator flags of the dictionary bytes, a scan, the end position of every word;
//   2. the payload length of the word of every phrase, a scan over the phrases -> where each
//      phrase's bytes go, and the length of the text;
//   3. one warp per phrase copies the bytes (coalesced 32-byte steps).
// With strip_w > 0 the input is a plain .dict (words with their w-byte overlaps, first word with
// its 0x02, last word with its w 0x02): the same bytes are skipped on the fly, so a parse can be
// inverted without a second run in -c mode.  This is the round trip parse -> unparse == text that
// tests/test_unparse_gpu.py and tools/fullsize_check.py run at full size.
#include "pfp_common.cuh"
#include "pfp_stages.cuh"
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

constexpr int UP_T = 256;

__global__ void __launch_bounds__(UP_T) up_flags_k(const u8 *__restrict__ dict, u64 n, u8 *__restrict__ flag) {
    const u64 i = (u64)blockIdx.x * UP_T + threadIdx.x;
    if (i < n) flag[i] = dict[i] == PFP_END_OF_WORD ? 1 : 0;
}

// wend[k] = position of the k-th terminator
__global__ void __launch_bounds__(UP_T) up_ends_k(const u8 *__restrict__ flag, const u32 *__restrict__ escan, u64 n,
                                                  u64 *__restrict__ wend) {
    const u64 i = (u64)blockIdx.x * UP_T + threadIdx.x;
    if (i < n && flag[i]) wend[escan[i]] = i;
}

// payload [b, e) of word k inside the dictionary bytes
__device__ __forceinline__ void up_payload(const u8 *__restrict__ dict, const u64 *__restrict__ wend, u32 k, u32 strip_w,
                                           u64 &b, u64 &e) {
    b = k ? wend[k - 1] + 1 : 0;
    e = wend[k];
    if (strip_w) {
        if (e - b >= strip_w) e -= strip_w; else e = b;
        if (b < e && dict[b] == PFP_DOLLAR) b++;             // only the first word of the text starts with 0x02
    }
}

__global__ void __launch_bounds__(UP_T) up_lens_k(const u8 *__restrict__ dict, const u64 *__restrict__ wend, u32 nwords,
                                                  u32 strip_w, const u32 *__restrict__ parse, u64 P,
                                                  u32 *__restrict__ plen, unsigned long long *__restrict__ flags) {
    const u64 j = (u64)blockIdx.x * UP_T + threadIdx.x;
    if (j >= P) return;
    const u32 r = parse[j];
    if (r == 0 || r - 1 >= nwords) {                          // "Invalid word ID in the parse file", unparse.c:120
        atomicOr(flags, PFP_ERRBIT_INTERNAL);
        plen[j] = 0;
        return;
    }
    u64 b, e;
    up_payload(dict, wend, r - 1, strip_w, b, e);
    plen[j] = (u32)(e - b);
}

__global__ void __launch_bounds__(UP_T) up_copy_k(const u8 *__restrict__ dict, const u64 *__restrict__ wend, u32 nwords,
                                                  u32 strip_w, const u32 *__restrict__ parse, u64 P,
                                                  const u64 *__restrict__ poff, u8 *__restrict__ text) {
    const u64 j = ((u64)blockIdx.x * UP_T + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    if (j >= P) return;
    const u32 r = parse[j];
    if (r == 0 || r - 1 >= nwords) return;
    u64 b, e;
    up_payload(dict, wend, r - 1, strip_w, b, e);
    u8 *dst = text + poff[j];
    for (u64 i = b + lane; i < e; i += 32) dst[i - b] = dict[i];
}

extern "C" int pfpb200_unparse_device(pfpb200_ctx *ctx, const uint8_t *d_dict, uint64_t dict_bytes, uint32_t strip_w,
                                      const uint32_t *d_parse, uint64_t n_phrases, const uint8_t **d_text,
                                      uint64_t *n_text, float *ms) {
    if (!ctx || !d_dict || !dict_bytes || !d_text || !n_text || (n_phrases && !d_parse)) return PFPB200_E_ARG;
    *d_text = nullptr;
    *n_text = 0;
    PFP_CUDA(ctx, cudaSetDevice(ctx->device));
    pfp_release_scratch(ctx);
    if (ctx->up_out) {                                       // only our own previous output is given back
        for (size_t i = 0; i < ctx->held.size(); i++)
            if (ctx->held[i] == ctx->up_out) {
                ctx->held[i] = ctx->held.back();
                ctx->held.pop_back();
                ctx->scratch.push_back(ctx->up_out);
                break;
            }
        ctx->up_out = nullptr;
        pfp_release_scratch(ctx);
    }
    ctx->err[0] = 0;
    PfpEvents evs(2);
    if (!evs.ok) return pfp_fail(ctx, PFPB200_E_CUDA, "cudaEventCreate failed");
    PFP_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, 3 * sizeof(u64), ctx->stream));
    PFP_CUDA(ctx, cudaEventRecord(evs[0], ctx->stream));
    u8 *flag = nullptr;
    u32 *escan = nullptr, *plen = nullptr;
    u64 *wend = nullptr, *poff = nullptr;
    PFP_TRY(pfp_alloc_t(ctx, &flag, dict_bytes));
    PFP_TRY(pfp_alloc_t(ctx, &escan, dict_bytes));
    up_flags_k<<<pfp_blocks(dict_bytes, UP_T), UP_T, 0, ctx->stream>>>(d_dict, dict_bytes, flag);
    PFP_LAUNCHED(ctx);
    PFP_TRY(pfp_exclusive_scan_u8_u32(ctx, flag, escan, dict_bytes, reinterpret_cast<u32 *>(&ctx->d_flags[1])));
    PFP_CUDA(ctx, cudaMemcpyAsync(&ctx->h_flags[1], &ctx->d_flags[1], sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    PFP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const u32 nwords = (u32)ctx->h_flags[1];
    PFP_TRY(pfp_alloc_t(ctx, &wend, (size_t)nwords + 1));
    up_ends_k<<<pfp_blocks(dict_bytes, UP_T), UP_T, 0, ctx->stream>>>(flag, escan, dict_bytes, wend);
    PFP_LAUNCHED(ctx);
    PFP_TRY(pfp_free_now(ctx, flag));
    PFP_TRY(pfp_free_now(ctx, escan));
    PFP_TRY(pfp_alloc_t(ctx, &plen, n_phrases));
    PFP_TRY(pfp_alloc_t(ctx, &poff, n_phrases));
    u64 total = 0;
    if (n_phrases) {
        up_lens_k<<<pfp_blocks(n_phrases, UP_T), UP_T, 0, ctx->stream>>>(
            d_dict, wend, nwords, strip_w, d_parse, n_phrases, plen, reinterpret_cast<unsigned long long *>(&ctx->d_flags[0]));
        PFP_LAUNCHED(ctx);
        PFP_TRY(pfp_exclusive_scan_u32_u64(ctx, plen, poff, n_phrases, &ctx->d_flags[2]));
        PFP_CUDA(ctx, cudaMemcpyAsync(&ctx->h_flags[0], &ctx->d_flags[0], 3 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        PFP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->h_flags[0] & PFP_ERRBIT_INTERNAL) {
            pfp_release_scratch(ctx);
            return pfp_fail(ctx, PFPB200_E_ARG, "Invalid word ID in the parse file");
        }
        total = ctx->h_flags[2];
    }
    u8 *text = nullptr;
    PFP_TRY(pfp_alloc_t(ctx, &text, total, true));
    ctx->up_out = text;
    if (n_phrases) {
        up_copy_k<<<pfp_blocks(n_phrases * 32, UP_T), UP_T, 0, ctx->stream>>>(d_dict, wend, nwords, strip_w, d_parse,
                                                                             n_phrases, poff, text);
        PFP_LAUNCHED(ctx);
    }
    PFP_CUDA(ctx, cudaEventRecord(evs[1], ctx->stream));
    PFP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ms) cudaEventElapsedTime(ms, evs[0], evs[1]);
    pfp_release_scratch(ctx);
    *d_text = text;
    *n_text = total;
    return PFPB200_OK;
}

static int up_open_size(const char *name, u64 *size) {
    const int fd = open(name, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return -1; }
    *size = (u64)st.st_size;
    return fd;
}

// `unparse <basename> [-o outfile]`: <basename>.dicz + <basename>.parse -> outfile (def. <basename>.out)
extern "C" int pfpb200_unparse_file(pfpb200_ctx *ctx, const char *basename, const char *outname, uint64_t *n_words,
                                    uint64_t *n_text, float *ms) {
    if (!ctx || !basename) return PFPB200_E_ARG;
    PFP_CUDA(ctx, cudaSetDevice(ctx->device));
    pfp_release_scratch(ctx);
    pfp_release_held(ctx);
    char name[4096], out[4096];
    u64 db = 0, pb = 0;
    snprintf(name, sizeof(name), "%s.dicz", basename);
    int fd = up_open_size(name, &db);
    if (fd < 0) return pfp_fail(ctx, PFPB200_E_IO, "cannot open %s: %s", name, strerror(errno));
    u8 *d_dict = nullptr;
    u32 *d_parse = nullptr;
    int rc = pfp_alloc_t(ctx, &d_dict, db, true);             // held: the call below releases the scratch
    if (rc == PFPB200_OK && db) rc = pfp_file_to_device(ctx, fd, 0, db, d_dict);
    close(fd);
    if (rc != PFPB200_OK) return rc;
    snprintf(name, sizeof(name), "%s.parse", basename);
    fd = up_open_size(name, &pb);
    if (fd < 0) return pfp_fail(ctx, PFPB200_E_IO, "cannot open %s: %s", name, strerror(errno));
    if (pb % 4 != 0) { close(fd); return pfp_fail(ctx, PFPB200_E_ARG, "Error reading parse file"); }
    rc = pfp_alloc_t(ctx, &d_parse, pb / 4, true);
    if (rc == PFPB200_OK && pb) rc = pfp_file_to_device(ctx, fd, 0, pb, reinterpret_cast<u8 *>(d_parse));
    close(fd);
    if (rc != PFPB200_OK) return rc;
    const u8 *text = nullptr;
    u64 nt = 0;
    rc = pfpb200_unparse_device(ctx, d_dict, db, 0, d_parse, pb / 4, &text, &nt, ms);
    if (rc != PFPB200_OK) return rc;
    if (outname) snprintf(out, sizeof(out), "%s", outname);
    else snprintf(out, sizeof(out), "%s.out", basename);
    rc = pfp_device_to_file(ctx, out, text, nt);
    if (n_text) *n_text = nt;
    if (n_words) *n_words = (u32)ctx->h_flags[1];
    pfp_release_scratch(ctx);
    return rc;
}
