// pfp_bwtparse.cu -- the stage after the parse: suffix array of the parse, its BWT, the inverted
// list and the permuted .last / .sai streams.  SURVEY.md section 8(f) row 3.
//
// Replaces bwtparse.c of the reference (main(), bwtparse.c:218-322): there sacak_int() (gSACA-K,
// gsa/gsacak.c) sorts the suffixes of T[0..n] (the n parse symbols + the end symbol T[n] = 0) on
// one core, then three loops build BWT / .bwlast / .bwsai (:243-272) and the inverted list by a
// counting sort (:294-298).  Here, for a parse resident in HBM:
//   * suffix array by PREFIX DOUBLING on ranks: round r sorts the pairs (rank of the first h
//     symbols, rank of the next h symbols) with the library's own LSD radix sort, h = 2, 4, 8, ...
//     -- the end symbol is unique and the smallest, so no suffix is a prefix of another and a
//     position past the end never has to be compared (its rank is read as 0).  A round is: keys
//     from ranks (one gather), radix passes over exactly 2 ceil(log2 N) key bits, group flags,
//     a scan, the new ranks.  A suffix's rank is the POSITION where its group starts, so ranks stay
//     valid while only some groups are refined: a suffix alone in its group is final and leaves
//     the active list (Larsson-Sadakane discarding) -- the last rounds sort a few per cent of the
//     suffixes.  It stops when the active list is empty.  A repetitive pan-genome parse needs
//     log2(longest repeat in phrases) + 1 rounds (8 for 100 haplotypes at 0.1 % divergence).
//   * BWT[i] = T[SA[i] - 1], bwlast[i] = last[SA[i] - 2], bwsai[i] = sai[SA[i] - 1] with the special
//     cases of bwtparse.c:246-270: gathers;
//   * ilist = positions i sorted stably by BWT[i] (bwtparse.c:294-298 is a counting sort with the
//     F[] array built from .occ): the same radix sort on (BWT[i], i) -- .occ is not needed.
// Outputs are byte-identical to the reference's .ilist / .bwlast / .bwsai (tests/test_bwtparse_gpu.py
// runs the unmodified bwtparse beside it; tests/golden/golden_bwtparse.npz holds its outputs).
#include "pfp_common.cuh"
#include "pfp_stages.cuh"
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

constexpr int BP_T = 256;

// first round: the pair of the first two symbols; T[n] = 0 is the end symbol (bwtparse.c:114)
__global__ void __launch_bounds__(BP_T) bp_init_k(const u32 *__restrict__ parse, u64 n, u64 *__restrict__ key,
                                                  u32 *__restrict__ val) {
    const u64 i = (u64)blockIdx.x * BP_T + threadIdx.x;
    if (i > n) return;
    const u64 t0 = i < n ? parse[i] : 0u, t1 = i + 1 < n ? parse[i + 1] : 0u;
    key[i] = (t0 << 32) | t1;
    val[i] = (u32)i;
}

// flag[k] = 1 where a new group of equal keys starts (sorted order of the active suffixes)
__global__ void __launch_bounds__(BP_T) bp_flags_k(const u64 *__restrict__ key, u64 M, u8 *__restrict__ flag) {
    const u64 k = (u64)blockIdx.x * BP_T + threadIdx.x;
    if (k >= M) return;
    flag[k] = (k == 0 || key[k] != key[k - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(BP_T) bp_iota_k(u32 *__restrict__ a, u64 n) {
    const u64 i = (u64)blockIdx.x * BP_T + threadIdx.x;
    if (i < n) a[i] = (u32)i;
}

// head[g] = slot (position in the suffix array) where group g of the active list starts
__global__ void __launch_bounds__(BP_T) bp_heads_k(const u8 *__restrict__ flag, const u32 *__restrict__ escan,
                                                   const u32 *__restrict__ slot, u64 M, u32 *__restrict__ head) {
    const u64 k = (u64)blockIdx.x * BP_T + threadIdx.x;
    if (k < M && flag[k]) head[escan[k]] = slot[k];
}

// the sorted active suffixes go back to their slots; a suffix's rank is the slot where its group
// starts (stable under refinement: only the groups that are split change); keep[k] = 0 for a group
// of one -- that suffix is in its final place and leaves the active list
__global__ void __launch_bounds__(BP_T) bp_update_k(const u8 *__restrict__ flag, const u32 *__restrict__ escan,
                                                    const u32 *__restrict__ slot, const u32 *__restrict__ val,
                                                    const u32 *__restrict__ head, u64 M, u32 *__restrict__ sa,
                                                    u32 *__restrict__ rank, u32 *__restrict__ rs,
                                                    u8 *__restrict__ keep) {
    const u64 k = (u64)blockIdx.x * BP_T + threadIdx.x;
    if (k >= M) return;
    const u32 f = flag[k], s = val[k];
    const u32 r = head[escan[k] + f - 1u];
    sa[slot[k]] = s;
    rank[s] = r;
    rs[k] = r;
    keep[k] = (f && (k + 1 == M || flag[k + 1])) ? 0 : 1;
}

// the suffixes that stay active, in order, with the next round's key: (rank, rank h symbols on)
__global__ void __launch_bounds__(BP_T) bp_compact_k(const u8 *__restrict__ keep, const u32 *__restrict__ kscan,
                                                     const u32 *__restrict__ slot, const u32 *__restrict__ val,
                                                     const u32 *__restrict__ rs, const u32 *__restrict__ rank, u64 M,
                                                     u64 N, u64 h, int b, u32 *__restrict__ slot2,
                                                     u32 *__restrict__ val2, u64 *__restrict__ key2) {
    const u64 k = (u64)blockIdx.x * BP_T + threadIdx.x;
    if (k >= M || !keep[k]) return;
    const u32 p = kscan[k], s = val[k];
    const u64 j = (u64)s + h;
    slot2[p] = slot[k];
    val2[p] = s;
    key2[p] = ((u64)rs[k] << b) | (j < N ? (u64)rank[j] + 1u : 0u);
}

// BWT of the parse and the permuted .last / .sai (bwtparse.c:243-272); SA[0] = n by construction
__global__ void __launch_bounds__(BP_T) bp_emit_k(const u32 *__restrict__ sa, const u32 *__restrict__ parse,
                                                  const u8 *__restrict__ last, const u8 *__restrict__ sai, u64 n,
                                                  u64 *__restrict__ bwt_key, u32 *__restrict__ pos,
                                                  u8 *__restrict__ bwlast, u8 *__restrict__ bwsai,
                                                  unsigned long long *__restrict__ flags) {
    const u64 i = (u64)blockIdx.x * BP_T + threadIdx.x;
    if (i > n) return;
    const u64 s = sa[i];
    if ((i == 0) != (s == n) || (s == 0 && i != 1)) atomicOr(flags, PFP_ERRBIT_INTERNAL);   // the asserts of :244,:250
    u32 sym = 0;
    u8 lc = 0;
    if (s != 0) {
        sym = parse[s - 1];
        lc = s == 1 ? last[n - 1] : last[s - 2];
    }
    bwt_key[i] = sym;
    pos[i] = (u32)i;
    bwlast[i] = lc;
    if (bwsai) {
#pragma unroll
        for (int k = 0; k < PFP_IBYTES; k++) bwsai[i * PFP_IBYTES + k] = s ? sai[(s - 1) * PFP_IBYTES + k] : (u8)0;
    }
}

__global__ void __launch_bounds__(BP_T) bp_max_k(const u32 *__restrict__ parse, u64 n, u32 *__restrict__ out) {
    u32 m = 0;
    for (u64 i = (u64)blockIdx.x * BP_T + threadIdx.x; i < n; i += (u64)gridDim.x * BP_T) m = max(m, parse[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// The outputs of a parse on this context stay valid (they are the natural inputs here); only the
// outputs of the previous bwtparse call are given back.
static int bp_begin(pfpb200_ctx *ctx) {
    PFP_CUDA(ctx, cudaSetDevice(ctx->device));
    pfp_release_scratch(ctx);
    for (int k = 0; k < 3; k++) {
        void *p = ctx->bp_out[k];
        ctx->bp_out[k] = nullptr;
        if (!p) continue;
        for (size_t i = 0; i < ctx->held.size(); i++)
            if (ctx->held[i] == p) {
                ctx->held[i] = ctx->held.back();
                ctx->held.pop_back();
                ctx->scratch.push_back(p);            // released with the scratch just below
                break;
            }
    }
    pfp_release_scratch(ctx);
    ctx->err[0] = 0;
    PFP_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, 3 * sizeof(u64), ctx->stream));
    return PFPB200_OK;
}

static int bits_for(u64 v) {   // bits needed to write values 0 .. v
    int b = 1;
    while (b < 64 && (v >> b) != 0) b++;
    return b;
}

// d_parse[n], d_last[n], d_sai[5n] (or null) on the device -> held device outputs
static int bwtparse_device_impl(pfpb200_ctx *ctx, const u32 *d_parse, u64 n, const u8 *d_last, const u8 *d_sai,
                                pfpb200_bwtparse_result *res) {
    memset(res, 0, sizeof(*res));
    if (n < 2) return pfp_fail(ctx, PFPB200_E_ARG, "bwtparse: the parse must hold at least two phrases");   // assert(n>1), :241
    if (n >= 0xFFFFFFFEull) return pfp_fail(ctx, PFPB200_E_LIMIT, "Input containing more than 2^32-2 phrases!");   // :101
    const u64 N = n + 1;
    const u32 nb = pfp_blocks(N, BP_T);
    PfpEvents evs(3);
    if (!evs.ok) return pfp_fail(ctx, PFPB200_E_CUDA, "cudaEventCreate failed");
    const u32 launches0 = ctx->launches;
    PFP_CUDA(ctx, cudaEventRecord(evs[0], ctx->stream));
    u64 *k0 = nullptr, *k1 = nullptr;
    u32 *v0 = nullptr, *v1 = nullptr, *rank = nullptr, *rs = nullptr, *escan = nullptr, *sa = nullptr;
    u32 *slot = nullptr, *slot2 = nullptr, *head = nullptr;
    u8 *flag = nullptr, *keep = nullptr;
    PFP_TRY(pfp_alloc_t(ctx, &k0, N));
    PFP_TRY(pfp_alloc_t(ctx, &k1, N));
    PFP_TRY(pfp_alloc_t(ctx, &v0, N));
    PFP_TRY(pfp_alloc_t(ctx, &v1, N));
    PFP_TRY(pfp_alloc_t(ctx, &sa, N));
    PFP_TRY(pfp_alloc_t(ctx, &rank, N));
    PFP_TRY(pfp_alloc_t(ctx, &rs, N));
    PFP_TRY(pfp_alloc_t(ctx, &escan, N));
    PFP_TRY(pfp_alloc_t(ctx, &slot, N));
    PFP_TRY(pfp_alloc_t(ctx, &slot2, N));
    PFP_TRY(pfp_alloc_t(ctx, &head, N));
    PFP_TRY(pfp_alloc_t(ctx, &flag, N));
    PFP_TRY(pfp_alloc_t(ctx, &keep, N));
    // largest symbol: key bits of the first round and of the inverted list
    u32 *d_max = reinterpret_cast<u32 *>(&ctx->d_flags[2]);
    PFP_CUDA(ctx, cudaMemsetAsync(&ctx->d_flags[1], 0, 2 * sizeof(u64), ctx->stream));
    bp_max_k<<<ctx->sm_count * 4, BP_T, 0, ctx->stream>>>(d_parse, n, d_max);
    PFP_LAUNCHED(ctx);
    PFP_CUDA(ctx, cudaMemcpyAsync(&ctx->h_flags[2], &ctx->d_flags[2], sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    bp_init_k<<<nb, BP_T, 0, ctx->stream>>>(d_parse, n, k0, v0);
    PFP_LAUNCHED(ctx);
    bp_iota_k<<<nb, BP_T, 0, ctx->stream>>>(slot, N);
    PFP_LAUNCHED(ctx);
    PFP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const u32 kmax = (u32)ctx->h_flags[2];
    const int sym_bits = bits_for(kmax);
    const int b = bits_for(N);                            // bits of a rank + 1
    u64 *ks = nullptr;
    u32 *vs = nullptr;
    // round 1: all suffixes by their first two symbols.  The low half sorts over sym_bits, the high half starts at bit 32
    PFP_TRY(pfp_radix_sort_pairs(ctx, k0, v0, k1, v1, N, 0, sym_bits, &ks, &vs));
    {
        u64 *ko = ks == k0 ? k1 : k0;
        u32 *vo = vs == v0 ? v1 : v0;
        PFP_TRY(pfp_radix_sort_pairs(ctx, ks, vs, ko, vo, N, 32, 32 + sym_bits, &ks, &vs));
    }
    u32 rounds = 1;
    u64 M = N;                                            // suffixes still in a group of two or more
    u32 *d_cnt = reinterpret_cast<u32 *>(&ctx->d_flags[1]);
    for (u64 h = 2;; h *= 2, rounds++) {
        // (ks, vs)[0..M): the active suffixes sorted by (rank, rank of the next h / 2 symbols); slot[k] = where the k-th goes
        const u32 mb = pfp_blocks(M, BP_T);
        bp_flags_k<<<mb, BP_T, 0, ctx->stream>>>(ks, M, flag);
        PFP_LAUNCHED(ctx);
        PFP_TRY(pfp_exclusive_scan_u8_u32(ctx, flag, escan, M, nullptr));
        bp_heads_k<<<mb, BP_T, 0, ctx->stream>>>(flag, escan, slot, M, head);
        PFP_LAUNCHED(ctx);
        bp_update_k<<<mb, BP_T, 0, ctx->stream>>>(flag, escan, slot, vs, head, M, sa, rank, rs, keep);
        PFP_LAUNCHED(ctx);
        PFP_TRY(pfp_exclusive_scan_u8_u32(ctx, keep, escan, M, d_cnt));
        PFP_CUDA(ctx, cudaMemcpyAsync(&ctx->h_flags[1], &ctx->d_flags[1], sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        PFP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const u64 M2 = (u32)ctx->h_flags[1];
        if (M2 == 0) break;                               // every suffix has its own rank: sa[] is the suffix array
        if (h >= 2 * N) return pfp_fail(ctx, PFPB200_E_INTERNAL, "bwtparse: prefix doubling did not converge");
        u64 *ko = ks == k0 ? k1 : k0;
        u32 *vo = vs == v0 ? v1 : v0;
        bp_compact_k<<<mb, BP_T, 0, ctx->stream>>>(keep, escan, slot, vs, rs, rank, M, N, h, b, slot2, vo, ko);
        PFP_LAUNCHED(ctx);
        { u32 *t = slot; slot = slot2; slot2 = t; }
        M = M2;
        PFP_TRY(pfp_radix_sort_pairs(ctx, ko, vo, ks, vs, M, 0, 2 * b, &ks, &vs));
    }
    PFP_CUDA(ctx, cudaEventRecord(evs[1], ctx->stream));
    // BWT, .bwlast, .bwsai, then the inverted list = positions sorted stably by BWT symbol
    u32 *ilist = nullptr;
    u8 *bwlast = nullptr, *bwsai = nullptr;
    PFP_TRY(pfp_alloc_t(ctx, &bwlast, N, true));
    if (d_sai) PFP_TRY(pfp_alloc_t(ctx, &bwsai, N * PFP_IBYTES, true));
    PFP_TRY(pfp_alloc_t(ctx, &ilist, N, true));
    u64 *bk = k0;                                          // the sort buffers are free: the suffix array has its own
    u32 *bv = v0;
    bp_emit_k<<<nb, BP_T, 0, ctx->stream>>>(sa, d_parse, d_last, d_sai, n, bk, bv, bwlast, bwsai,
                                            reinterpret_cast<unsigned long long *>(&ctx->d_flags[0]));
    PFP_LAUNCHED(ctx);
    u64 *ik = nullptr;
    u32 *iv = nullptr;
    PFP_TRY(pfp_radix_sort_pairs(ctx, bk, bv, k1, v1, N, 0, sym_bits, &ik, &iv));
    PFP_CUDA(ctx, cudaMemcpyAsync(ilist, iv, N * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
    PFP_CUDA(ctx, cudaEventRecord(evs[2], ctx->stream));
    PFP_CUDA(ctx, cudaMemcpyAsync(&ctx->h_flags[0], &ctx->d_flags[0], sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    PFP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_flags[0] & PFP_ERRBIT_INTERNAL)
        return pfp_fail(ctx, PFPB200_E_INTERNAL, "bwtparse: the parse does not start with its smallest, unique phrase");
    ctx->bp_out[0] = ilist; ctx->bp_out[1] = bwlast; ctx->bp_out[2] = bwsai;
    res->ilist = ilist;
    res->bwlast = bwlast;
    res->bwsai = bwsai;
    res->n_out = N;
    res->alphabet = (u64)kmax + 1;
    res->rounds = rounds;
    res->launches = ctx->launches - launches0;
    cudaEventElapsedTime(&res->ms_sa, evs[0], evs[1]);
    cudaEventElapsedTime(&res->ms_lists, evs[1], evs[2]);
    res->ms_total = res->ms_sa + res->ms_lists;
    return PFPB200_OK;
}

extern "C" int pfpb200_bwtparse_device(pfpb200_ctx *ctx, const uint32_t *d_parse, uint64_t n_phrases,
                                       const uint8_t *d_last, const uint8_t *d_sai,
                                       pfpb200_bwtparse_result *res) {
    if (!ctx || !d_parse || !d_last || !res) return PFPB200_E_ARG;
    PFP_TRY(bp_begin(ctx));
    const int rc = bwtparse_device_impl(ctx, d_parse, n_phrases, d_last, d_sai, res);
    pfp_release_scratch(ctx);
    return rc;
}

extern "C" int pfpb200_bwtparse_host(pfpb200_ctx *ctx, const uint32_t *parse, uint64_t n, const uint8_t *last,
                                     const uint8_t *sai, uint32_t *ilist, uint8_t *bwlast, uint8_t *bwsai,
                                     pfpb200_bwtparse_result *res) {
    if (!ctx || !parse || !last || !ilist || !bwlast || !res || (sai && !bwsai)) return PFPB200_E_ARG;
    PFP_TRY(bp_begin(ctx));
    u32 *d_parse = nullptr;
    u8 *d_last = nullptr, *d_sai = nullptr;
    int rc = pfp_alloc_t(ctx, &d_parse, n);
    if (rc == PFPB200_OK) rc = pfp_alloc_t(ctx, &d_last, n);
    if (rc == PFPB200_OK && sai) rc = pfp_alloc_t(ctx, &d_sai, n * PFP_IBYTES);
    if (rc != PFPB200_OK) { pfp_release_scratch(ctx); return rc; }
    auto copy = [&](void *d, const void *s, size_t bytes, cudaMemcpyKind k) {
        return bytes == 0 || cudaMemcpyAsync(d, s, bytes, k, ctx->stream) == cudaSuccess;
    };
    bool ok = copy(d_parse, parse, n * sizeof(u32), cudaMemcpyHostToDevice) &&
              copy(d_last, last, n, cudaMemcpyHostToDevice) &&
              (!sai || copy(d_sai, sai, n * PFP_IBYTES, cudaMemcpyHostToDevice));
    if (!ok) { pfp_release_scratch(ctx); return pfp_fail(ctx, PFPB200_E_CUDA, "bwtparse: host to device copy failed"); }
    rc = bwtparse_device_impl(ctx, d_parse, n, d_last, d_sai, res);
    if (rc == PFPB200_OK) {
        ok = copy(ilist, res->ilist, res->n_out * sizeof(u32), cudaMemcpyDeviceToHost) &&
             copy(bwlast, res->bwlast, res->n_out, cudaMemcpyDeviceToHost) &&
             (!sai || copy(bwsai, res->bwsai, res->n_out * PFP_IBYTES, cudaMemcpyDeviceToHost)) &&
             cudaStreamSynchronize(ctx->stream) == cudaSuccess;
        if (!ok) rc = pfp_fail(ctx, PFPB200_E_CUDA, "bwtparse: device to host copy failed");
    }
    pfp_release_scratch(ctx);
    return rc;
}

// ---- files: <base>.parse, <base>[.<i>].last, <base>[.<i>].sai -> <base>.ilist .bwlast .bwsai ----------------
static int open_size(const char *name, u64 *size) {
    const int fd = open(name, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return -1; }
    *size = (u64)st.st_size;
    return fd;
}

// the .last / .sai stream of `want` bytes, in one file or in nseg segment files <base>.<i>.<ext>
// (mopen_aux_file / mfread, utils.c:57-110), streamed to d_dst
static int segments_to_device(pfpb200_ctx *ctx, const char *base, const char *ext, int nseg, u64 want, u8 *d_dst) {
    char name[4096];
    u64 got = 0;
    for (int i = 0; i < (nseg > 0 ? nseg : 1); i++) {
        if (nseg > 0) snprintf(name, sizeof(name), "%s.%d.%s", base, i, ext);
        else snprintf(name, sizeof(name), "%s.%s", base, ext);
        u64 sz = 0;
        const int fd = open_size(name, &sz);
        if (fd < 0) return pfp_fail(ctx, PFPB200_E_IO, "cannot open %s: %s", name, strerror(errno));
        if (got + sz > want) sz = want - got;
        const int rc = sz ? pfp_file_to_device(ctx, fd, 0, sz, d_dst + got) : PFPB200_OK;
        close(fd);
        if (rc != PFPB200_OK) return rc;
        got += sz;
    }
    if (got != want) return pfp_fail(ctx, PFPB200_E_IO, "%s: %llu bytes, %llu expected", ext, (unsigned long long)got,
                                     (unsigned long long)want);
    return PFPB200_OK;
}

extern "C" int pfpb200_bwtparse_file(pfpb200_ctx *ctx, const char *basename, int sa_info, int nseg,
                                     pfpb200_bwtparse_result *res) {
    if (!ctx || !basename || !res) return PFPB200_E_ARG;
    PFP_TRY(bp_begin(ctx));
    char name[4096];
    snprintf(name, sizeof(name), "%s.parse", basename);
    u64 bytes = 0;
    const int fd = open_size(name, &bytes);
    if (fd < 0) return pfp_fail(ctx, PFPB200_E_IO, "cannot open %s: %s", name, strerror(errno));
    if (bytes % 4 != 0) {
        close(fd);
        return pfp_fail(ctx, PFPB200_E_ARG, "Invalid input file: size not multiple of 4");      // bwtparse.c:81
    }
    const u64 n = bytes / 4;
    u32 *d_parse = nullptr;
    u8 *d_last = nullptr, *d_sai = nullptr;
    int rc = pfp_alloc_t(ctx, &d_parse, n);
    if (rc == PFPB200_OK && n) rc = pfp_file_to_device(ctx, fd, 0, bytes, reinterpret_cast<u8 *>(d_parse));
    close(fd);
    if (rc == PFPB200_OK) rc = pfp_alloc_t(ctx, &d_last, n);
    if (rc == PFPB200_OK) rc = segments_to_device(ctx, basename, "last", nseg, n, d_last);
    if (rc == PFPB200_OK && sa_info) rc = pfp_alloc_t(ctx, &d_sai, n * PFP_IBYTES);
    if (rc == PFPB200_OK && sa_info) rc = segments_to_device(ctx, basename, "sai", nseg, n * PFP_IBYTES, d_sai);
    if (rc == PFPB200_OK) rc = bwtparse_device_impl(ctx, d_parse, n, d_last, d_sai, res);
    if (rc == PFPB200_OK) {
        snprintf(name, sizeof(name), "%s.bwlast", basename);
        rc = pfp_device_to_file(ctx, name, res->bwlast, res->n_out);
    }
    if (rc == PFPB200_OK && sa_info) {
        snprintf(name, sizeof(name), "%s.bwsai", basename);
        rc = pfp_device_to_file(ctx, name, res->bwsai, res->n_out * PFP_IBYTES);
    }
    if (rc == PFPB200_OK) {
        snprintf(name, sizeof(name), "%s.ilist", basename);
        rc = pfp_device_to_file(ctx, name, res->ilist, res->n_out * sizeof(u32));
    }
    pfp_release_scratch(ctx);
    return rc;
}
