// pfp_rank.cu -- K4 lexicographic ranking of the distinct words, .dict/.occ emission, K5 remap.
//
// Replaces std::sort with pstringCompare (newscan.cpp:387-390,636), writeDictOcc (:394-441)
// and remapParse (:443-466).  Order = unsigned-byte lexicographic, as std::string compares.
//
// The distinct words sit in a pool, each zero-padded to 8-byte words.  Ranking is an MSD
// refinement on 8-byte big-endian chunks: round r sorts the still-tied words by
// (tie-group, chunk r) with the LSD radix sort (chunk bits first, then the group id, stable),
// writes them back into their group's slots and splits groups where the chunk changes.  A
// word's padding is 0x00, below every text byte (> 0x02), so shorter-is-smaller falls out.
#include "pfp_common.cuh"
#include "pfp_stages.cuh"

__device__ __forceinline__ u64 bswap64(u64 v) {
    u32 lo = (u32)v, hi = (u32)(v >> 32);
    return ((u64)__byte_perm(lo, 0, 0x0123) << 32) | (u64)__byte_perm(hi, 0, 0x0123);
}

__device__ __forceinline__ u64 word_key(const u64 *__restrict__ pool, const u64 *__restrict__ uoff,
                                        const u32 *__restrict__ uwords, u32 u, u32 r) {
    return (r < uwords[u]) ? bswap64(__ldg(pool + uoff[u] + r)) : 0ull;
}

__global__ void rank_keys0_k(const u64 *pool, const u64 *uoff, const u32 *uwords, u64 d, u64 *keys,
                             u32 *vals) {
    u64 u = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= d) return;
    keys[u] = word_key(pool, uoff, uwords, (u32)u, 0);
    vals[u] = (u32)u;
}

__global__ void rank_heads0_k(const u64 *__restrict__ ks, u64 d, u8 *__restrict__ head) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    head[i] = (i == 0 || ks[i] != ks[i - 1]) ? 1 : 0;
}

__global__ void rank_active_k(const u8 *__restrict__ head, u64 d, u8 *__restrict__ act,
                              u8 *__restrict__ gh) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    // first position at or before i that is a head == start of my group; I am alone iff I am a
    // head and my successor is one too
    bool single = head[i] && (i + 1 == d || head[i + 1]);
    act[i] = single ? 0 : 1;
    gh[i] = (!single && head[i]) ? 1 : 0;
}

__global__ void rank_compact_k(const u8 *__restrict__ act, const u8 *__restrict__ gh,
                               const u32 *__restrict__ ascan, const u32 *__restrict__ gscan,
                               const u32 *__restrict__ ord, u64 d, u32 r, const u64 *pool,
                               const u64 *uoff, const u32 *uwords, u32 *__restrict__ apos,
                               u64 *__restrict__ keys, u32 *__restrict__ vals,
                               u32 *__restrict__ gid_of_u) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d || !act[i]) return;
    u32 j = ascan[i];
    u32 u = ord[i];
    apos[j] = (u32)i;
    keys[j] = word_key(pool, uoff, uwords, u, r);
    vals[j] = u;
    gid_of_u[u] = gscan[i] + gh[i] - 1;
}

__global__ void rank_gidkeys_k(const u32 *__restrict__ vs, const u32 *__restrict__ gid_of_u, u64 m,
                               u64 *__restrict__ keys2) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) keys2[j] = gid_of_u[vs[j]];
}

__global__ void rank_writeback_k(const u32 *__restrict__ apos, const u64 *__restrict__ gk,
                                 const u32 *__restrict__ vs, u64 m, u32 r, const u64 *pool,
                                 const u64 *uoff, const u32 *uwords, u32 *__restrict__ ord,
                                 u8 *__restrict__ head) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    u32 i = apos[j];
    u32 u = vs[j];
    ord[i] = u;
    if (j > 0 && gk[j] == gk[j - 1]) {
        u32 up = vs[j - 1];
        if (word_key(pool, uoff, uwords, u, r) != word_key(pool, uoff, uwords, up, r)) head[i] = 1;
    }
}

int pfp_rank_stage(pfpb200_ctx *ctx, const DictArrays &D, u32 **order, u32 *rounds) {
    const int TB = 256;
    const u64 d = D.d;
    u64 *k0 = nullptr, *k1 = nullptr, *ks = nullptr;
    u32 *v0 = nullptr, *v1 = nullptr, *vs = nullptr;
    u32 *ord = nullptr;
    PFP_TRY(pfp_alloc_t(ctx, &k0, d));
    PFP_TRY(pfp_alloc_t(ctx, &k1, d));
    PFP_TRY(pfp_alloc_t(ctx, &v0, d));
    PFP_TRY(pfp_alloc_t(ctx, &v1, d));
    PFP_TRY(pfp_alloc_t(ctx, &ord, d));
    u8 *head = nullptr, *act = nullptr, *gh = nullptr;
    PFP_TRY(pfp_alloc_t(ctx, &head, d));
    const u32 nbd = pfp_blocks(d, TB);
    rank_keys0_k<<<nbd, TB, 0, ctx->stream>>>(D.pool, D.uoff, D.uwords, d, k0, v0);
    PFP_LAUNCHED(ctx);
    PFP_TRY(pfp_radix_sort_pairs(ctx, k0, v0, k1, v1, d, 0, 64, &ks, &vs));
    rank_heads0_k<<<nbd, TB, 0, ctx->stream>>>(ks, d, head);
    PFP_LAUNCHED(ctx);
    PFP_CUDA(ctx, cudaMemcpyAsync(ord, vs, d * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));

    u32 *ascan = nullptr, *gscan = nullptr, *apos = nullptr, *gid_of_u = nullptr;
    u32 r = 1;
    const u32 max_rounds = (D.max_len + 7) / 8 + 1;
    bool allocated = false;
    for (;; r++) {
        if (d < 2) break;
        if (!allocated) {
            PFP_TRY(pfp_alloc_t(ctx, &act, d));
            PFP_TRY(pfp_alloc_t(ctx, &gh, d));
            PFP_TRY(pfp_alloc_t(ctx, &ascan, d));
            PFP_TRY(pfp_alloc_t(ctx, &gscan, d));
            PFP_TRY(pfp_alloc_t(ctx, &apos, d));
            PFP_TRY(pfp_alloc_t(ctx, &gid_of_u, d));
            allocated = true;
        }
        rank_active_k<<<nbd, TB, 0, ctx->stream>>>(head, d, act, gh);
        PFP_LAUNCHED(ctx);
        PFP_CUDA(ctx, cudaMemsetAsync(&ctx->d_flags[1], 0, 2 * sizeof(u64), ctx->stream));
        PFP_TRY(pfp_exclusive_scan_u8_u32(ctx, act, ascan, d, reinterpret_cast<u32 *>(&ctx->d_flags[1])));
        PFP_TRY(pfp_exclusive_scan_u8_u32(ctx, gh, gscan, d, reinterpret_cast<u32 *>(&ctx->d_flags[2])));
        PFP_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 3 * sizeof(u64),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        PFP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        u64 m = (u32)ctx->h_flags[1], ng = (u32)ctx->h_flags[2];
        if (m == 0) break;
        if (r > max_rounds)
            return pfp_fail(ctx, PFPB200_E_INTERNAL,
                            "ranking did not separate %llu words after %u rounds (duplicate words?)",
                            (unsigned long long)m, r);
        rank_compact_k<<<nbd, TB, 0, ctx->stream>>>(act, gh, ascan, gscan, ord, d, r, D.pool, D.uoff,
                                                    D.uwords, apos, k0, v0, gid_of_u);
        PFP_LAUNCHED(ctx);
        PFP_TRY(pfp_radix_sort_pairs(ctx, k0, v0, k1, v1, m, 0, 64, &ks, &vs));
        // second key: the tie group, stable, only as many bits as there are groups
        u64 *g_in = (ks == k0) ? k1 : k0;
        u32 *v_other = (vs == v0) ? v1 : v0;
        const u32 nbm = pfp_blocks(m, TB);
        rank_gidkeys_k<<<nbm, TB, 0, ctx->stream>>>(vs, gid_of_u, m, g_in);
        PFP_LAUNCHED(ctx);
        int bits = 1;
        while (bits < 32 && (1ull << bits) < ng) bits++;
        u64 *gs = nullptr;
        u32 *vs2 = nullptr;
        PFP_TRY(pfp_radix_sort_pairs(ctx, g_in, vs, ks, v_other, m, 0, bits, &gs, &vs2));
        rank_writeback_k<<<nbm, TB, 0, ctx->stream>>>(apos, gs, vs2, m, r, D.pool, D.uoff, D.uwords,
                                                      ord, head);
        PFP_LAUNCHED(ctx);
    }
    *rounds = r;
    *order = ord;
    PFP_TRY(pfp_free_now(ctx, k0));
    PFP_TRY(pfp_free_now(ctx, k1));
    PFP_TRY(pfp_free_now(ctx, v0));
    PFP_TRY(pfp_free_now(ctx, v1));
    PFP_TRY(pfp_free_now(ctx, head));
    PFP_TRY(pfp_free_now(ctx, act));
    PFP_TRY(pfp_free_now(ctx, gh));
    PFP_TRY(pfp_free_now(ctx, ascan));
    PFP_TRY(pfp_free_now(ctx, gscan));
    PFP_TRY(pfp_free_now(ctx, apos));
    PFP_TRY(pfp_free_now(ctx, gid_of_u));
    return PFPB200_OK;
}

// ---- .dict / .occ -----------------------------------------------------------------------------------
__device__ __forceinline__ void out_span(const u64 *pool, u64 off, u32 len, u32 strip_w, u32 *skip,
                                         u32 *outlen) {
    // -c mode (newscan.cpp:410-413): drop the last w bytes and a leading 0x02
    *skip = 0;
    *outlen = len;
    if (strip_w) {
        u32 l = len - strip_w;
        if ((u8)(pool[off] & 0xFF) == PFP_DOLLAR) { *skip = 1; l -= 1; }
        *outlen = l;
    }
}

__global__ void dict_layout_k(const u32 *__restrict__ ord, u64 d, const DictArrays D, u32 strip_w,
                              u32 *__restrict__ rank_of_uid, u32 *__restrict__ occ,
                              u32 *__restrict__ dl) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    u32 u = ord[i];
    rank_of_uid[u] = (u32)i + 1;                     // 1-based (newscan.cpp:405,436)
    occ[i] = D.count[u];                             // newscan.cpp:433
    u32 skip, outlen;
    out_span(D.pool, D.uoff[u], D.ulen[u], strip_w, &skip, &outlen);
    dl[i] = outlen + 1;                              // + EndOfWord (newscan.cpp:416)
}

constexpr int DC_T = 256;
constexpr int DC_GROUP = 8;

__global__ void __launch_bounds__(DC_T) dict_copy_k(const u32 *__restrict__ ord,
                                                    const u64 *__restrict__ doff, u64 d,
                                                    const DictArrays D, u32 strip_w, u64 total,
                                                    u8 *__restrict__ dict) {
    const u32 li = threadIdx.x & (DC_GROUP - 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) dict[total] = (u8)PFP_END_OF_DICT;   // :438
    for (u64 i = (u64)blockIdx.x * (DC_T / DC_GROUP) + (threadIdx.x / DC_GROUP); i < d;
         i += (u64)gridDim.x * (DC_T / DC_GROUP)) {
        u32 u = ord[i];
        u32 skip, outlen;
        u64 off = D.uoff[u];
        out_span(D.pool, off, D.ulen[u], strip_w, &skip, &outlen);
        const u8 *src = reinterpret_cast<const u8 *>(D.pool + off) + skip;
        u8 *dst = dict + doff[i];
        if (skip == 0) {
            // aligned 8-byte source words, byte stores (destination is unaligned)
            u32 nw = (outlen + 7) >> 3;
            for (u32 k = li; k < nw; k += DC_GROUP) {
                u64 v = __ldg(D.pool + off + k);
                u32 nb = outlen - 8 * k < 8 ? outlen - 8 * k : 8;
                for (u32 b = 0; b < nb; b++) dst[8 * k + b] = (u8)(v >> (8 * b));
            }
        } else {
            for (u32 b = li; b < outlen; b += DC_GROUP) dst[b] = src[b];
        }
        if (li == 0) dst[outlen] = (u8)PFP_END_OF_WORD;
    }
}

int pfp_dict_stage(pfpb200_ctx *ctx, const DictArrays &D, const u32 *order, u32 strip_w, u8 **dict,
                   u64 *dict_bytes, u32 **occ, u32 **rank_of_uid) {
    const int TB = 256;
    const u64 d = D.d;
    u32 *dl = nullptr;
    u64 *doff = nullptr;
    PFP_TRY(pfp_alloc_t(ctx, occ, d, true));
    PFP_TRY(pfp_alloc_t(ctx, rank_of_uid, d));
    PFP_TRY(pfp_alloc_t(ctx, &dl, d));
    PFP_TRY(pfp_alloc_t(ctx, &doff, d));
    dict_layout_k<<<pfp_blocks(d, TB), TB, 0, ctx->stream>>>(order, d, D, strip_w, *rank_of_uid, *occ, dl);
    PFP_LAUNCHED(ctx);
    PFP_TRY(pfp_exclusive_scan_u32_u64(ctx, dl, doff, d, &ctx->d_flags[1]));
    PFP_CUDA(ctx, cudaMemcpyAsync(&ctx->h_flags[1], &ctx->d_flags[1], sizeof(u64),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    PFP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    u64 total = ctx->h_flags[1];
    *dict_bytes = total + 1;
    PFP_TRY(pfp_alloc_t(ctx, dict, (size_t)(total + 1), true));
    u64 want = (d + (DC_T / DC_GROUP) - 1) / (DC_T / DC_GROUP);
    u64 maxb = (u64)ctx->sm_count * 32;
    u32 nb = (u32)(want < maxb ? want : maxb);
    if (nb == 0) nb = 1;
    dict_copy_k<<<nb, DC_T, 0, ctx->stream>>>(order, doff, d, D, strip_w, total, *dict);
    PFP_LAUNCHED(ctx);
    PFP_TRY(pfp_free_now(ctx, dl));
    PFP_TRY(pfp_free_now(ctx, doff));
    return PFPB200_OK;
}

// ---- K5 -----------------------------------------------------------------------------------------------
__global__ void remap_k(const u32 *__restrict__ uid, const u32 *__restrict__ rank_of_uid, u64 P,
                        u32 *__restrict__ parse) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < P) parse[j] = __ldg(rank_of_uid + uid[j]);    // newscan.cpp:456-458
}

int pfp_remap_stage(pfpb200_ctx *ctx, const u32 *uid, const u32 *rank_of_uid, u64 P, u32 *parse) {
    remap_k<<<pfp_blocks(P, 256), 256, 0, ctx->stream>>>(uid, rank_of_uid, P, parse);
    PFP_LAUNCHED(ctx);
    return PFPB200_OK;
}
