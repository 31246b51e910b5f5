/* pfp_oracle.c -- TEST INFRASTRUCTURE ONLY (see pfp_oracle.h).
 *
 * Plain-C CPU restatement of the prefix-free parsing stage of alshai/Big-BWT.
 * Each function names the reference lines it follows.  The restatement keeps the
 * reference's arithmetic (same moduli, same update order) but not its data
 * structures: the text is held once in memory with its virtual 0x02 borders, phrases
 * are (start,len) views into it, and the dictionary is an open-addressing table keyed
 * by the reference's own 64-bit phrase hash with a byte compare on every hit, which is
 * what `-P` (newscan.cpp:256-269) amounts to and leaves the output files unchanged.
 */
#define _POSIX_C_SOURCE 200809L
#include "pfp_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define DOLLAR 2u      /* utils.h:5 */
#define END_OF_WORD 1u /* utils.h:6 */
#define END_OF_DICT 0u /* utils.h:7 */
#define IBYTES 5       /* utils.h:10 */

static const uint64_t WIN_PRIME = 1999999973ULL;          /* newscan.cpp:172 */
static const uint64_t WORD_PRIME = 27162335252586509ULL;  /* newscan.cpp:232 */

static double now_sec(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- rolling window, newscan.cpp:168-216 ---------------------------------------- */
typedef struct {
    uint32_t w;
    int *ring;          /* last w symbols, zero before w symbols were seen (:188-192) */
    uint64_t hash, seen;
    uint64_t top_pow;   /* 256^(w-1) mod WIN_PRIME (:180-181) */
} roll_t;

static int roll_init(roll_t *r, uint32_t w) {
    r->w = w;
    r->ring = (int *)calloc(w, sizeof(int));
    if (!r->ring) return -1;
    r->hash = r->seen = 0;
    r->top_pow = 1;
    for (uint32_t i = 1; i < w; i++) r->top_pow = (r->top_pow * 256) % WIN_PRIME;
    return 0;
}

/* newscan.cpp:194-202: drop the oldest symbol, shift, add the new one */
static uint64_t roll_push(roll_t *r, int c) {
    uint32_t slot = (uint32_t)(r->seen++ % r->w);
    r->hash += WIN_PRIME - ((uint64_t)r->ring[slot] * r->top_pow) % WIN_PRIME;
    r->hash = (256 * r->hash + (uint64_t)c) % WIN_PRIME;
    r->ring[slot] = c;
    return r->hash;
}

uint64_t pfp_oracle_window_hash(const uint8_t *s, uint32_t w) {
    roll_t r;
    if (roll_init(&r, w)) return 0;
    uint64_t h = 0;
    for (uint32_t i = 0; i < w; i++) h = roll_push(&r, s[i]);
    free(r.ring);
    return h;
}

/* ---- 64-bit phrase hash, newscan.cpp:229-239 ------------------------------------- */
uint64_t pfp_oracle_kr_hash(const uint8_t *s, uint64_t len) {
    uint64_t h = 0;
    for (uint64_t k = 0; k < len; k++) h = (256 * h + s[k]) % WORD_PRIME;
    return h;
}

/* number of leading bytes that are valid text (> 0x02): the sequential scanner stops at
 * the first invalid byte and keeps what it has (newscan.cpp:341,364) */
static uint64_t valid_prefix(const uint8_t *t, uint64_t n) {
    uint64_t i = 0;
    while (i < n && t[i] > DOLLAR) i++;
    return i;
}

/* ---- trigger positions: newscan.cpp:363-371 with the |word|<=w rule of :247-248 --- */
static uint64_t scan_triggers(const uint8_t *text, uint64_t n, uint32_t w, uint32_t p,
                              uint32_t flags, uint64_t **out_e) {
    roll_t r;
    if (roll_init(&r, w)) return (uint64_t)-1;
    uint64_t cap = n / (p > 4 ? p / 2 : 2) + 16, k = 0;
    uint64_t *e = (uint64_t *)malloc(cap * sizeof(uint64_t));
    if (!e) { free(r.ring); return (uint64_t)-1; }
    /* sequential scanner: a word of length <= w is not cut (word = 0x02 + e+1 symbols), so
     * e >= w-1; helper threads test parsed > w instead, so e >= w (pscan.hpp:93) */
    uint64_t first_ok = (flags & PFP_ORACLE_THREADED_RULE) ? w : (uint64_t)w - 1;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t h = roll_push(&r, text[i]);
        if (h % p == 0 && i >= first_ok) {
            if (k == cap) {
                cap *= 2;
                uint64_t *ne = (uint64_t *)realloc(e, cap * sizeof(uint64_t));
                if (!ne) { free(e); free(r.ring); return (uint64_t)-1; }
                e = ne;
            }
            e[k++] = i;
        }
    }
    free(r.ring);
    *out_e = e;
    return k;
}

uint64_t pfp_oracle_triggers(const uint8_t *text, uint64_t n, uint32_t w, uint32_t p,
                             uint32_t flags, uint64_t *out, uint64_t cap) {
    if (w < 1 || p < 1) return 0;
    n = valid_prefix(text, n);
    uint64_t *e = NULL;
    uint64_t k = scan_triggers(text, n, w, p, flags, &e);
    if (k == (uint64_t)-1) return 0;
    for (uint64_t i = 0; i < k && i < cap; i++) out[i] = e[i];
    free(e);
    return k;
}

/* ---- dictionary ------------------------------------------------------------------- */
typedef struct {
    uint64_t key;    /* reference phrase hash (+ probe displacement) */
    uint64_t start;  /* offset of the first occurrence in the bordered text */
    uint32_t len;
    uint32_t occ;
    uint32_t rank;   /* 1-based, assigned after the sort */
    uint32_t used;
} entry_t;

/* entries live in a growable array (stable indices); the probe table holds index+1 */
typedef struct {
    entry_t *ent;
    uint64_t count, cap;
    uint64_t *slot;
    uint64_t mask;
} dict_t;

static uint64_t slot_of(uint64_t key, uint64_t mask) {
    return ((key * 0x9E3779B97F4A7C15ULL) >> 20) & mask;
}

static int dict_grow(dict_t *d) {
    uint64_t ncap = d->slot ? (d->mask + 1) * 2 : 1024;
    uint64_t *ns = (uint64_t *)calloc(ncap, sizeof(uint64_t));
    if (!ns) return -1;
    for (uint64_t i = 0; i < d->count; i++) {
        uint64_t s = slot_of(d->ent[i].key, ncap - 1);
        while (ns[s]) s = (s + 1) & (ncap - 1);
        ns[s] = i + 1;
    }
    free(d->slot);
    d->slot = ns;
    d->mask = ncap - 1;
    return 0;
}

/* find-or-insert, returns the entry index or UINT64_MAX; same string <=> same entry (byte
 * compare on every hit, as newscan.cpp:258,282 do) */
static uint64_t dict_touch(dict_t *d, const uint8_t *ext, uint64_t start, uint32_t len) {
    if (!d->slot || d->count * 10 >= (d->mask + 1) * 6)
        if (dict_grow(d)) return UINT64_MAX;
    uint64_t key = pfp_oracle_kr_hash(ext + start, len);
    uint64_t s = slot_of(key, d->mask);
    for (;;) {
        if (!d->slot[s]) {
            if (d->count == d->cap) {
                uint64_t nc = d->cap ? d->cap * 2 : 1024;
                entry_t *ne = (entry_t *)realloc(d->ent, nc * sizeof(entry_t));
                if (!ne) return UINT64_MAX;
                d->ent = ne; d->cap = nc;
            }
            entry_t *e = &d->ent[d->count];
            e->used = 1; e->key = key; e->start = start; e->len = len; e->occ = 0; e->rank = 0;
            d->slot[s] = ++d->count;
            return d->count - 1;
        }
        entry_t *e = &d->ent[d->slot[s] - 1];
        if (e->key == key && e->len == len && memcmp(ext + e->start, ext + start, len) == 0)
            return d->slot[s] - 1;
        s = (s + 1) & d->mask;
    }
}

/* std::string ordering (newscan.cpp:387-390): unsigned bytes, shorter first on a tie */
static const uint8_t *g_sort_base;
static int entry_cmp(const void *a, const void *b) {
    const entry_t *x = *(entry_t *const *)a, *y = *(entry_t *const *)b;
    uint32_t m = x->len < y->len ? x->len : y->len;
    int c = memcmp(g_sort_base + x->start, g_sort_base + y->start, m);
    if (c) return c;
    return (x->len > y->len) - (x->len < y->len);
}

static void put_le(uint8_t *dst, uint64_t v, int nbytes) {
    for (int i = 0; i < nbytes; i++) dst[i] = (uint8_t)(v >> (8 * i));
}

int pfp_oracle_parse(const uint8_t *text, uint64_t n, uint32_t w, uint32_t p,
                     uint32_t flags, pfp_oracle_result *out) {
    memset(out, 0, sizeof(*out));
    if (w < 1 || p < 1) return -1;
    double t0 = now_sec();
    n = valid_prefix(text, n);
    out->n_text = n;

    /* bordered text: one 0x02 in front (newscan.cpp:329), w of them behind (:376) */
    uint8_t *ext = (uint8_t *)malloc(n + w + 1);
    if (!ext) return -2;
    ext[0] = DOLLAR;
    memcpy(ext + 1, text, n);
    memset(ext + 1 + n, DOLLAR, w);

    uint64_t *e = NULL;
    uint64_t k = scan_triggers(text, n, w, p, flags, &e);
    if (k == (uint64_t)-1) { free(ext); return -2; }

    uint64_t np = k + 1;
    out->n_phrases = np;
    uint64_t *which = (uint64_t *)malloc(np * sizeof(uint64_t));
    out->last = (uint8_t *)malloc(np);
    out->sai = (uint8_t *)malloc(np * IBYTES);
    out->parse = (uint32_t *)malloc(np * sizeof(uint32_t));
    dict_t d = {0};
    if (!which || !out->last || !out->sai || !out->parse) goto oom;

    /* phrase cut, newscan.cpp:245-304.  In ext coordinates text position i is i+1. */
    uint64_t pos = 0;  /* end position + 1 of the previous word, as in the reference */
    for (uint64_t j = 0; j < np; j++) {
        uint64_t start = (j == 0) ? 0 : e[j - 1] + 1 - w + 1;           /* keeps w overlap (:303) */
        uint64_t endx = (j < k) ? e[j] + 1 : n + w;                     /* inclusive, ext coords  */
        uint64_t len = endx - start + 1;
        if (len > 0xFFFFFFFFu) goto oom;
        uint64_t ei = dict_touch(&d, ext, start, (uint32_t)len);
        if (ei == UINT64_MAX) goto oom;
        if (d.ent[ei].occ == 0xFFFFFFFFu) goto oom;                     /* :277-281 */
        d.ent[ei].occ++;
        which[j] = ei;
        out->last[j] = ext[start + len - w - 1];                        /* :296 */
        if (pos == 0) pos = len - 1; else pos += len - w;               /* :299-300 */
        put_le(out->sai + j * IBYTES, pos, IBYTES);                     /* :301 */
    }
    out->sec_scan = now_sec() - t0;

    /* sort + write, newscan.cpp:622-639 and writeDictOcc :394-441 */
    t0 = now_sec();
    uint64_t nd = d.count;
    out->n_distinct = nd;
    entry_t **order = (entry_t **)malloc((nd ? nd : 1) * sizeof(entry_t *));
    if (!order) goto oom;
    uint64_t q = 0, sum = 0;
    for (uint64_t i = 0; i < nd; i++) { order[q++] = &d.ent[i]; sum += d.ent[i].len; }
    g_sort_base = ext;
    qsort(order, nd, sizeof(entry_t *), entry_cmp);
    out->sum_word_len = sum;
    out->dict_len = sum + nd + 1;
    out->dict = (uint8_t *)malloc(out->dict_len);
    out->occ = (uint32_t *)malloc((nd ? nd : 1) * sizeof(uint32_t));
    if (!out->dict || !out->occ) { free(order); goto oom; }
    uint64_t off = 0;
    for (uint64_t r = 0; r < nd; r++) {
        memcpy(out->dict + off, ext + order[r]->start, order[r]->len);
        off += order[r]->len;
        out->dict[off++] = END_OF_WORD;                                 /* :416 */
        out->occ[r] = order[r]->occ;                                    /* :433 */
        order[r]->rank = (uint32_t)(r + 1);                             /* :436 */
    }
    out->dict[off++] = END_OF_DICT;                                     /* :438 */
    free(order);
    out->sec_sort = now_sec() - t0;

    /* remapParse, newscan.cpp:443-466 */
    t0 = now_sec();
    for (uint64_t j = 0; j < np; j++) out->parse[j] = d.ent[which[j]].rank;
    out->sec_remap = now_sec() - t0;

    free(which); free(d.ent); free(d.slot); free(e); free(ext);
    return 0;
oom:
    free(which); free(d.ent); free(d.slot); free(e); free(ext);
    pfp_oracle_free(out);
    return -2;
}

void pfp_oracle_free(pfp_oracle_result *r) {
    free(r->dict); free(r->occ); free(r->parse); free(r->last); free(r->sai);
    r->dict = NULL; r->occ = NULL; r->parse = NULL; r->last = NULL; r->sai = NULL;
}

/* ---- FASTA/FASTQ text definition, kseq.h:177-218 + newscan.cpp:338-349 ------------- */
/* A small cursor over the file bytes standing in for kstream_t. */
typedef struct { const uint8_t *b; uint64_t n, i; } cur_t;
static int cur_getc(cur_t *c) { return c->i < c->n ? (int)c->b[c->i++] : -1; }

/* append the rest of the current line to dst[*l..], consume the '\n', drop one trailing
 * '\r' (kseq.h:141: only when the accumulated length exceeds 1).  Returns -1 if nothing at
 * all was available (EOF before any byte), else 0. */
static int cur_line_append(cur_t *c, uint8_t *dst, uint64_t *l) {
    if (c->i >= c->n) return -1;
    while (c->i < c->n && c->b[c->i] != '\n') dst[(*l)++] = c->b[c->i++];
    if (c->i < c->n) c->i++;
    if (*l > 1 && dst[*l - 1] == '\r') (*l)--;
    return 0;
}

uint64_t pfp_oracle_fasta_extract(const uint8_t *file, uint64_t n, uint8_t *out,
                                  int *truncated) {
    cur_t c = { file, n, 0 };
    uint64_t total = 0;
    int last_char = 0, ch;
    uint8_t *seq = (uint8_t *)malloc(n + 2);
    uint8_t *qual = (uint8_t *)malloc(n + 2);
    if (truncated) *truncated = 0;
    if (!seq || !qual) { free(seq); free(qual); return 0; }
    for (;;) {
        /* kseq.h:181-185: find the next header unless its first char was already read */
        if (last_char == 0) {
            while ((ch = cur_getc(&c)) >= 0 && ch != '>' && ch != '@') {}
            if (ch < 0) break;
            last_char = ch;
        }
        /* kseq.h:187-188: name up to whitespace, then the comment up to end of line */
        {
            int any = 0, delim = 0;
            if (c.i >= c.n) break;                       /* ks_getuntil returns -1 at EOF */
            while (c.i < c.n) {
                int x = c.b[c.i++];
                any = 1;
                if (x == ' ' || (x >= '\t' && x <= '\r')) { delim = x; break; }   /* isspace */
            }
            (void)any;
            if (delim != '\n' && delim != 0)
                while (c.i < c.n && c.b[c.i++] != '\n') {}
        }
        /* kseq.h:193-197: sequence lines up to a line starting with > + @ */
        uint64_t sl = 0;
        while ((ch = cur_getc(&c)) >= 0 && ch != '>' && ch != '+' && ch != '@') {
            if (ch == '\n') continue;
            seq[sl++] = (uint8_t)ch;
            cur_line_append(&c, seq, &sl);
        }
        if (ch == '>' || ch == '@') last_char = ch;
        int fastq_ok = 1;
        if (ch == '+') {
            /* kseq.h:207-215 */
            uint64_t ql = 0;
            while ((ch = cur_getc(&c)) >= 0 && ch != '\n') {}
            if (ch < 0) fastq_ok = 0;                    /* -2: no quality string */
            else {
                while (cur_line_append(&c, qual, &ql) >= 0 && ql < sl) {}
                last_char = 0;
                if (ql != sl) fastq_ok = 0;              /* -2: length mismatch    */
            }
        }
        if (!fastq_ok) break;                            /* kseq_read < 0 ends the loop (:338) */
        /* newscan.cpp:339-347: the (signed) char goes through toupper, stop at c <= Dollar.
         * glibc's toupper maps -128..-2 to the byte's unsigned value and leaves EOF (-1, byte
         * 0xFF) alone, so bytes 0x80..0xFE are ordinary symbols and 0xFF ends the input
         * (measured against newscanNT.x: golden cases fasta_high_byte, fasta_ff_byte). */
        for (uint64_t i = 0; i < sl; i++) {
            int v = seq[i];
            if (v >= 'a' && v <= 'z') v -= 32;
            if (v <= (int)DOLLAR || v == 0xFF) { if (truncated) *truncated = 1; goto done; }
            out[total++] = (uint8_t)v;
        }
        if (ch < 0) break;   /* EOF: the next kseq_read would return -1 */
    }
done:
    free(seq); free(qual);
    return total;
}
