/* pfp_io.c -- host side of the drop-in boundary, plain C: input readers (plain text and
 * FASTA/FASTQ with the semantics of the reference's kseq reader) and writers of the five
 * output files in the reference's formats (utils.h:14-26, utils.c:33-54). */
#define _GNU_SOURCE
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>
#include "../../include/pfpb200.h"

/* gzip-compressed input, as the reference's kseq reader takes it through gzopen()/gzread()
 * (newscan.cpp:332-336): inflate the whole file into memory */
static int read_gz(const char *path, uint8_t **buf, uint64_t *n, char *err, size_t errlen) {
    gzFile g = gzopen(path, "rb");
    if (!g) { snprintf(err, errlen, "%s: %s", path, strerror(errno)); return -1; }
    gzbuffer(g, 1u << 20);
    uint64_t cap = (uint64_t)1 << 24, got = 0;
    uint8_t *b = (uint8_t *)malloc(cap);
    if (!b) { snprintf(err, errlen, "%s: out of memory", path); gzclose(g); return -1; }
    for (;;) {
        if (got == cap) {
            uint8_t *nb = (uint8_t *)realloc(b, cap * 2);
            if (!nb) { snprintf(err, errlen, "%s: out of memory inflating", path); free(b); gzclose(g); return -1; }
            b = nb;
            cap *= 2;
        }
        uint64_t want = cap - got;
        if (want > ((uint64_t)1 << 30)) want = (uint64_t)1 << 30;
        int r = gzread(g, b + got, (unsigned)want);
        if (r < 0) {
            int ec = 0;
            snprintf(err, errlen, "%s: %s", path, gzerror(g, &ec));
            free(b);
            gzclose(g);
            return -1;
        }
        if (r == 0) break;
        got += (uint64_t)r;
    }
    gzclose(g);
    *buf = b;
    *n = got;
    return 0;
}

/* gz_ok: a file starting with the gzip magic is inflated (FASTA mode; the reference reads plain
 * text through ifstream, newscan.cpp:354-358, so a .gz given without -f is parsed as bytes) */
int pfp_io_read_file(const char *path, int gz_ok, uint8_t **buf, uint64_t *n, char *err, size_t errlen) {
    *buf = NULL;
    *n = 0;
    FILE *f = fopen(path, "rb");
    if (!f) { snprintf(err, errlen, "%s: %s", path, strerror(errno)); return -1; }
    if (gz_ok) {
        int c0 = fgetc(f), c1 = fgetc(f);
        if (c0 == 0x1f && c1 == 0x8b) { fclose(f); return read_gz(path, buf, n, err, errlen); }
        rewind(f);
    }
    struct stat st;
    if (fstat(fileno(f), &st) != 0) { snprintf(err, errlen, "%s: %s", path, strerror(errno)); fclose(f); return -1; }
    uint64_t size = (uint64_t)st.st_size;
    uint8_t *b = (uint8_t *)malloc(size ? size : 1);
    if (!b) { snprintf(err, errlen, "%s: out of memory reading %llu bytes", path, (unsigned long long)size); fclose(f); return -1; }
    uint64_t got = 0;
    while (got < size) {
        size_t r = fread(b + got, 1, (size_t)(size - got), f);
        if (r == 0) break;
        got += r;
    }
    fclose(f);
    if (got != size) { snprintf(err, errlen, "%s: short read", path); free(b); return -1; }
    *buf = b;
    *n = size;
    return 0;
}

/* ---- FASTA / FASTQ ---------------------------------------------------------------------------
 * The text of `-f` mode is what kseq_read() returns record after record (kseq.h:177-218),
 * upper-cased and cut at the first invalid symbol (newscan.cpp:338-349):
 *   - a record starts at the next '>' or '@' (anywhere, when no header char is pending);
 *   - its name runs to the first whitespace, the comment to the end of that line;
 *   - sequence lines follow until a line whose first char is '>', '@' or '+'; newlines are
 *     dropped and one trailing '\r' per line too, once the record holds more than one byte;
 *   - '+' opens FASTQ qualities, read until they are as long as the sequence; a record with
 *     missing or mismatched qualities ends the input without contributing;
 *   - symbols 0x00..0x02 and 0xFF end the input (the reference pipes the signed char through
 *     toupper(), which returns EOF unchanged); a-z are upper-cased. */
static int is_space(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

/* append the rest of the line at *pi to dst[*len..] and step over its '\n'.  At EOF nothing
 * happens (kseq's ks_getuntil2 returns -1 before its '\r' rule); otherwise one trailing '\r' is
 * dropped once the accumulated string is longer than one byte (kseq.h:141).
 * Returns -1 at EOF, else 0. */
static int take_line(const uint8_t *f, uint64_t n, uint64_t *pi, uint8_t *dst, uint64_t *len) {
    uint64_t i = *pi;
    if (i >= n) return -1;
    const uint8_t *nl = (const uint8_t *)memchr(f + i, '\n', (size_t)(n - i));
    uint64_t stop = nl ? (uint64_t)(nl - f) : n;
    memmove(dst + *len, f + i, (size_t)(stop - i));
    *len += stop - i;
    *pi = nl ? stop + 1 : n;
    if (*len > 1 && dst[*len - 1] == '\r') (*len)--;
    return 0;
}

uint64_t pfpb200_fasta_extract(const uint8_t *f, uint64_t n, uint8_t *out, int *truncated) {
    uint64_t i = 0, total = 0;
    int pending = 0; /* header char already consumed */
    if (truncated) *truncated = 0;
    for (;;) {
        if (!pending) {
            while (i < n && f[i] != '>' && f[i] != '@') i++;
            if (i >= n) break;
            i++;
            pending = 1;
        }
        if (i >= n) break; /* header char was the last byte: no record */
        /* name, then comment */
        int delim = 0;
        while (i < n) {
            int c = f[i++];
            if (is_space(c)) { delim = c; break; }
        }
        if (delim != 0 && delim != '\n') {
            const uint8_t *nl = (const uint8_t *)memchr(f + i, '\n', (size_t)(n - i));
            i = nl ? (uint64_t)(nl - f) + 1 : n;
        }
        /* sequence lines, written tentatively behind the committed text */
        uint8_t *seq = out + total;
        uint64_t sl = 0;
        int c = -1;
        while (i < n) {
            c = f[i++];
            if (c == '>' || c == '+' || c == '@') break;
            if (c == '\n') { c = -1; continue; }
            seq[sl++] = (uint8_t)c;
            take_line(f, n, &i, seq, &sl);
            c = -1;
        }
        if (c == '>' || c == '@') pending = 1; /* stays set otherwise, like kseq's last_char */
        if (c == '+') {
            const uint8_t *nl = (const uint8_t *)memchr(f + i, '\n', (size_t)(n - i));
            if (!nl) break;                      /* no quality string */
            i = (uint64_t)(nl - f) + 1;
            /* qualities go to the free space behind the tentative sequence: sequence and
             * quality bytes are copies of disjoint file bytes, so total + sl + ql <= n */
            uint8_t *qual = seq + sl;
            uint64_t ql = 0;
            while (take_line(f, n, &i, qual, &ql) == 0 && ql < sl) {}
            pending = 0;
            if (ql != sl) break;                 /* truncated / mismatched qualities */
        }
        /* toupper + validity */
        for (uint64_t k = 0; k < sl; k++) {
            unsigned v = seq[k];
            if (v - 'a' < 26u) v -= 32;
            if (v <= 2u || v == 0xFFu) { if (truncated) *truncated = 1; return total + k; }
            seq[k] = (uint8_t)v;
        }
        total += sl;
    }
    return total;
}

/* ---- writers --------------------------------------------------------------------------------- */
static int write_all(const char *base, const char *ext, int seg, const void *p, uint64_t bytes,
                     char *err, size_t errlen) {
    char *name = NULL;
    int e = seg < 0 ? asprintf(&name, "%s.%s", base, ext)            /* utils.c:33-41 */
                    : asprintf(&name, "%s.%d.%s", base, seg, ext);   /* utils.c:44-54 */
    if (e < 1) { snprintf(err, errlen, "asprintf failed"); return -1; }
    FILE *f = fopen(name, "wb");
    if (!f) { snprintf(err, errlen, "%s: %s", name, strerror(errno)); free(name); return -1; }
    uint64_t done = 0;
    while (done < bytes) {
        size_t r = fwrite((const uint8_t *)p + done, 1, (size_t)(bytes - done), f);
        if (r == 0) break;
        done += r;
    }
    int bad = (done != bytes) | (fclose(f) != 0);
    if (bad) snprintf(err, errlen, "%s: write error", name);
    free(name);
    return bad ? -1 : 0;
}

int pfp_io_write_outputs(const char *path, const pfpb200_opts *opts, const pfpb200_outputs *o,
                         char *err, size_t errlen) {
    const char *dict_ext = (opts->flags & PFPB200_F_COMPRESS) ? "dicz" : "dict";
    if (write_all(path, dict_ext, -1, o->dict, o->dict_bytes, err, errlen)) return -1;
    if (write_all(path, "occ", -1, o->occ, 4 * o->n_distinct, err, errlen)) return -1;
    if (write_all(path, "parse", -1, o->parse, 4 * o->n_phrases, err, errlen)) return -1;
    const int T = opts->nseg;
    if (T <= 0) {
        if (write_all(path, "last", -1, o->last, o->n_phrases, err, errlen)) return -1;
        if (o->sai && write_all(path, "sai", -1, o->sai, 5 * o->n_phrases, err, errlen)) return -1;
    } else {
        /* newscan.hpp:274-276: one .last/.sai segment per helper thread; bwtparse -t T reads them
         * back to back (bwtparse.c:179,195; utils.c:57-105), so any split is equivalent */
        uint64_t P = o->n_phrases, per = (P + (uint64_t)T - 1) / (uint64_t)T;
        for (int s = 0; s < T; s++) {
            uint64_t a = (uint64_t)s * per, b = a + per;
            if (a > P) a = P;
            if (b > P) b = P;
            if (write_all(path, "last", s, o->last + a, b - a, err, errlen)) return -1;
            if (o->sai && write_all(path, "sai", s, o->sai + 5 * a, 5 * (b - a), err, errlen)) return -1;
        }
    }
    return 0;
}

/* The text T the parser sees for `path` (host memory): file bytes, or -- with PFPB200_F_FASTA --
 * the kseq-equivalent extraction of a plain or gzip-compressed FASTA/FASTQ file.  No GPU needed. */
int pfpb200_read_input(const char *path, uint32_t flags, uint8_t **text, uint64_t *n_text, int *truncated) {
    char err[256];
    if (!path || !text || !n_text) return PFPB200_E_ARG;
    *text = NULL;
    *n_text = 0;
    if (truncated) *truncated = 0;
    uint8_t *file = NULL;
    uint64_t fn = 0;
    if (pfp_io_read_file(path, (flags & PFPB200_F_FASTA) != 0, &file, &fn, err, sizeof(err)) != 0) return PFPB200_E_IO;
    if (!(flags & PFPB200_F_FASTA)) { *text = file; *n_text = fn; return PFPB200_OK; }
    uint8_t *seq = (uint8_t *)malloc(fn ? fn : 1);
    if (!seq) { free(file); return PFPB200_E_NOMEM; }
    int tr = 0;
    *n_text = pfpb200_fasta_extract(file, fn, seq, &tr);
    if (truncated) *truncated = tr;
    free(file);
    *text = seq;
    return PFPB200_OK;
}

void pfpb200_free_host(void *p) { free(p); }
