/* pfp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the prefix-free-parsing stage of alshai/Big-BWT
 * (reference: newscan.cpp / pscan.cpp).  Used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the checker
 * for the CUDA path.  It is never linked into, imported by or executed from
 * the product (big-bwt_b200/): the product fails loudly without its CUDA
 * library.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * byte-for-byte against outputs of the reference's own newscanNT.x (built
 * unmodified into oracle/_ref/ by oracle/Makefile), both live when
 * oracle/_ref exists and through the committed fixtures in tests/golden/.
 */
#ifndef PFP_ORACLE_H
#define PFP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flags */
#define PFP_ORACLE_THREADED_RULE 1u /* first trigger must end at e >= w (newscan.hpp:66,206,
                                       pscan.hpp:93) instead of e >= w-1 (newscan.cpp:247-248) */

typedef struct {
    uint64_t n_text;       /* symbols consumed (input is cut at the first byte <= 0x02) */
    uint64_t n_phrases;    /* words in the parse                                        */
    uint64_t n_distinct;   /* dictionary words                                          */
    uint64_t sum_word_len; /* "Sum of lenghts of dictionary words" (newscan.cpp:633)    */
    uint64_t dict_len;     /* bytes of .dict = sum_word_len + n_distinct + 1            */
    uint8_t  *dict;        /* .dict  (newscan.cpp:406-438)                              */
    uint32_t *occ;         /* .occ   (newscan.cpp:433)                                  */
    uint32_t *parse;       /* .parse (newscan.cpp:456-458), 1-based ranks               */
    uint8_t  *last;        /* .last  (newscan.cpp:296)                                  */
    uint8_t  *sai;         /* .sai   (newscan.cpp:299-301), 5 bytes LE per phrase       */
    double   sec_scan, sec_sort, sec_remap;
} pfp_oracle_result;

/* Parse text[0..n) with window w and modulus p. Returns 0, or <0 on bad args / OOM. */
int pfp_oracle_parse(const uint8_t *text, uint64_t n, uint32_t w, uint32_t p,
                     uint32_t flags, pfp_oracle_result *out);
void pfp_oracle_free(pfp_oracle_result *r);

/* Just the trigger end positions E (ascending), the first stage of the path.
 * Returns the count; writes at most cap of them. */
uint64_t pfp_oracle_triggers(const uint8_t *text, uint64_t n, uint32_t w, uint32_t p,
                             uint32_t flags, uint64_t *out, uint64_t cap);

/* kseq-equivalent FASTA/FASTQ sequence extraction (kseq.h:177-218) followed by the
 * per-byte toupper / validity rule of newscan.cpp:339-348.  out must hold n bytes.
 * Returns the number of text bytes; *truncated = 1 if an invalid byte stopped it. */
uint64_t pfp_oracle_fasta_extract(const uint8_t *file, uint64_t n, uint8_t *out,
                                  int *truncated);

/* The reference's own 64-bit phrase hash (newscan.cpp:229-239); exposed for KATs. */
uint64_t pfp_oracle_kr_hash(const uint8_t *s, uint64_t len);
/* The reference's window hash of the w bytes s[0..w) (newscan.cpp:194-202). */
uint64_t pfp_oracle_window_hash(const uint8_t *s, uint32_t w);

#ifdef __cplusplus
}
#endif
#endif
