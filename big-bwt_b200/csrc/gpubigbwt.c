/* gpubigbwt.c -- the `bigbwt` driver (bigbwt:35-150) as ONE process on the GPU: the input file streams
 * into HBM, the parse, the BWT of the parse and the final BWT are computed there
 * (pfpb200_bigbwt_file), and only <file>.bwt (and .sa / .ssa / .esa) are written -- the reference
 * runs newscan, bwtparse and pfbwt as three processes that hand each other files.  Options as
 * bigbwt's: -w -p -f -s -e -S -k; ours: -g (CUDA device). */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "../../include/pfpb200.h"

static void usage(const char *exe) {
    printf("Usage: %s <input filename> [options]\n", exe);
    printf("  Options: \n");
    printf("\t-w W\tsliding window size, def. 10\n");
    printf("\t-p M\thash modulus, def. 100\n");
    printf("\t-f  \tread fasta\n");
    printf("\t-s  \tcompute the start of run-length encoded BWT intervals of the sampled SA (.ssa)\n");
    printf("\t-e  \tcompute the end of run-length encoded BWT intervals of the sampled SA (.esa)\n");
    printf("\t-S  \tcompute the full suffix array (.sa)\n");
    printf("\t-k  \tkeep the intermediate files of the three stages\n");
    printf("\t-t T\taccepted for compatibility (helper threads of the reference's stages)\n");
    printf("\t-g G\tCUDA device index, def. 0\n");
    exit(1);
}

int main(int argc, char **argv) {
    pfpb200_opts o = {10, 100, 0, 0};
    long w = 10, p = 100;
    int c, device = 0, keep = 0;
    unsigned flags = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    puts("==== Command line:");
    for (int i = 0; i < argc; i++) printf(" %s", argv[i]);
    puts("");
    while ((c = getopt(argc, argv, "w:p:fseSkhg:t:v")) != -1) {
        switch (c) {
            case 'w': w = strtol(optarg, NULL, 10); break;
            case 'p': p = strtol(optarg, NULL, 10); break;
            case 'f': o.flags |= PFPB200_F_FASTA; break;
            case 's': flags |= PFPB200_PFBWT_SSA; break;
            case 'e': flags |= PFPB200_PFBWT_ESA; break;
            case 'S': flags |= PFPB200_PFBWT_SA; break;
            case 'k': keep = 1; break;
            case 't': break;                       /* bigbwt -t T: helper threads of the CPU stages, nothing to do here */
            case 'v': break;
            case 'g': device = atoi(optarg); break;
            case 'h': usage(argv[0]); break;
            default: puts("Unknown option. Use -h for help."); exit(1);
        }
    }
    if (argc != optind + 1) { puts("Invalid number of arguments"); usage(argv[0]); }
    if (w < 4) { puts("Windows size must be at least 4"); exit(1); }
    if (p < 10) { puts("Modulus must be at least 10"); exit(1); }
    if ((flags & PFPB200_PFBWT_SA) && (flags & (PFPB200_PFBWT_SSA | PFPB200_PFBWT_ESA))) {
        puts("You can either compute the full SA or a sample of it, not both. Exiting...");      /* bigbwt:59-61 */
        exit(1);
    }
    if (w > 65536 || p > 0x7FFFFFFFL) { puts("Option value too large"); exit(1); }
    o.w = (uint32_t)w; o.p = (uint32_t)p;
    pfpb200_ctx *ctx = NULL;
    int rc = pfpb200_create(device, &ctx);
    if (rc != PFPB200_OK) {
        fprintf(stderr, "gpubigbwt: cannot use CUDA device %d: %s\n", device, pfpb200_strerror(rc));
        return 1;
    }
    pfpb200_stats st;
    pfpb200_bwtparse_result bp;
    pfpb200_pfbwt_result r;
    rc = pfpb200_bigbwt_file(ctx, argv[optind], &o, flags, keep, &st, &bp, &r);
    if (rc != PFPB200_OK) {
        fprintf(stderr, "gpubigbwt: %s: %s\n", pfpb200_strerror(rc), pfpb200_last_error(ctx));
        pfpb200_destroy(ctx);
        return 1;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    printf("Total input symbols: %llu\n", (unsigned long long)st.n_text);
    printf("Found %llu distinct words\n", (unsigned long long)st.n_distinct);
    printf("Total number of words: %llu\n", (unsigned long long)st.n_phrases);
    printf("Easy bwt chars: %llu\n", (unsigned long long)r.easy);
    printf("Hard bwt chars: %llu\n", (unsigned long long)r.hard);
    printf("GPU: parse %.3f ms, bwtparse %.3f ms (%u rounds), pfbwt %.3f ms (suffix sort %.3f in %u rounds, BWT%s %.3f)\n",
           st.ms_total, bp.ms_total, bp.rounds, r.ms_total, r.ms_sa, r.rounds, flags ? " + SA" : "", r.ms_fill);
    printf("File read: %.3f s; stages after the parse + file write: %.3f s\n", st.sec_read, st.sec_write);
    printf("==== Elapsed time: %.3f wall clock seconds\n",
           (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
    pfpb200_destroy(ctx);
    return 0;
}
