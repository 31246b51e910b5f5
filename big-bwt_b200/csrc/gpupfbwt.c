/* gpupfbwt.c -- pfbwt.x-compatible command line over libpfpb200 (reference main(): pfbwt.cpp:318-407,
 * options :257-315).  Install it as pfbwt.x / pfbwtNT.x (/ pfbwt64.x / pfbwtNT64.x) next to the
 * unchanged `bigbwt` script, which runs `pfbwt*.x -w W [-s] [-e] [-S] [-t T] <file>` and only looks
 * at the exit status (bigbwt:130-150,231-240). */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "../../include/pfpb200.h"

static void print_help(const char *name, int w) {
    printf("Usage: %s <input filename> [options]\n", name);
    printf("  Options: \n");
    printf("\t-w W\tsliding window size, def. %d\n", w);
    printf("\t-t M\taccepted for compatibility (helper threads of the reference)\n");
    printf("\t-g G\tCUDA device index, def. 0\n");
    printf("\t-h  \tshow help and exit\n");
    printf("\t-s  \tcompute sampled suffix array\n");
    printf("\t-e  \tcompute sampled suffix array at the ends of the runs\n");
    printf("\t-S  \tcompute full suffix array\n");
    exit(1);
}

int main(int argc, char **argv) {
    int c, w = 10, device = 0, th = 0;
    unsigned flags = 0;
    time_t start = time(NULL);
    puts("==== Command line:");
    for (int i = 0; i < argc; i++) printf(" %s", argv[i]);
    puts("");
    while ((c = getopt(argc, argv, "t:w:sehSg:")) != -1) {
        switch (c) {
            case 's': flags |= PFPB200_PFBWT_SSA; break;
            case 'e': flags |= PFPB200_PFBWT_ESA; break;
            case 'S': flags |= PFPB200_PFBWT_SA; break;
            case 'w': w = atoi(optarg); break;
            case 't': th = atoi(optarg); break;
            case 'g': device = atoi(optarg); break;
            case 'h': print_help(argv[0], w); break;
            default: puts("Unknown option. Use -h for help."); exit(1);
        }
    }
    if (argc != optind + 1) { puts("Invalid number of arguments"); print_help(argv[0], w); }
    if ((flags & PFPB200_PFBWT_SA) && (flags & (PFPB200_PFBWT_SSA | PFPB200_PFBWT_ESA))) {
        printf("You can either require the sampled SA or the full SA, not both");
        exit(1);
    }
    if (w < 4) { puts("Windows size must be at least 4"); exit(1); }
    if (th < 0) { puts("Number of threads cannot be negative"); exit(1); }
    pfpb200_ctx *ctx = NULL;
    int rc = pfpb200_create(device, &ctx);
    if (rc != PFPB200_OK) {
        fprintf(stderr, "gpupfbwt: cannot use CUDA device %d: %s\n", device, pfpb200_strerror(rc));
        return 1;
    }
    pfpb200_pfbwt_result r;
    rc = pfpb200_pfbwt_file(ctx, argv[optind], (uint32_t)w, flags, &r);
    if (rc != PFPB200_OK) {
        fprintf(stderr, "gpupfbwt: %s: %s\n", pfpb200_strerror(rc), pfpb200_last_error(ctx));
        pfpb200_destroy(ctx);
        return 1;
    }
    printf("Dictionary file size: %llu\n", (unsigned long long)r.dict_bytes);
    printf("Dictionary words: %llu\n", (unsigned long long)r.dict_words);
    printf("Parsing size: %llu\n", (unsigned long long)r.parse_size);
    printf("bwlast file size: %llu\n", (unsigned long long)r.parse_size);
    printf("Full words: %llu\n", (unsigned long long)r.dict_words);
    printf("Easy bwt chars: %llu\n", (unsigned long long)r.easy);
    printf("Hard bwt chars: %llu\n", (unsigned long long)r.hard);
    printf("GPU pfbwt: %.3f ms (dictionary suffix sort %.3f in %u doubling rounds, BWT%s %.3f), %u kernel launches\n",
           r.ms_total, r.ms_sa, r.rounds, flags ? " + SA" : "", r.ms_fill, r.launches);
    printf("==== Elapsed time: %.0f wall clock seconds\n", difftime(time(NULL), start));
    pfpb200_destroy(ctx);
    return 0;
}
