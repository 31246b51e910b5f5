/* gpuunparse.c -- unparse-compatible command line over libpfpb200 (reference main():
 * unparse.c:76-137, options :33-64): restores the original file from <basename>.dicz and
 * <basename>.parse, the files `bigbwt -c` keeps. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "../../include/pfpb200.h"

static void print_help(const char *name) {
    printf("Usage: %s <basename> [options]\n\n", name);
    puts("Restore the original file given a prefix free parse (files .dicz and .parse)");
    puts("  Options:");
    puts("\t-o outfile   output file (def. <basename>.out)");
    puts("\t-g G         CUDA device index (def. 0)");
    puts("\t-h           show help and exit");
    exit(1);
}

int main(int argc, char **argv) {
    int c, device = 0;
    const char *outname = NULL;
    puts("==== Command line:");
    for (int i = 0; i < argc; i++) printf(" %s", argv[i]);
    puts("\n");
    while ((c = getopt(argc, argv, "ho:g:")) != -1) {
        switch (c) {
            case 'o': outname = optarg; break;
            case 'g': device = atoi(optarg); break;
            case 'h': print_help(argv[0]); break;
            default: puts("Unknown option. Use -h for help."); exit(1);
        }
    }
    if (argc != optind + 1) print_help(argv[0]);
    time_t start = time(NULL);
    pfpb200_ctx *ctx = NULL;
    int rc = pfpb200_create(device, &ctx);
    if (rc != PFPB200_OK) {
        fprintf(stderr, "gpuunparse: cannot use CUDA device %d: %s\n", device, pfpb200_strerror(rc));
        return 1;
    }
    uint64_t words = 0, n = 0;
    float ms = 0;
    rc = pfpb200_unparse_file(ctx, argv[optind], outname, &words, &n, &ms);
    if (rc != PFPB200_OK) {
        fprintf(stderr, "gpuunparse: %s: %s\n", pfpb200_strerror(rc), pfpb200_last_error(ctx));
        pfpb200_destroy(ctx);
        return 1;
    }
    fprintf(stderr, "Found %llu dictionary words\n", (unsigned long long)words);
    if (outname) fprintf(stderr, "Recovering file %s\n", outname);
    else fprintf(stderr, "Recovering file %s.out\n", argv[optind]);
    printf("GPU unparse: %llu bytes in %.3f ms\n", (unsigned long long)n, ms);
    printf("==== Elapsed time: %.0f wall clock seconds\n", difftime(time(NULL), start));
    pfpb200_destroy(ctx);
    return 0;
}
