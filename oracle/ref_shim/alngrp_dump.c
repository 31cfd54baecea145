/* alngrp_dump.c — TEST INFRASTRUCTURE, not reference code: a driver around the UNMODIFIED
 * reference functions saiset_create / alngrp_create (saiset.c:15-78), linked against the reference
 * objects built by Makefile.ref.  Used to pin oracle/alnoracle.c:orc_alngrp_merge (scope row N4).
 *
 *   alngrp_dump <n_reads> <a.sai> <b.sai> [...]   (single-end: the same file is given for both ends)
 * stdout, per read: u32 n, then n x { u32 dbidx, 16-byte bwt_aln1_t }.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "saiset.h"

void bwa_print_sam_PG(void) {} /* lives in the reference's main.cpp, which this driver replaces; never called here */

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: alngrp_dump <n_reads> <x.sai>...\n"); return 1; }
    int n_reads = atoi(argv[1]), n = argc - 2, i, r;
    const char ***files = calloc(n, sizeof(char **));
    dbset_t dbs;
    dbs.count = n;
    dbs.db = calloc(n, sizeof(bwtdb_t *));
    for (i = 0; i < n; ++i) {
        files[i] = calloc(2, sizeof(char *));
        files[i][0] = files[i][1] = argv[2 + i];
    }
    saiset_t *s = saiset_create(n, files);
    for (r = 0; r < n_reads; ++r) {
        alngrp_t *ag = alngrp_create(&dbs, s, 0);
        uint32_t cnt = (uint32_t)ag->n;
        fwrite(&cnt, 4, 1, stdout);
        for (i = 0; i < (int)cnt; ++i) {
            fwrite(&ag->a[i].dbidx, 4, 1, stdout);
            fwrite(&ag->a[i].aln, sizeof(bwt_aln1_t), 1, stdout);
        }
        alngrp_destroy(ag);
    }
    return 0;
}
