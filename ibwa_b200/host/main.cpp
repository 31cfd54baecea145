/*
 * b200aln — command-line front end.  Replaces the `aln` sub-command of the
 * reference's dispatcher (main.cpp:35-59, `ibwa aln ...`): same options, same
 * index files in, same .sai bytes out.  Every other sub-command of the
 * reference (index, samse, sampe, bwasw ...) is out of scope and stays with
 * the reference binary.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/b200aln.h"

int main(int argc, char *argv[])
{
    if (argc < 2 || strcmp(argv[1], "aln") != 0) {
        fprintf(stderr, "\nProgram: b200aln (%s)\n", b200aln_version());
        fprintf(stderr, "Usage:   b200aln aln [options] <prefix> <in.fq>\n\n");
        fprintf(stderr, "Only the `aln` stage is provided; use the reference binary for index/samse/sampe.\n\n");
        return 1;
    }
    /* The process ends with the run: the contexts' device memory (tens of GB) is left to the driver's teardown
     * instead of being freed buffer by buffer, and no destructor of the CUDA runtime has to run. */
    setenv("B200ALN_FAST_EXIT", "1", 0);
    const int rc = b200aln_aln_main(argc - 1, argv + 1);
    fflush(NULL);
    _exit(rc);
}
