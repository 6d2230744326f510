/* oracle_search.c -- command-line front end of the CPU restatement
 * (TEST INFRASTRUCTURE; see oracle.h).  Same argv contract as the reference
 * search binary (itree.c:1357-1377) so outputs can be cmp'ed directly. */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: oracle_search compTree.ctr seqs.fa out.txt [threads] [SPEED <X>] [RC]\n"); return 1; }
    int do_rc = !strcmp(argv[argc - 1], "RC");
    argc -= do_rc;
    if (!strcmp(argv[argc - 2], "SPEED")) argc -= 2;
    int threads = argc >= 5 ? atoi(argv[4]) : 1;
    char err[256] = "";
    OrcDB *db = orc_db_load(argv[1], err, sizeof err);
    if (!db) { puts(err); return 0; }
    OrcStats st;
    int rc = orc_search_file(db, argv[2], argv[3], do_rc, threads, 0, &st, err, sizeof err);
    if (rc) { fprintf(stderr, "%s\n", err); return rc; }
    printf("Good finds: %llu\nSearched %llu queries\n", (unsigned long long)st.good_finds, (unsigned long long)st.reads);
    printf("lookups=%llu hits=%llu probes=%llu sect_idx=%llu sect_bkt=%llu bytes_per_lookup=%.2f\n",
           (unsigned long long)st.lookups, (unsigned long long)st.hits, (unsigned long long)st.probes,
           (unsigned long long)st.sect_idx, (unsigned long long)st.sect_bkt,
           st.lookups ? 32.0 * (double)(st.sect_idx + st.sect_bkt) / (double)st.lookups : 0.0);
    orc_db_free(db);
    return 0;
}
