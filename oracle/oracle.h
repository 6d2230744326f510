/* oracle.h -- CPU restatement of UTree's SEARCH_GG hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under utree_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status: the reference ships no tests or golden vectors
 * (SURVEY.md 4.1), so this oracle is pinned by EXECUTING the reference:
 * tests/test_oracle_golden.py diffs its output byte-for-byte against the
 * committed fixtures in tests/golden/ that were produced by
 * oracle/_ref/utree-search_gg and oracle/_ref/utree-search (built from
 * /root/reference/itree.c by oracle/Makefile; scripts/make_golden*.py), and
 * tests/test_gpu_scale.py runs those binaries live beside the CUDA path.
 *
 * Every function cites the reference lines it restates (file itree.c).
 */
#ifndef UTREE_ORACLE_H
#define UTREE_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcDB OrcDB;

/* Per-lookup / per-run accounting used for the roofline. */
typedef struct {
    uint64_t reads;        /* records parsed                               */
    uint64_t good_finds;   /* reads with >=1 hit (itree.c:1029)            */
    uint64_t lookups;      /* XT_getIX32 calls                             */
    uint64_t hits;         /* lookups with ix < maxIX                      */
    uint64_t probes;       /* suffix compares incl. the final equality     */
    uint64_t sect_idx;     /* distinct 32B sectors touched in BinIx        */
    uint64_t sect_bkt;     /* distinct 32B sectors touched in the records  */
    uint64_t out_bytes;    /* bytes of output text                         */
} OrcStats;

/* Result of one read's vote (itree.c:1028-1098). */
enum { ORC_NONE = 0, ORC_STAR = 1, ORC_WALK = 2 };
#define ORC_DV_EMPTY 0xFFFFFFFFu /* taxonomy "" (dv == -1)      */
#define ORC_DV_FULL  0xFFFFFFFEu /* whole label (dv == -2)      */
typedef struct {
    uint32_t kind;      /* ORC_NONE / ORC_STAR / ORC_WALK                  */
    uint32_t label;     /* label id whose string (prefix) is printed       */
    uint32_t cut;       /* WALK: dv (prefix length, or EMPTY / FULL)       */
    uint32_t found;     /* foundUniq                                       */
    uint32_t uix;       /* distinct labels                                 */
    uint32_t sl, ol;    /* WALK only                                       */
} OrcVote;

/* itree.c:733-828 XT_read32 + :1154-1223 label tail.  NULL on failure,
 * message in err. */
OrcDB *orc_db_load(const char *path, char *err, size_t errlen);
void orc_db_free(OrcDB *db);
uint64_t orc_db_num_nodes(const OrcDB *db);
uint32_t orc_db_max_ix(const OrcDB *db);     /* sampIX + 1 (itree.c:855)     */
uint32_t orc_db_ix_bytes(const OrcDB *db);   /* sizeof(IXTYPE) of the file   */
uint32_t orc_db_binix_bytes(const OrcDB *db);/* on-disk BinIx entry width    */
const char *orc_db_label(const OrcDB *db, uint32_t ix);
/* raw views (for tests that cross-check the product loader) */
const uint64_t *orc_db_binix(const OrcDB *db);
const uint8_t *orc_db_records(const OrcDB *db);

/* itree.c:720-730 XT_getIX32 + :699-707 xtSuffixBS.  Returns label id or
 * 0xFFFFFFFF (BAD_IX widened).  st may be NULL. */
uint32_t orc_lookup(const OrcDB *db, uint64_t word, OrcStats *st);

/* itree.c:887-898 (RC concat) + :906-933 (slide).  seq/len is the trimmed
 * sequence line.  Writes up to cap hit ids to hits (pass NULL to only count);
 * if words != NULL also records every looked-up word (cap_words entries) and
 * returns their number in *n_words.  Returns foundUniq. */
uint64_t orc_slide(const OrcDB *db, const char *seq, uint32_t len, int do_rc,
                   uint32_t *hits, uint64_t cap,
                   uint64_t *words, uint64_t cap_words, uint64_t *n_words,
                   OrcStats *st);

/* itree.c:1028-1098.  hits may be permuted freely. */
void orc_vote(const OrcDB *db, const uint32_t *hits, uint64_t n, OrcVote *out);

/* Formats one output line exactly as itree.c:1032/1040/1096.  Returns bytes
 * written (0 for ORC_NONE).  buf must hold name + label + 64. */
size_t orc_format(const OrcDB *db, const char *name, const OrcVote *v, char *buf);

/* Whole search, input order (== reference threads=1), itree.c:833-1108.
 * threads>1 parallelises over reads but still emits in input order.
 * max_reads==0 means all.  Returns 0 ok, else the reference's exit code
 * (1 bad fasta, 2 format error) with message in err. */
int orc_search_file(const OrcDB *db, const char *fasta, const char *out,
                    int do_rc, int threads, uint64_t max_reads,
                    OrcStats *st, char *err, size_t errlen);

/* The non-GG binary (-D SEARCH, itree.c:948-951, 969-1007): the slide skips
 * PACKSIZE/SPARSITY - 1 windows after a hit, the vote is a top-2 plurality
 * over the hits plus one stale entry left by an earlier read; sequential. */
#define ORC_SHALLOW_SKIP 7
uint64_t orc_slide_shallow(const OrcDB *db, const char *seq, uint32_t len, int do_rc,
                           uint32_t *hits, uint64_t cap, OrcStats *st);
int orc_search_file_shallow(const OrcDB *db, const char *fasta, const char *out, int do_rc,
                            OrcStats *st, char *err, size_t errlen);

/* helpers exported for unit tests */
uint64_t orc_revcomp_word(uint64_t w);                 /* rc of a 32-mer word */
int orc_pack_word(const char *bases32, uint64_t *w);   /* 0 if non-ACGT       */

#ifdef __cplusplus
}
#endif
#endif
