/* utb_internal.h -- shared between the host C files and the CUDA TU. */
#ifndef UTB_INTERNAL_H
#define UTB_INTERNAL_H
#include "../../include/utree_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define UTB_NUMBINS ((1u << 24) + 1u)   /* itree.c:693 */
#define UTB_LINELEN 16777216u           /* itree.c:836 */
#define UTB_MAXSEQ (UTB_LINELEN - 2u)   /* longest sequence line the reference handles */
#define UTB_BAD32 0xFFFFFFFFu

/* Host view of a CTR file (memory-mapped; label table parsed). */
struct utb_ctr {
    int fd;
    void *map;
    size_t map_len;
    uint64_t num_nodes;
    uint32_t ix_bytes;      /* 2 / 4 */
    uint32_t binix_bytes;   /* 4 / 8 */
    uint32_t sz;            /* 5 + ix_bytes */
    uint32_t max_ix;        /* number of distinct labels */
    const uint8_t *binix_raw;   /* UTB_NUMBINS entries of binix_bytes */
    const uint8_t *recs;        /* num_nodes * sz bytes */
    uint64_t last_bin;
    /* label table: NUL-terminated strings back to back, padded */
    char *blob;
    size_t blob_len;
    uint32_t *off;          /* max_ix + 1 offsets into blob */
    uint32_t *rank;         /* rank[ix]: position of label ix in strcmp order */
    uint32_t *by_rank;      /* inverse permutation */
};

void utb_set_error(const char *fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
