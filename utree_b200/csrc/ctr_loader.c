/* ctr_loader.c -- host-side CTR parser (replaces XT_read32, itree.c:733-828,
 * and the label-tail reader readSamplesFPdelim / READ_ADD_SAMPLES /
 * addSampleUdX, itree.c:1154-1223, 191-220).
 *
 * The file is memory-mapped read-only; prefix index and record blob stay as
 * views into the mapping (they are streamed to HBM by utb_db_upload), only
 * the label tail is parsed into a compact table:
 *   - ids number the DISTINCT label strings in order of first appearance
 *     (the reference's BST insert returns the old id for a repeated string);
 *   - rank[ix] is the label's position in strcmp (unsigned byte) order, which
 *     turns the reference's qsort-by-string (itree.c:1041) into an integer
 *     sort on the device.
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include "utb_internal.h"
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static __thread char g_err[512];
void utb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *utb_last_error(void) { return g_err; }

static uint64_t fnv1a(const char *s, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= (uint8_t)s[i]; h *= 1099511628211ull; }
    return h;
}

static int cmp_label(const void *a, const void *b, void *arg) {
    const utb_ctr *c = (const utb_ctr *)arg;
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return strcmp(c->blob + c->off[x], c->blob + c->off[y]);
}

/* Label tail -> blob/off/rank.  Each line is "label\tcount\n"; the label is
 * everything before the first tab (itree.c:1161-1162). */
static int parse_tail(utb_ctr *c, const char *tail, size_t n) {
    size_t lines = 0;
    for (const char *p = tail, *e = tail + n; p < e;) {
        const char *nl = memchr(p, '\n', (size_t)(e - p));
        ++lines;
        if (!nl) break;
        p = nl + 1;
    }
    c->blob = (char *)calloc(n + 128, 1);
    c->off = (uint32_t *)malloc((lines + 2) * sizeof(uint32_t));
    size_t cap = 16;
    while (cap < 2 * lines + 2) cap <<= 1;
    uint32_t *tab = (uint32_t *)malloc(cap * sizeof(uint32_t));
    if (!c->blob || !c->off || !tab) { free(tab); utb_set_error("out of memory (labels)"); return UTB_ERR_NOMEM; }
    memset(tab, 0xFF, cap * sizeof(uint32_t));
    uint32_t count = 0;
    size_t w = 0;
    for (const char *p = tail, *e = tail + n; p < e;) {
        const char *nl = memchr(p, '\n', (size_t)(e - p));
        size_t linelen = nl ? (size_t)(nl - p) : (size_t)(e - p);
        const char *tb = memchr(p, '\t', linelen);
        if (!tb) {  /* the reference scans past the line here: out of contract */
            free(tab);
            utb_set_error("label line %u has no tab", count);
            return UTB_ERR_FORMAT;
        }
        size_t ll = (size_t)(tb - p);
        size_t ll0 = strnlen(p, ll); /* an embedded NUL ends the C string */
        uint64_t h = fnv1a(p, ll0) & (cap - 1);
        int dup = 0;
        while (tab[h] != UTB_BAD32) {
            const char *q = c->blob + c->off[tab[h]];
            if (strlen(q) == ll0 && !memcmp(q, p, ll0)) { dup = 1; break; }
            h = (h + 1) & (cap - 1);
        }
        if (!dup) {
            tab[h] = count;
            c->off[count++] = (uint32_t)w;
            memcpy(c->blob + w, p, ll0);
            w += ll0 + 1;
        }
        if (!nl) break;
        p = nl + 1;
    }
    free(tab);
    c->off[count] = (uint32_t)w;
    c->blob_len = w + 64; /* zero padding so warp-wide compares may over-read */
    c->max_ix = count;
    c->rank = (uint32_t *)malloc((size_t)(count + 1) * sizeof(uint32_t));
    c->by_rank = (uint32_t *)malloc((size_t)(count + 1) * sizeof(uint32_t));
    if (!c->rank || !c->by_rank) { utb_set_error("out of memory (ranks)"); return UTB_ERR_NOMEM; }
    for (uint32_t i = 0; i < count; ++i) c->by_rank[i] = i;
    qsort_r(c->by_rank, count, sizeof(uint32_t), cmp_label, c);
    for (uint32_t i = 0; i < count; ++i) c->rank[c->by_rank[i]] = i;
    return UTB_OK;
}

int utb_ctr_open(const char *path, utb_ctr **out) {
    if (!path || !out) { utb_set_error("utb_ctr_open: null argument"); return UTB_ERR_ARG; }
    *out = NULL;
    int fd = open(path, O_RDONLY);
    if (fd < 0) { utb_set_error("Invalid DB file"); return UTB_ERR_IO; }             /* itree.c:735 */
    struct stat st;
    if (fstat(fd, &st) || st.st_size < 32) { close(fd); utb_set_error("Tree malformatted."); return UTB_ERR_FORMAT; }
    void *map = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) { close(fd); utb_set_error("mmap failed: %s", strerror(errno)); return UTB_ERR_IO; }
    const uint8_t *base = (const uint8_t *)map;
    uint64_t md[4];
    memcpy(md, base, 32);
    if (!md[3]) { munmap(map, (size_t)st.st_size); close(fd); utb_set_error("Tree malformatted."); return UTB_ERR_FORMAT; } /* :738 */
    /* itree.c:746-751: the reference binary is compiled for one WTYPE/IXTYPE;
     * one runtime build serves both label widths here (PACKSIZE=32, NO_COUNT). */
    if (md[0] != 8 || md[1] != 0 || (md[2] != 2 && md[2] != 4)) {
        munmap(map, (size_t)st.st_size); close(fd);
        utb_set_error("ERROR. Input tree requires PACKSIZE=%u, CNTTYPE size %u, IXTYPE size %u",
                      (unsigned)(md[0] << 2), (unsigned)md[1], (unsigned)md[2]);
        return UTB_ERR_FORMAT;
    }
    utb_ctr *c = (utb_ctr *)calloc(1, sizeof(*c));
    if (!c) { munmap(map, (size_t)st.st_size); close(fd); utb_set_error("out of memory"); return UTB_ERR_NOMEM; }
    c->fd = fd; c->map = map; c->map_len = (size_t)st.st_size;
    c->num_nodes = md[3];
    c->ix_bytes = (uint32_t)md[2];
    c->sz = 5 + c->ix_bytes;                                               /* itree.c:691 */
    c->binix_bytes = c->num_nodes < 0xFFFFFFFFull ? 4 : 8;                 /* itree.c:757 */
    uint64_t need = 32 + (uint64_t)UTB_NUMBINS * c->binix_bytes + c->num_nodes * c->sz;
    if ((uint64_t)st.st_size < need) {                                     /* itree.c:767-768 */
        utb_ctr_close(c); utb_set_error("Error in reading tree."); return UTB_ERR_FORMAT;
    }
    c->binix_raw = base + 32;
    c->recs = c->binix_raw + (uint64_t)UTB_NUMBINS * c->binix_bytes;
    if (c->binix_bytes == 4) { uint32_t v; memcpy(&v, c->binix_raw + (uint64_t)(UTB_NUMBINS - 1) * 4, 4); c->last_bin = v; }
    else memcpy(&c->last_bin, c->binix_raw + (uint64_t)(UTB_NUMBINS - 1) * 8, 8);
    int rc = parse_tail(c, (const char *)base + need, (size_t)((uint64_t)st.st_size - need));
    if (rc) { utb_ctr_close(c); return rc; }
    *out = c;
    return UTB_OK;
}

/* What XT_read32 prints before it gives up on a truncated tree (itree.c:754-768): the header fields and the
 * number of prefix-index entries its fread loop got.  Returns 0 when the header itself is readable and valid. */
int utb_ctr_probe(const char *path, uint64_t md[4], uint64_t *binix_read) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    int ok = fread(md, 8, 4, f) == 4 && md[3] && md[0] == 8 && md[1] == 0 && (md[2] == 2 || md[2] == 4);
    if (ok) {
        struct stat st;
        const uint64_t e = md[3] < 0xFFFFFFFFull ? 4 : 8;
        uint64_t got = !fstat(fileno(f), &st) && st.st_size > 32 ? ((uint64_t)st.st_size - 32) / e : 0;
        *binix_read = got < UTB_NUMBINS ? got : UTB_NUMBINS;
    }
    fclose(f);
    return ok ? 0 : -1;
}

void utb_ctr_close(utb_ctr *c) {
    if (!c) return;
    if (c->map) munmap(c->map, c->map_len);
    if (c->fd >= 0) close(c->fd);
    free(c->blob); free(c->off); free(c->rank); free(c->by_rank);
    free(c);
}

uint64_t utb_ctr_num_nodes(const utb_ctr *c) { return c->num_nodes; }
uint32_t utb_ctr_ix_bytes(const utb_ctr *c) { return c->ix_bytes; }
uint32_t utb_ctr_binix_bytes(const utb_ctr *c) { return c->binix_bytes; }
uint32_t utb_ctr_max_ix(const utb_ctr *c) { return c->max_ix; }
uint64_t utb_ctr_last_bin(const utb_ctr *c) { return c->last_bin; }
const char *utb_ctr_label(const utb_ctr *c, uint32_t ix) { return ix < c->max_ix ? c->blob + c->off[ix] : NULL; }
uint32_t utb_ctr_label_rank(const utb_ctr *c, uint32_t ix) { return ix < c->max_ix ? c->rank[ix] : UTB_BAD32; }
