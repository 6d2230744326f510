/* compress.c -- `utree-compress preTree.ubt compTree.ctr` (XT_cmp32, itree.c:1234-1315;
 * SURVEY 8f-2): turns the builder's .ubt (header, (u64 word, IXTYPE id) records ascending by
 * word, label tail) into the .ctr the search path loads (App. A): header, BinIx over the
 * 24-bit prefixes, records with the 3 prefix bytes dropped, label tail.
 *
 * The output is byte-identical to the reference compressor's, INCLUDING its first-bin quirk
 * (itree.c:1284-1288): index 0 doubles as "unset", so a first bin that holds
 * exactly one record is lost and its record is folded into the next non-empty bin.  The
 * search path is exact on such files (quirk_bin), so reproducing the quirk keeps every
 * downstream tool -- including the reference's own search -- byte-compatible.
 *
 * Pure host code: the transformation is a single streaming pass (the records are already
 * sorted), I/O bound; there is nothing for the device to do here.
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include "utb_internal.h"
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

/* Returns UTB_OK; on failure the message is in utb_last_error().  *n_records / *n_labels
 * are optional outputs. */
static __thread uint64_t g_count_sum;   /* sum of the label counts written by the last utb_compress_ubt on this thread */
int utb_compress_ubt(const char *ubt_path, const char *ctr_path, uint64_t *n_records, uint32_t *n_labels) {
    if (!ubt_path || !ctr_path) { utb_set_error("utb_compress_ubt: null argument"); return UTB_ERR_ARG; }
    int fd = open(ubt_path, O_RDONLY);
    if (fd < 0) { utb_set_error("Invalid input filename"); return UTB_ERR_IO; }              /* itree.c:1236 */
    struct stat st;
    if (fstat(fd, &st) || st.st_size < 32) { close(fd); utb_set_error("Tree malformatted."); return UTB_ERR_FORMAT; }
    const uint8_t *base = (const uint8_t *)mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (base == MAP_FAILED) { close(fd); utb_set_error("mmap failed: %s", strerror(errno)); return UTB_ERR_IO; }
    int rc = UTB_OK;
    FILE *fo = NULL;
    uint32_t *binix32 = NULL; uint64_t *binix = NULL; uint8_t *buf = NULL;
    uint64_t md[4];
    memcpy(md, base, 32);
    if (!md[3]) { utb_set_error("Tree malformatted."); rc = UTB_ERR_FORMAT; goto done; }      /* itree.c:1239 */
    if (md[0] != 8 || md[1] != 0 || (md[2] != 2 && md[2] != 4)) {                             /* itree.c:1247-1252 */
        utb_set_error("ERROR. Input tree requires PACKSIZE=%u, CNTTYPE size %u, IXTYPE size %u",
                      (unsigned)(md[0] << 2), (unsigned)md[1], (unsigned)md[2]);
        rc = UTB_ERR_FORMAT; goto done;
    }
    const uint64_t n = md[3];
    const uint32_t ixb = (uint32_t)md[2], dr_sz = 8 + ixb, sz = 5 + ixb;
    if ((uint64_t)st.st_size < 32 + n * dr_sz) { utb_set_error("Error in reading tree."); rc = UTB_ERR_FORMAT; goto done; }
    const uint8_t *recs = base + 32;
    const uint8_t *tail = recs + n * dr_sz;
    const size_t tail_len = (size_t)((uint64_t)st.st_size - 32 - n * dr_sz);

    /* BinIx exactly as itree.c:1281-1289 builds it */
    binix = (uint64_t *)calloc(UTB_NUMBINS, sizeof(uint64_t));
    if (!binix) { utb_set_error("out of memory"); rc = UTB_ERR_NOMEM; goto done; }
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t w; memcpy(&w, recs + i * dr_sz, 8);
        uint32_t v = (uint32_t)(w >> 40);
        if (!binix[v]) binix[v] = i;                               /* index 0 is indistinguishable from "unset" */
    }
    binix[UTB_NUMBINS - 1] = n;
    size_t u = 0;
    for (; !binix[u]; ++u);
    binix[u] = 0;
    for (size_t i = UTB_NUMBINS - 2; i > u; --i) if (!binix[i]) binix[i] = binix[i + 1];

    fo = fopen(ctr_path, "wb");
    if (!fo) { utb_set_error("Invalid output filename"); rc = UTB_ERR_IO; goto done; }        /* itree.c:1299 */
    setvbuf(fo, NULL, _IOFBF, (size_t)8 << 20);
    if (fwrite(md, 8, 4, fo) != 4) { rc = UTB_ERR_IO; goto werr; }
    if (n < 0xFFFFFFFFull) {                                       /* itree.c:1303-1304: 4-byte entries */
        binix32 = (uint32_t *)malloc((size_t)UTB_NUMBINS * 4);
        if (!binix32) { utb_set_error("out of memory"); rc = UTB_ERR_NOMEM; goto done; }
        for (size_t i = 0; i < UTB_NUMBINS; ++i) binix32[i] = (uint32_t)binix[i];
        if (fwrite(binix32, 4, UTB_NUMBINS, fo) != UTB_NUMBINS) { rc = UTB_ERR_IO; goto werr; }
    } else if (fwrite(binix, 8, UTB_NUMBINS, fo) != UTB_NUMBINS) { rc = UTB_ERR_IO; goto werr; }
    /* records: low 5 bytes of the word + the id (itree.c:1306-1309), in 8 MiB blocks */
    {
        const size_t BLK = ((size_t)8 << 20) / sz;
        buf = (uint8_t *)malloc(BLK * sz);
        if (!buf) { utb_set_error("out of memory"); rc = UTB_ERR_NOMEM; goto done; }
        for (uint64_t i0 = 0; i0 < n; i0 += BLK) {
            size_t c = (size_t)(n - i0 < BLK ? n - i0 : BLK);
            for (size_t j = 0; j < c; ++j) {
                const uint8_t *r = recs + (i0 + j) * dr_sz;
                memcpy(buf + j * sz, r, 5);
                memcpy(buf + j * sz + 5, r + 8, ixb);
            }
            if (fwrite(buf, sz, c, fo) != c) { rc = UTB_ERR_IO; goto werr; }
        }
    }
    /* label tail: re-emitted through the label reader (itree.c:1270, 1311-1313): distinct labels in order
     * of first appearance, each with its decimal count.  READ_ADD_SAMPLES stores the count of EVERY line
     * into SampCnts[sampIX] (itree.c:1208-1209), sampIX being the label added last: a repeated label
     * overwrites the count of whichever label is newest at that point. */
    {
        uint32_t labels = 0;
        size_t cap = 16, lines = 0;
        for (size_t i = 0; i < tail_len; ++i) lines += tail[i] == '\n';
        while (cap < 2 * (lines + 1) + 2) cap <<= 1;
        const char **seen = (const char **)calloc(cap, sizeof(char *));
        size_t *seen_len = (size_t *)calloc(cap, sizeof(size_t));
        const char **lab = (const char **)malloc((lines + 2) * sizeof(char *));
        size_t *lab_len = (size_t *)malloc((lines + 2) * sizeof(size_t));
        unsigned long long *lab_cnt = (unsigned long long *)malloc((lines + 2) * sizeof(unsigned long long));
#define FREE_LABELS() do { free(seen); free(seen_len); free(lab); free(lab_len); free(lab_cnt); } while (0)
        if (!seen || !seen_len || !lab || !lab_len || !lab_cnt) { FREE_LABELS(); utb_set_error("out of memory"); rc = UTB_ERR_NOMEM; goto done; }
        for (const char *p = (const char *)tail, *e = p + tail_len; p < e;) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
            size_t ll = nl ? (size_t)(nl - p) : (size_t)(e - p);
            const char *tb = (const char *)memchr(p, '\t', ll);
            if (!tb) { FREE_LABELS(); utb_set_error("label line %u has no tab", labels); rc = UTB_ERR_FORMAT; goto done; }
            size_t len = (size_t)(tb - p);
            uint64_t h = 1469598103934665603ull;
            for (size_t k = 0; k < len; ++k) { h ^= (uint8_t)p[k]; h *= 1099511628211ull; }
            size_t s = (size_t)h & (cap - 1);
            int dup = 0;
            while (seen[s]) { if (seen_len[s] == len && !memcmp(seen[s], p, len)) { dup = 1; break; } s = (s + 1) & (cap - 1); }
            if (!dup) { seen[s] = p; seen_len[s] = len; lab[labels] = p; lab_len[labels] = len; ++labels; }
            lab_cnt[labels - 1] = strtoull(tb + 1, NULL, 10);      /* atol + %llu round trip */
            if (!nl) break;
            p = nl + 1;
        }
        uint64_t total = 0;
        for (uint32_t i = 0; i < labels; ++i) {
            total += lab_cnt[i];
            if (fwrite(lab[i], 1, lab_len[i], fo) != lab_len[i] || fprintf(fo, "\t%llu\n", lab_cnt[i]) < 0) { FREE_LABELS(); rc = UTB_ERR_IO; goto werr; }
        }
        FREE_LABELS();
#undef FREE_LABELS
        if (n_labels) *n_labels = labels;
        g_count_sum = total;
    }
    if (n_records) *n_records = n;
    goto done;
werr:
    utb_set_error("write error on %s", ctr_path);
done:
    if (fo && fclose(fo) && !rc) { rc = UTB_ERR_IO; utb_set_error("write error on %s", ctr_path); }
    free(binix); free(binix32); free(buf);
    munmap((void *)base, (size_t)st.st_size);
    close(fd);
    return rc;
}

/* The reference CLI contract of the COMPRESS build (itree.c:1352-1355, 1234-1315). */
int utb_compress_main(int argc, char **argv) {
    if (argc != 3) { puts("[v2.0RF SigNature Edition] usage: xtree-compress preTree.ubt compTree.ctr"); return 1; }
    uint64_t n = 0; uint32_t labels = 0;
    /* header line first, as the reference prints it before compressing */
    {
        FILE *f = fopen(argv[1], "rb");
        uint64_t md[4] = {0, 0, 0, 0};
        if (f && fread(md, 8, 4, f) == 4 && md[3] && md[0] == 8 && md[1] == 0 && (md[2] == 2 || md[2] == 4)) {
            printf("Nodes in input tree: %llu (PACKSIZE=%u, CNTTYPE=%s, IXTYPE=%s, el=%u)\n", (unsigned long long)md[3], 32u,
                   "NA", md[2] == 2 ? "uint16_t" : "uint32_t", (unsigned)(8 + md[2]));
            puts(md[3] < 0xFFFFFFFFull ? "Using 32-bit counters"
                                       : "Holy smokes. Looks like we have over 4 billion k-mers here.\nPlease complain to the developer.\nTrying something anyway...");
        }
        if (f) fclose(f);
    }
    int rc = utb_compress_ubt(argv[1], argv[2], &n, &labels);
    if (rc) { puts(utb_last_error()); return 0; }                  /* the reference exits 0 on every error here */
    (void)n;
    printf("Total nodes in tree: %llu [%llu labels]\n", (unsigned long long)g_count_sum, (unsigned long long)labels);   /* itree.c:1310-1314: the sum of the label counts */
    return 0;
}
