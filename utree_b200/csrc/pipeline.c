/* pipeline.c -- host side of the search: FASTA framer (the reference's record
 * reader XT_INITIATE_WS, itree.c:860-901), batch scheduler over one or more
 * GPUs, and the ordered output formatter (the fprintf lines of itree.c:1032,
 * 1040, 1096).
 *
 * Threads: the calling thread reads the input straight into the pinned
 * staging buffer of the next free batch slot, frames the records in place
 * (no per-read copies: the device receives the file bytes as they are, plus
 * one (offset,length) pair per read) and submits the batch; one formatter
 * thread waits for batches in sequence order and emits the text.  Slots are
 * dealt round-robin over the devices, the database being replicated on each,
 * so multi-GPU needs no collective: the ordered merge below is the only
 * cross-device step.
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include "utb_internal.h"
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

int utb_batch_last_ms(utb_batch *b, float ms[4]);
uint64_t utb_batch_launches(const utb_batch *b);

#define SLOTS_PER_DEVICE 3
#define DEFAULT_BATCH_BYTES ((size_t)64 << 20)

typedef struct {
    utb_batch *b;
    int dev_index;
    /* filled by the framer */
    size_t n_bytes, n_reads;
    uint32_t *name_off, *name_len;
    uint64_t first_read;           /* global index of the slot's first read */
    /* hand-over */
    int state;                     /* 0 free, 1 submitted */
} slot_t;

struct utb_searcher {
    const utb_ctr *ctr;
    int n_devices;
    int *devices;
    utb_db **dbs;
    int n_slots;
    slot_t *slots;
    int host_threads;
    size_t batch_bytes, batch_reads;
    int verbose;                   /* CLI: progress lines on stdout */
};

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int utb_searcher_create(const utb_ctr *ctr, const int *devices, int n_devices,
                        int host_threads, utb_searcher **out) {
    if (!ctr || !out || n_devices < 1 || !devices) { utb_set_error("utb_searcher_create: bad argument"); return UTB_ERR_ARG; }
    *out = NULL;
    utb_searcher *s = (utb_searcher *)calloc(1, sizeof(*s));
    if (!s) { utb_set_error("out of memory"); return UTB_ERR_NOMEM; }
    s->ctr = ctr; s->n_devices = n_devices; s->host_threads = host_threads < 1 ? 1 : host_threads;
    const char *e = getenv("UTB_BATCH_MB");
    s->batch_bytes = e && atoi(e) > 0 ? (size_t)atoi(e) << 20 : DEFAULT_BATCH_BYTES;
    if (s->batch_bytes < 2 * (size_t)UTB_LINELEN + 4096) s->batch_bytes = 2 * (size_t)UTB_LINELEN + 4096; /* one max record must fit */
    s->batch_reads = s->batch_bytes / 64;
    s->devices = (int *)malloc(sizeof(int) * (size_t)n_devices);
    s->dbs = (utb_db **)calloc((size_t)n_devices, sizeof(utb_db *));
    s->n_slots = n_devices * SLOTS_PER_DEVICE;
    s->slots = (slot_t *)calloc((size_t)s->n_slots, sizeof(slot_t));
    if (!s->devices || !s->dbs || !s->slots) { utb_searcher_destroy(s); utb_set_error("out of memory"); return UTB_ERR_NOMEM; }
    for (int d = 0; d < n_devices; ++d) {
        s->devices[d] = devices[d];
        int rc = utb_db_upload(ctr, devices[d], &s->dbs[d]);
        if (rc) { utb_searcher_destroy(s); return rc; }
    }
    for (int i = 0; i < s->n_slots; ++i) {
        slot_t *sl = &s->slots[i];
        sl->dev_index = i % n_devices;
        int rc = utb_batch_create(s->dbs[sl->dev_index], s->batch_bytes, s->batch_reads, &sl->b);
        if (rc) { utb_searcher_destroy(s); return rc; }
        sl->name_off = (uint32_t *)malloc(s->batch_reads * 4);
        sl->name_len = (uint32_t *)malloc(s->batch_reads * 4);
        if (!sl->name_off || !sl->name_len) { utb_searcher_destroy(s); utb_set_error("out of memory"); return UTB_ERR_NOMEM; }
    }
    *out = s;
    return UTB_OK;
}

void utb_searcher_destroy(utb_searcher *s) {
    if (!s) return;
    if (s->slots) for (int i = 0; i < s->n_slots; ++i) {
        utb_batch_destroy(s->slots[i].b);
        free(s->slots[i].name_off); free(s->slots[i].name_len);
    }
    if (s->dbs) for (int d = 0; d < s->n_devices; ++d) utb_db_free(s->dbs[d]);
    free(s->slots); free(s->dbs); free(s->devices); free(s);
}

void utb_searcher_set_verbose(utb_searcher *s, int v) { if (s) s->verbose = v; }

/* ---- input source / output sink ------------------------------------------ */
typedef struct {
    int fd;                 /* >= 0: file */
    const char *mem;        /* else memory */
    size_t mem_len, mem_pos;
    int eof;
} source_t;

static ssize_t src_read(source_t *s, char *dst, size_t cap) {
    if (s->fd >= 0) {
        size_t got = 0;
        while (got < cap) {
            ssize_t k = read(s->fd, dst + got, cap - got);
            if (k < 0) { if (errno == EINTR) continue; return -1; }
            if (k == 0) { s->eof = 1; break; }
            got += (size_t)k;
        }
        return (ssize_t)got;
    }
    size_t left = s->mem_len - s->mem_pos, k = left < cap ? left : cap;
    memcpy(dst, s->mem + s->mem_pos, k);
    s->mem_pos += k;
    if (s->mem_pos == s->mem_len) s->eof = 1;
    return (ssize_t)k;
}

typedef struct {
    FILE *fp;               /* file sink, or */
    char *mem; size_t len, cap;   /* growing memory sink */
    int failed;
} sink_t;

static void sink_write(sink_t *k, const char *p, size_t n) {
    if (!n) return;
    if (k->fp) { if (fwrite(p, 1, n, k->fp) != n) k->failed = 1; return; }
    if (k->len + n > k->cap) {
        size_t nc = k->cap ? k->cap : (size_t)1 << 20;
        while (nc < k->len + n) nc <<= 1;
        char *q = (char *)realloc(k->mem, nc);
        if (!q) { k->failed = 1; return; }
        k->mem = q; k->cap = nc;
    }
    memcpy(k->mem + k->len, p, n);
    k->len += n;
}

/* ---- formatter ------------------------------------------------------------ */
typedef struct {
    utb_searcher *s;
    sink_t *sink;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    uint64_t submitted;     /* batches handed to the devices */
    int done_reading;
    int error;              /* sticky UTB_ERR_* from the device side */
    char errmsg[512];
    utb_stats st;
} run_t;

static inline char *put_u32(char *p, uint32_t v) {
    char t[10]; int n = 0;
    do { t[n++] = (char)('0' + v % 10u); v /= 10u; } while (v);
    while (n) *p++ = t[--n];
    return p;
}

/* One line per read with >= 1 hit; shapes of itree.c:1032, 1040, 1096. */
static size_t format_batch(const utb_ctr *c, const slot_t *sl, const utb_result *res, char *out, uint64_t *good) {
    const char *bytes = utb_batch_bytes(sl->b);
    char *p = out;
    uint64_t g = 0;
    for (size_t r = 0; r < sl->n_reads; ++r) {
        const utb_result *v = &res[r];
        if (v->kind == UTB_NONE) continue;
        ++g;
        memcpy(p, bytes + sl->name_off[r], sl->name_len[r]); p += sl->name_len[r];
        *p++ = '\t';
        const char *lab = c->blob + c->off[v->label];
        size_t ll = c->off[v->label + 1] - c->off[v->label] - 1;
        if (v->kind == UTB_WALK) {
            if (v->cut == UTB_CUT_EMPTY) ll = 0;                   /* dv == -1: "" (itree.c:1087) */
            else if (v->cut != UTB_CUT_FULL && v->cut < ll) ll = v->cut;   /* first dv bytes (itree.c:1088) */
        }
        memcpy(p, lab, ll); p += ll;
        *p++ = '\t';
        p = put_u32(p, v->found); *p++ = '\t';
        p = put_u32(p, v->uix); *p++ = '\t';
        if (v->kind == UTB_STAR) *p++ = '*';
        else { p = put_u32(p, v->sl); *p++ = ';'; p = put_u32(p, v->ol); }
        *p++ = '\n';
    }
    *good += g;
    return (size_t)(p - out);
}

static void *formatter_main(void *arg) {
    run_t *R = (run_t *)arg;
    utb_searcher *s = R->s;
    size_t max_label = 0;
    for (uint32_t i = 0; i < s->ctr->max_ix; ++i) {
        size_t l = s->ctr->off[i + 1] - s->ctr->off[i];
        if (l > max_label) max_label = l;
    }
    char *obuf = NULL; size_t ocap = 0;
    for (uint64_t seq = 0;; ++seq) {
        pthread_mutex_lock(&R->mu);
        while (R->submitted <= seq && !R->done_reading) pthread_cond_wait(&R->cv, &R->mu);
        int have = R->submitted > seq;
        pthread_mutex_unlock(&R->mu);
        if (!have) break;
        slot_t *sl = &s->slots[seq % (uint64_t)s->n_slots];
        const utb_result *res = NULL;
        int rc = utb_batch_wait(sl->b, &res);
        if (rc && !R->error) { R->error = rc; snprintf(R->errmsg, sizeof R->errmsg, "%s", utb_last_error()); }
        if (!rc) {
            /* worst case per line: name + label + 4 tabs + 4 numbers + newline */
            size_t need = sl->n_bytes + sl->n_reads * (max_label + 64);
            if (need > ocap) { free(obuf); ocap = need + (need >> 2); obuf = (char *)malloc(ocap); }
            if (!obuf) { R->error = UTB_ERR_NOMEM; snprintf(R->errmsg, sizeof R->errmsg, "out of memory (formatter)"); ocap = 0; }
            else {
                size_t n = format_batch(s->ctr, sl, res, obuf, &R->st.good_finds);
                sink_write(R->sink, obuf, n);
                R->st.out_bytes += n;
            }
            uint64_t lk = 0, ht = 0; float ms[4] = {0, 0, 0, 0};
            utb_batch_counts(sl->b, &lk, &ht);
            utb_batch_last_ms(sl->b, ms);
            R->st.lookups += lk; R->st.hits += ht;
            R->st.seconds_device += 1e-3 * (double)ms[3];
            R->st.d2h_bytes += sl->n_reads * sizeof(utb_result) + 32;
            if (s->verbose) {   /* itree.c:878 */
                uint64_t a = sl->first_read, b = sl->first_read + sl->n_reads;
                for (uint64_t m = (a >> 20) + 1; (m << 20) <= b; ++m)
                    printf("Searched %llu queries...\n", (unsigned long long)(m << 20));
            }
        }
        pthread_mutex_lock(&R->mu);
        sl->state = 0;
        pthread_cond_broadcast(&R->cv);
        pthread_mutex_unlock(&R->mu);
    }
    free(obuf);
    return NULL;
}

/* ---- framer ---------------------------------------------------------------- */
/* Frames complete records of buf[0..fill) into the slot's tables, following
 * the reference reader (itree.c:866-890).  Returns the number of bytes
 * consumed; *fmt_err is set (with a message) when the record at the returned
 * position is malformed -- the records before it are still valid. */
static size_t frame_records(slot_t *sl, size_t fill, int eof, size_t max_reads, uint64_t max_slots,
                            uint64_t first_read, int *fmt_err, char *msg, size_t msglen, int *full) {
    char *buf = utb_batch_bytes(sl->b);
    uint64_t *seq_off = utb_batch_seq_off(sl->b);
    uint32_t *seq_len = utb_batch_seq_len(sl->b);
    size_t pos = 0, n = 0;
    uint64_t groups = 0;
    *fmt_err = 0; *full = 0;
    while (pos < fill) {
        char *h = buf + pos;
        char *hnl = (char *)memchr(h, '\n', fill - pos);
        if (!hnl && !eof) break;                                   /* header line incomplete */
        size_t sstart = hnl ? (size_t)(hnl - buf) + 1 : fill;
        if (sstart >= fill) {
            if (!eof) break;
            /* fgets for the sequence line fails: itree.c:871-872 */
            snprintf(msg, msglen, "ERROR: can't read sequence L %llu", (unsigned long long)(first_read + n));
            *fmt_err = 1; break;
        }
        char *sq = buf + sstart;
        char *snl = (char *)memchr(sq, '\n', fill - sstart);
        if (!snl && !eof) break;                                   /* sequence line incomplete */
        size_t send = snl ? (size_t)(snl - buf) + 1 : fill;        /* one past the line incl. '\n' */
        if ((size_t)(sstart - pos) >= UTB_LINELEN || send - sstart >= UTB_LINELEN) {
            snprintf(msg, msglen, "ERROR: line longer than %u bytes near query %llu (limit of the reference reader)",
                     UTB_LINELEN - 1, (unsigned long long)(first_read + n + 1));
            *fmt_err = 1; break;
        }
        if (*h != '>') {                                           /* itree.c:880 */
            snprintf(msg, msglen, "ERROR: no header '>' [L %llu]", (unsigned long long)(first_read + n + 1));
            *fmt_err = 1; break;
        }
        if (*sq == '>') {                                          /* itree.c:886 */
            snprintf(msg, msglen, "ERROR: sequence begins '>' [L %llu]", (unsigned long long)(first_read + n + 1));
            *fmt_err = 1; break;
        }
        /* strlen(): an embedded NUL ends the line (itree.c:887) */
        size_t length = send - sstart;
        char *nul = (char *)memchr(sq, 0, length);
        if (nul) length = (size_t)(nul - sq);
        if (!length) {                                             /* itree.c:888 */
            snprintf(msg, msglen, "ERROR: empty query line %llu", (unsigned long long)(first_read + n + 1));
            *fmt_err = 1; break;
        }
        if (sq[length - 1] == '\n') --length;                      /* itree.c:889 */
        if (length && sq[length - 1] == '\r') --length;            /* itree.c:890 */
        uint64_t g = utb_read_slots((uint32_t)length);
        if (n == max_reads || groups + g > max_slots) { *full = 1; break; }
        /* name: after '>' up to the first ' ', '\n' or NUL (itree.c:881-882) */
        size_t hl = (size_t)(sstart - pos);                        /* header line incl. '\n' if any */
        size_t nl = 1;
        while (nl < hl && h[nl] && h[nl] != ' ' && h[nl] != '\n') ++nl;
        sl->name_off[n] = (uint32_t)(pos + 1);
        sl->name_len[n] = (uint32_t)(nl - 1);
        seq_off[n] = sstart;
        seq_len[n] = (uint32_t)length;
        groups += g;
        ++n;
        pos = send;
    }
    sl->n_reads = n;
    return pos;
}

static int run_search(utb_searcher *s, source_t *src, sink_t *sink, int do_rc, utb_stats *stats, int *ref_exit) {
    run_t R;
    memset(&R, 0, sizeof R);
    R.s = s; R.sink = sink;
    pthread_mutex_init(&R.mu, NULL);
    pthread_cond_init(&R.cv, NULL);
    if (ref_exit) *ref_exit = 0;
    double t0 = now_s();
    uint64_t launches0 = 0;
    for (int i = 0; i < s->n_slots; ++i) { s->slots[i].state = 0; launches0 += utb_batch_launches(s->slots[i].b); }
    pthread_t fmt;
    if (pthread_create(&fmt, NULL, formatter_main, &R)) { utb_set_error("cannot start formatter thread"); return UTB_ERR_NOMEM; }

    char *carry = (char *)malloc(s->batch_bytes);
    size_t carry_len = 0;
    int rc = UTB_OK, fmt_err = 0;
    char fmt_msg[256] = "";
    uint64_t seq = 0, n_reads_total = 0;
    if (!carry) { rc = UTB_ERR_NOMEM; utb_set_error("out of memory (carry)"); }
    while (!rc) {
        slot_t *sl = &s->slots[seq % (uint64_t)s->n_slots];
        pthread_mutex_lock(&R.mu);
        while (sl->state != 0) pthread_cond_wait(&R.cv, &R.mu);
        int dev_err = R.error;
        pthread_mutex_unlock(&R.mu);
        if (dev_err) break;
        char *buf = utb_batch_bytes(sl->b);
        size_t cap = utb_batch_max_bytes(sl->b), fill = carry_len;
        if (carry_len) memcpy(buf, carry, carry_len);
        carry_len = 0;
        if (!src->eof) {
            ssize_t k = src_read(src, buf + fill, cap - fill);
            if (k < 0) { rc = UTB_ERR_IO; utb_set_error("read error on input: %s", strerror(errno)); break; }
            fill += (size_t)k;
        }
        if (!fill) break;                                          /* clean EOF */
        int full = 0;
        size_t used = frame_records(sl, fill, src->eof, utb_batch_max_reads(sl->b), utb_batch_max_slots(sl->b),
                                    n_reads_total, &fmt_err, fmt_msg, sizeof fmt_msg, &full);
        if (!fmt_err && used == 0 && !full && fill == cap) {
            fmt_err = 1;
            snprintf(fmt_msg, sizeof fmt_msg, "ERROR: record larger than the %zu-byte batch buffer", cap);
        }
        sl->n_bytes = used;
        sl->first_read = n_reads_total;
        n_reads_total += sl->n_reads;
        if (!fmt_err && used < fill) { carry_len = fill - used; memcpy(carry, buf + used, carry_len); }
        if (sl->n_reads) {
            int r2 = utb_batch_submit(sl->b, used, sl->n_reads, do_rc);
            if (r2) { rc = r2; break; }
            R.st.h2d_bytes += used + sl->n_reads * 16 + 4;
            pthread_mutex_lock(&R.mu);
            sl->state = 1;
            R.submitted = ++seq;
            pthread_cond_broadcast(&R.cv);
            pthread_mutex_unlock(&R.mu);
        }
        if (fmt_err) break;
        if (src->eof && !carry_len) break;
    }
    pthread_mutex_lock(&R.mu);
    R.done_reading = 1;
    pthread_cond_broadcast(&R.cv);
    pthread_mutex_unlock(&R.mu);
    pthread_join(fmt, NULL);
    free(carry);
    pthread_mutex_destroy(&R.mu);
    pthread_cond_destroy(&R.cv);
    if (!rc && R.error) { rc = R.error; utb_set_error("%s", R.errmsg); }
    if (!rc && sink->failed) { rc = UTB_ERR_IO; utb_set_error("write error on output"); }
    R.st.reads = n_reads_total + (fmt_err ? 1 : 0);                /* the reference counts the bad record too */
    R.st.batches = seq;
    for (int i = 0; i < s->n_slots; ++i) R.st.kernel_launches += utb_batch_launches(s->slots[i].b);
    R.st.kernel_launches -= launches0;
    R.st.seconds_total = now_s() - t0;
    if (stats) *stats = R.st;
    if (!rc && fmt_err) {
        utb_set_error("%s", fmt_msg);
        if (ref_exit) *ref_exit = 2;                               /* itree.c:872-888 */
        return UTB_ERR_FORMAT;
    }
    return rc;
}

int utb_search_file(utb_searcher *s, const char *fasta_path, const char *out_path,
                    int do_rc, utb_stats *stats, int *ref_exit) {
    if (!s || !fasta_path || !out_path) { utb_set_error("utb_search_file: null argument"); return UTB_ERR_ARG; }
    if (ref_exit) *ref_exit = 0;
    int fd = open(fasta_path, O_RDONLY);
    if (fd < 0) { utb_set_error("Invalid input files"); if (ref_exit) *ref_exit = 1; return UTB_ERR_IO; }   /* itree.c:835 */
#ifdef POSIX_FADV_SEQUENTIAL
    posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
    FILE *fo = fopen(out_path, "wb");
    if (!fo) { close(fd); utb_set_error("cannot open output file %s", out_path); if (ref_exit) *ref_exit = 1; return UTB_ERR_IO; }
    setvbuf(fo, NULL, _IOFBF, (size_t)4 << 20);
    source_t src; memset(&src, 0, sizeof src); src.fd = fd;
    sink_t sink; memset(&sink, 0, sizeof sink); sink.fp = fo;
    int rc = run_search(s, &src, &sink, do_rc, stats, ref_exit);
    if (fclose(fo) && !rc) { rc = UTB_ERR_IO; utb_set_error("write error on output"); }
    close(fd);
    return rc;
}

int utb_search_mem(utb_searcher *s, const char *fasta, size_t n, int do_rc,
                   char **out, size_t *out_len, utb_stats *stats, int *ref_exit) {
    if (!s || (!fasta && n) || !out || !out_len) { utb_set_error("utb_search_mem: null argument"); return UTB_ERR_ARG; }
    source_t src; memset(&src, 0, sizeof src); src.fd = -1; src.mem = fasta; src.mem_len = n; src.eof = n == 0;
    sink_t sink; memset(&sink, 0, sizeof sink);
    int rc = run_search(s, &src, &sink, do_rc, stats, ref_exit);
    *out = sink.mem; *out_len = sink.len;
    return rc;
}

/* ---- CLI (itree.c:1357-1377; stdout lines of SURVEY App. C) ---------------- */
static const char *TYPEARR[9] = {"NA", "uint8_t", "uint16_t", "NA", "uint32_t", "NA", "NA", "NA", "uint64_t"};

int utb_main(int argc, char **argv) {
    if (argc < 4) {                                                /* itree.c:1358-1360 */
        printf("[v2.0RF SigNature Edition] usage: xtree-searchGG compTree.ctr fastaToSearch.fa output.txt [threads] [SPEED <X>] [RC]\n");
        return 1;
    }
    printf("This is UTree [v2.0RF SigNature Edition]\n");
    int do_rc = !strcmp(argv[argc - 1], "RC");                     /* itree.c:1362-1364 */
    argc -= do_rc;
    int speed = 0;
    if (!strcmp(argv[argc - 2], "SPEED")) { speed = atoi(argv[argc - 1]); argc -= 2; }
    printf("Reverse complement consideration is %sabled.\n", do_rc ? "en" : "dis");
    printf("Searching at speed %d.\n", speed);
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    int threads = argc >= 5 ? atoi(argv[4]) : (int)(ncpu > 0 ? ncpu : 1);   /* itree.c:1368 */
    printf("Using up to %d threads.\n", threads);

    utb_ctr *ctr = NULL;
    int rc = utb_ctr_open(argv[1], &ctr);
    if (rc) { puts(utb_last_error()); return 0; }                  /* itree.c:735, 738, 750: exit(0) */
    if (ctr->num_nodes < 0xFFFFFFFFull) puts("Using 32-bit counters");   /* itree.c:754-755 */
    else puts("Holey smokes, a tree of over 4 billion k-mers. Here goes...");
    printf("%llu elements read.\n", (unsigned long long)UTB_NUMBINS);     /* itree.c:761 */
    printf("Nodes in input tree: %llu (PACKSIZE=%u, CNTTYPE=%s, IXTYPE=%s, SZ=%d)\n",   /* itree.c:764 */
           (unsigned long long)ctr->num_nodes, 32u, "NA", TYPEARR[ctr->ix_bytes], (int)ctr->sz);
    printf("Read %llu nodes.\n", (unsigned long long)ctr->num_nodes);     /* itree.c:769 */
    if (ctr->last_bin != ctr->num_nodes)                           /* itree.c:792-793 */
        printf("Warning: detected nodes %u != %u\n", (unsigned)ctr->last_bin, (unsigned)ctr->num_nodes);

    int ndev = 0, devs[64];
    rc = utb_device_count(&ndev);
    if (rc) { fprintf(stderr, "utree-b200: %s\n", utb_last_error()); utb_ctr_close(ctr); return 4; }
    int n = 0;
    const char *e = getenv("UTB_DEVICES");                         /* e.g. "0,1,2,3"; default: all visible */
    if (e && *e) {
        char *dup = strdup(e), *save = NULL;
        for (char *t = strtok_r(dup, ",", &save); t && n < 64; t = strtok_r(NULL, ",", &save)) devs[n++] = atoi(t);
        free(dup);
    } else for (; n < ndev && n < 64; ++n) devs[n] = n;
    utb_searcher *s = NULL;
    rc = utb_searcher_create(ctr, devs, n, threads, &s);
    if (rc) { fprintf(stderr, "utree-b200: %s\n", utb_last_error()); utb_ctr_close(ctr); return rc == UTB_ERR_NOMEM ? 3 : 4; }
    puts("Tree read.");                                            /* itree.c:826 */
    fflush(stdout);
    s->verbose = 1;
    utb_stats st;
    int ref_exit = 0;
    rc = utb_search_file(s, argv[2], argv[3], do_rc, &st, &ref_exit);
    if (rc == UTB_ERR_IO && ref_exit == 1) { puts(utb_last_error()); utb_searcher_destroy(s); utb_ctr_close(ctr); return 1; }
    if (rc == UTB_ERR_FORMAT && ref_exit) {                        /* output so far is flushed, as exit() does */
        fflush(stdout);
        fprintf(stderr, "%s\n", utb_last_error());
        utb_searcher_destroy(s); utb_ctr_close(ctr);
        return ref_exit;
    }
    if (rc) { fprintf(stderr, "utree-b200: %s\n", utb_last_error()); utb_searcher_destroy(s); utb_ctr_close(ctr); return 4; }
    printf("Good finds: %llu\n", (unsigned long long)st.good_finds);       /* itree.c:1106 */
    printf("Searched %llu queries\n", (unsigned long long)st.reads);       /* itree.c:1375 */
    if (getenv("UTB_STATS"))
        fprintf(stderr, "utree-b200: %d GPU(s), %llu batches, %llu lookups, %llu hits, %.3f s total, %.3f s device\n",
                n, (unsigned long long)st.batches, (unsigned long long)st.lookups, (unsigned long long)st.hits,
                st.seconds_total, st.seconds_device);
    utb_searcher_destroy(s);
    utb_ctr_close(ctr);
    return 0;
}
