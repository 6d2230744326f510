/* oracle.c -- CPU restatement of UTree SEARCH_GG (see oracle.h header note:
 * TEST INFRASTRUCTURE, never linked into the product).
 *
 * Written from the algorithm description in SURVEY.md App. A/B and checked
 * line-by-line against /root/reference/itree.c; the citations below name the
 * reference lines each block restates.  It deliberately keeps the reference's
 * in-memory layout (u64-widened BinIx, byte-packed records with unaligned
 * 8-byte loads) so that sector accounting is in terms of the on-disk format.
 */
#define _FILE_OFFSET_BITS 64
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NUMBINS ((1u << 24) + 1u)          /* itree.c:693 */
#define SXBITS 40                          /* itree.c:694 */
#define SUFMASK 0xFFFFFFFFFFull            /* itree.c:780-783 */
#define LINELEN 16777216                   /* itree.c:836 */
#define BAD32 0xFFFFFFFFu

struct OrcDB {
    uint64_t num_nodes;
    uint32_t ix_bytes;     /* 2 or 4 */
    uint32_t sz;           /* 5 + ix_bytes (itree.c:691) */
    uint32_t binix_bytes;  /* 4 or 8 on disk (itree.c:757) */
    uint64_t *binix;       /* widened (itree.c:756-759) */
    uint8_t *dump;         /* records + 32 slack (itree.c:766) */
    uint32_t max_ix;       /* sampIX + 1 */
    char **labels;         /* SampStrings */
    char *label_blob;
};

static void seterr(char *err, size_t n, const char *msg) {
    if (err && n) { strncpy(err, msg, n - 1); err[n - 1] = 0; }
}

/* ---- label tail: itree.c:1154-1223 READ_ADD_SAMPLES / addSampleUdX -------
 * Each line "label\tcount\n"; the label is the bytes before the first tab.
 * A label string seen before does NOT get a new id (addSampleUdX returns the
 * existing one, itree.c:219-220), so ids number the DISTINCT labels in order
 * of first appearance.  The reference keeps a BST; we use an open hash. */
static uint64_t hash_str(const char *s, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= (uint8_t)s[i]; h *= 1099511628211ull; }
    return h;
}

static int parse_labels(OrcDB *db, char *blob, size_t n) {
    size_t lines = 0;
    for (size_t i = 0; i < n; ++i) lines += blob[i] == '\n';
    if (n && blob[n - 1] != '\n') ++lines;
    db->labels = (char **)calloc(lines + 1, sizeof(char *));
    size_t cap = 16; while (cap < 2 * lines + 2) cap <<= 1;
    uint32_t *tab = (uint32_t *)malloc(cap * sizeof(uint32_t));
    memset(tab, 0xFF, cap * sizeof(uint32_t));
    uint32_t count = 0;
    size_t pos = 0;
    while (pos < n) {
        char *line = blob + pos;
        char *nl = (char *)memchr(line, '\n', n - pos);
        size_t linelen = nl ? (size_t)(nl - line) : n - pos;
        char *tab_c = (char *)memchr(line, '\t', linelen);
        if (!tab_c) { free(tab); return -1; } /* reference would run off the line */
        size_t ll = (size_t)(tab_c - line);
        *tab_c = 0;
        uint64_t h = hash_str(line, ll) & (cap - 1);
        int dup = 0;
        while (tab[h] != BAD32) {
            if (!strcmp(db->labels[tab[h]], line)) { dup = 1; break; }
            h = (h + 1) & (cap - 1);
        }
        if (!dup) { tab[h] = count; db->labels[count++] = line; }
        pos += linelen + 1;
    }
    free(tab);
    db->max_ix = count; /* sampIX = count-1; maxIX = sampIX+1 (itree.c:855) */
    return 0;
}

OrcDB *orc_db_load(const char *path, char *err, size_t errlen) {
    FILE *fp = fopen(path, "rb");
    if (!fp) { seterr(err, errlen, "Invalid DB file"); return NULL; }  /* :735 */
    uint64_t md[4] = {0, 0, 0, 0};
    if (fread(md, 8, 4, fp) < 4 || !md[3]) {                            /* :737-738 */
        fclose(fp); seterr(err, errlen, "Tree malformatted."); return NULL;
    }
    /* :746-751 -- the reference binary is compiled for one IXTYPE; this
     * restatement dispatches on the header instead (PACKSIZE=32, NO_COUNT). */
    if (md[0] != 8 || md[1] != 0 || (md[2] != 2 && md[2] != 4)) {
        fclose(fp); seterr(err, errlen, "ERROR. Input tree requires other PACKSIZE/CNTTYPE/IXTYPE");
        return NULL;
    }
    OrcDB *db = (OrcDB *)calloc(1, sizeof(*db));
    db->num_nodes = md[3];
    db->ix_bytes = (uint32_t)md[2];
    db->sz = 5 + db->ix_bytes;
    db->binix_bytes = db->num_nodes < 0xFFFFFFFFull ? 4 : 8;          /* :757 */
    db->binix = (uint64_t *)calloc(NUMBINS, 8);                         /* zero-extend explicitly */
    if (db->binix_bytes == 4) {
        uint32_t *tmp = (uint32_t *)malloc((size_t)NUMBINS * 4);
        if (fread(tmp, 4, NUMBINS, fp) != NUMBINS) { free(tmp); goto bad; }
        for (size_t i = 0; i < NUMBINS; ++i) db->binix[i] = tmp[i];
        free(tmp);
    } else if (fread(db->binix, 8, NUMBINS, fp) != NUMBINS) goto bad;
    db->dump = (uint8_t *)calloc(db->num_nodes * db->sz + 32, 1);       /* :766 */
    if (fread(db->dump, db->sz, db->num_nodes, fp) != db->num_nodes) goto bad; /* :767-768 */
    {
        off_t here = ftello(fp); fseeko(fp, 0, SEEK_END);
        off_t end = ftello(fp); fseeko(fp, here, SEEK_SET);
        size_t n = (size_t)(end - here);
        db->label_blob = (char *)malloc(n + 1);
        if (fread(db->label_blob, 1, n, fp) != n) goto bad;
        db->label_blob[n] = 0;
        if (parse_labels(db, db->label_blob, n)) goto bad;
    }
    fclose(fp);
    return db;
bad:
    fclose(fp);
    seterr(err, errlen, "Error in reading tree.");
    orc_db_free(db);
    return NULL;
}

void orc_db_free(OrcDB *db) {
    if (!db) return;
    free(db->binix); free(db->dump); free(db->labels); free(db->label_blob); free(db);
}
uint64_t orc_db_num_nodes(const OrcDB *db) { return db->num_nodes; }
uint32_t orc_db_max_ix(const OrcDB *db) { return db->max_ix; }
uint32_t orc_db_ix_bytes(const OrcDB *db) { return db->ix_bytes; }
uint32_t orc_db_binix_bytes(const OrcDB *db) { return db->binix_bytes; }
const char *orc_db_label(const OrcDB *db, uint32_t ix) { return ix < db->max_ix ? db->labels[ix] : NULL; }
const uint64_t *orc_db_binix(const OrcDB *db) { return db->binix; }
const uint8_t *orc_db_records(const OrcDB *db) { return db->dump; }

/* ---- sector accounting -------------------------------------- */
typedef struct { uint64_t s[96]; int n; } SectSet;
static void sect_add(SectSet *ss, uint64_t lo_byte, uint64_t hi_byte_incl) {
    for (uint64_t s = lo_byte >> 5; s <= (hi_byte_incl >> 5); ++s) {
        int seen = 0;
        for (int i = 0; i < ss->n; ++i) if (ss->s[i] == s) { seen = 1; break; }
        if (!seen && ss->n < 96) ss->s[ss->n++] = s;
    }
}

static inline uint64_t load_suffix(const uint8_t *dump, uint32_t sz, uint64_t rec) {
    uint64_t v; memcpy(&v, dump + rec * sz, 8);                        /* :676 SUFFIX_AT */
    return v & SUFMASK;
}

/* itree.c:720-730 + 699-707 */
uint32_t orc_lookup(const OrcDB *db, uint64_t word, OrcStats *st) {
    uint64_t p = word >> SXBITS, s = word & SUFMASK;                    /* :722 */
    uint64_t a = db->binix[p], b = db->binix[p + 1];                    /* :724 */
    SectSet sb; sb.n = 0;
    if (st) {
        SectSet si; si.n = 0;
        sect_add(&si, p * db->binix_bytes, (p + 2) * db->binix_bytes - 1);
        st->sect_idx += (uint64_t)si.n;
        st->lookups++;
    }
    if (a >= b) return BAD32;                                           /* :726 */
    uint64_t pos = a, size = b - a - 1;                                 /* :728 */
    while (size) {                                                      /* :701-705 */
        uint64_t w = size >> 1;
        if (st) { sect_add(&sb, (pos + w + 1) * db->sz, (pos + w + 1) * db->sz + 7); st->probes++; }
        if (load_suffix(db->dump, db->sz, pos + w + 1) <= s) { pos += w + 1; size -= w + 1; }
        else size = w;
    }
    if (st) { sect_add(&sb, pos * db->sz, pos * db->sz + 7); st->probes++; }
    uint32_t ix = BAD32;
    if (load_suffix(db->dump, db->sz, pos) == s) {                      /* :706 */
        const uint8_t *r = db->dump + pos * db->sz + 5;                 /* :677 IX_AT */
        if (db->ix_bytes == 2) { uint16_t v; memcpy(&v, r, 2); ix = v; }
        else memcpy(&ix, r, 4);
        if (st) sect_add(&sb, pos * db->sz + 5, pos * db->sz + db->sz - 1);
        /* BAD_IX of a uint16_t build is 0xFFFF; it can never be < maxIX
         * (<= 0xFFFE labels), so widening is transparent. */
    }
    if (st) st->sect_bkt += (uint64_t)sb.n;
    return ix;
}

/* ---- 2-bit coding: itree.c:110-121 --------------------------------------- */
static uint8_t C2X[256];
static char RCT[256];
static int tables_ready = 0;
static void init_tables(void) {
    if (tables_ready) return;
    memset(C2X, 255, 256);
    C2X['a'] = C2X['A'] = 0; C2X['c'] = C2X['C'] = 1;
    C2X['g'] = C2X['G'] = 2; C2X['t'] = C2X['T'] = 3;
    memset(RCT, 'N', 256);                                              /* :838-841 */
    RCT['A'] = RCT['a'] = 'T'; RCT['C'] = RCT['c'] = 'G';
    RCT['G'] = RCT['g'] = 'C'; RCT['T'] = RCT['t'] = 'A';
    tables_ready = 1;
}

uint64_t orc_revcomp_word(uint64_t w) {
    uint64_t r = 0; w = ~w;
    for (int i = 0; i < 32; ++i) { r = (r << 2) | (w & 3); w >>= 2; }
    return r;
}
int orc_pack_word(const char *b, uint64_t *out) {
    init_tables();
    uint64_t w = 0;
    for (int i = 0; i < 32; ++i) {
        uint8_t c = C2X[(uint8_t)b[i]];
        if (c == 255) return 0;
        w = (w << 2) | c;
    }
    *out = w; return 1;
}

/* itree.c:887-898 + 906-933.  Bytes >= 0x80 index C2Xb/RC with a negative
 * char in the reference (UB); scope is 7-bit input and they are treated as
 * non-ACGT here. */
static uint64_t slide_impl(const OrcDB *db, const char *seq, uint32_t len0, int do_rc,
                           uint32_t *hits, uint64_t cap,
                           uint64_t *words, uint64_t cap_words, uint64_t *n_words,
                           OrcStats *st, int skip) {
    init_tables();
    char *buf = NULL; const char *src = seq; int length = (int)len0;
    if (do_rc) {                                                        /* :891-897 */
        buf = (char *)malloc((size_t)len0 * 2 + 2);
        memcpy(buf, seq, len0);
        buf[len0] = 'N';
        for (int x = length + 1; x <= length << 1; ++x)
            buf[x] = RCT[(uint8_t)buf[length + length - x]];
        length = (length << 1) + 1;
        buf[length] = 0;
        src = buf;
    }
    uint64_t found = 0, nw = 0, w = 0;
    const int k1 = 31, kv = 31;
    for (int i = kv, z = -4; i < length; ++i) {                         /* :906 */
        int j;
        if (i < z + kv) { w <<= (i - z - 1) << 1; j = z + 1; }          /* :920 */
        else { w = 0; j = i - k1; }                                     /* :921 */
        for (int p = j; j <= i; ++j) {                                  /* :922-925 */
            uint8_t c = C2X[(uint8_t)src[j]];
            if (c == 255) { i += j - p; z = 0; break; }
            w <<= 2; w |= c;
        }
        if (j <= i) continue;                                           /* :926 */
        z = i;
        if (words && nw < cap_words) words[nw] = w;
        ++nw;
        uint32_t ix = orc_lookup(db, w, st);                            /* :928 */
        if (ix < db->max_ix) {                                          /* :929 */
            if (hits && found < cap) hits[found] = ix;
            ++found;
            if (st) st->hits++;
            i += skip;                                                  /* XT_SHALLOWVOTE :950: PACKSIZE/SPARSITY - 1 (0 for the GG vote) */
        }
    }
    if (n_words) *n_words = nw;
    free(buf);
    return found;
}
uint64_t orc_slide(const OrcDB *db, const char *seq, uint32_t len0, int do_rc,
                   uint32_t *hits, uint64_t cap,
                   uint64_t *words, uint64_t cap_words, uint64_t *n_words,
                   OrcStats *st) {
    return slide_impl(db, seq, len0, do_rc, hits, cap, words, cap_words, n_words, st, 0);
}
/* the non-GG binary (-D SEARCH): after a hit the next window looked up ends 8 bases on */
uint64_t orc_slide_shallow(const OrcDB *db, const char *seq, uint32_t len0, int do_rc, uint32_t *hits, uint64_t cap, OrcStats *st) {
    return slide_impl(db, seq, len0, do_rc, hits, cap, NULL, 0, NULL, st, ORC_SHALLOW_SKIP);
}

/* ---- vote: itree.c:1028-1098 --------------------------------------------- */
typedef struct { const char *s; uint32_t n; uint32_t ix; } TaxCnt;
static int by_str(const void *a, const void *b) {                      /* :831-832 */
    return strcmp(((const TaxCnt *)a)->s, ((const TaxCnt *)b)->s);
}
static inline uint32_t cutoff_of(uint32_t x) {                          /* :1044-1046 */
    uint32_t c = x - x / 4;
    c += ((x >> 1) >= c);
    return c;
}

void orc_vote(const OrcDB *db, const uint32_t *hits, uint64_t n64, OrcVote *out) {
    memset(out, 0, sizeof(*out));
    uint32_t n = (uint32_t)n64;
    if (!n) { out->kind = ORC_NONE; return; }                           /* :1028 */
    out->found = n;
    if (n == 1) {                                                       /* :1031-1032 */
        out->kind = ORC_STAR; out->label = hits[0]; out->uix = 1; return;
    }
    uint32_t *H = (uint32_t *)calloc(db->max_ix, 4);
    TaxCnt *T = (TaxCnt *)malloc((size_t)db->max_ix * sizeof(TaxCnt));
    for (uint32_t i = 0; i < n; ++i) ++H[hits[i]];                       /* :1033-1034 */
    uint32_t uix = 0;
    for (uint32_t i = n; i; --i) {                                       /* :1036-1038 */
        uint32_t t = hits[i - 1];
        if (H[t]) { T[uix].s = db->labels[t]; T[uix].n = H[t]; T[uix].ix = t; ++uix; H[t] = 0; }
    }
    out->uix = uix;
    if (uix == 1) {                                                     /* :1039-1040 */
        out->kind = ORC_STAR; out->label = hits[0]; free(H); free(T); return;
    }
    qsort(T, uix, sizeof(*T), by_str);                                  /* :1041 */
    uint32_t cutoff = cutoff_of(n), st = 0, ed = uix, dv = 0xFFFFFFFFu, orun = n, sl, ol;
    for (;;) {                                                          /* :1047 */
        uint32_t run = T[st].n, td = dv;
        for (uint32_t z = st + 1; z < ed; ++z) {                         /* :1050 */
            const char *s1 = T[z - 1].s, *s2 = T[z].s;
            if (!s1[(uint32_t)(dv + (dv == 0xFFFFFFFFu))]) {            /* :1052 */
                run = T[z].n; st = z;
                orun -= T[z - 1].n;
                cutoff = cutoff_of(orun);
                continue;
            }
            for (td = dv + 1; s1[td] && s1[td] == s2[td]; ++td)        /* :1060-1061 */
                if (s1[td] == ';') break;
            if (s1[td] == s2[td]) run += T[z].n;                        /* :1062 */
            else if ((!s1[td] && s2[td] == ';') ||
                     ((s1[td] == ';' || !s1[td]) && td > 0 && s1[td - 1] == '_')) { /* :1063; td==0 reads s1[-1] in the reference (out of contract) */
                run = T[z].n; st = z;
                orun -= T[z - 1].n;
                cutoff = cutoff_of(orun);
            }
            else if (run >= cutoff) { ed = z; break; }                  /* :1068 */
            else { run = T[z].n; st = z; }                              /* :1069 */
        }
        sl = run; ol = orun;                                            /* :1071 */
        if (run < cutoff) break;                                        /* :1072 */
        if (st + 1 >= ed) {                                             /* :1073-1080 */
            if (T[ed - 1].n >= cutoff) dv = 0xFFFFFFFEu;
            break;
        }
        orun = run; dv = td; cutoff = cutoff_of(run);                   /* :1082-1085 */
    }
    out->kind = ORC_WALK; out->label = T[ed - 1].ix; out->cut = dv;
    out->sl = sl; out->ol = ol;
    free(H); free(T);
}

static size_t put_u32(char *p, uint32_t v) {
    char t[12]; int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; ++i) p[i] = t[n - 1 - i];
    return (size_t)n;
}

/* itree.c:1032, 1040, 1087-1096 */
size_t orc_format(const OrcDB *db, const char *name, const OrcVote *v, char *buf) {
    if (v->kind == ORC_NONE) return 0;
    char *p = buf;
    size_t nl = strlen(name); memcpy(p, name, nl); p += nl; *p++ = '\t';
    const char *lab = db->labels[v->label];
    size_t ll = strlen(lab);
    if (v->kind == ORC_WALK) {
        if (v->cut == ORC_DV_EMPTY) ll = 0;
        else if (v->cut != ORC_DV_FULL && v->cut < ll) ll = v->cut;     /* memcpy dv bytes then %s stops at NUL */
    }
    memcpy(p, lab, ll); p += ll; *p++ = '\t';
    p += put_u32(p, v->found); *p++ = '\t';
    p += put_u32(p, v->uix); *p++ = '\t';
    if (v->kind == ORC_STAR) *p++ = '*';
    else { p += put_u32(p, v->sl); *p++ = ';'; p += put_u32(p, v->ol); }
    *p++ = '\n';
    return (size_t)(p - buf);
}

/* ---- whole-file driver: itree.c:860-901 reader + 1009-1101 ---------------- */
typedef struct { char *name; char *seq; uint32_t len; } Rec;

int orc_search_file(const OrcDB *db, const char *fasta, const char *out,
                    int do_rc, int threads, uint64_t max_reads,
                    OrcStats *stats, char *err, size_t errlen) {
    init_tables();
    FILE *fp = fopen(fasta, "rb");
    if (!fp) { seterr(err, errlen, "Invalid input files"); return 1; }  /* :835 */
    FILE *fo = fopen(out, "wb");
    if (!fo) { fclose(fp); seterr(err, errlen, "cannot open output"); return 1; }
    if (threads < 1) threads = 1;
    char *line = (char *)malloc((size_t)LINELEN + 2), *line2 = (char *)malloc((size_t)LINELEN + 2);
    const size_t CHUNK = 16384;
    Rec *recs = (Rec *)calloc(CHUNK, sizeof(Rec));
    char **outs = (char **)calloc(CHUNK, sizeof(char *));
    size_t *outn = (size_t *)calloc(CHUNK, sizeof(size_t));
    OrcStats tot; memset(&tot, 0, sizeof(tot));
    uint64_t li = 0; int rc = 0, eof = 0;
    while (!eof && !rc) {
        size_t nrec = 0;
        while (nrec < CHUNK) {
            if (max_reads && li >= max_reads) { eof = 1; break; }
            if (!fgets(line, LINELEN, fp)) { eof = 1; break; }          /* :869 */
            if (!fgets(line2, LINELEN, fp)) {                            /* :871-872 */
                char m[96]; snprintf(m, sizeof m, "ERROR: can't read sequence L %u", (unsigned)li);
                seterr(err, errlen, m); rc = 2; break;
            }
            ++li;                                                       /* :877 */
            if (line[0] != '>') { seterr(err, errlen, "ERROR: no header '>'"); rc = 2; break; } /* :880 */
            char *src = line;
            while (*++src && *src != ' ' && *src != '\n');              /* :881 */
            *src = 0;                                                   /* :882 */
            if (line2[0] == '>') { seterr(err, errlen, "ERROR: sequence begins '>'"); rc = 2; break; } /* :886 */
            int length = (int)strlen(line2);                            /* :887 */
            if (!length) { seterr(err, errlen, "ERROR: empty query line"); rc = 2; break; }   /* :888 */
            if (line2[length - 1] == '\n') --length;                    /* :889 */
            if (length > 0 && line2[length - 1] == '\r') --length;      /* :890 (length==0 reads line2[-1] in the reference) */
            Rec *r = &recs[nrec++];
            r->name = strdup(line + 1);
            r->seq = (char *)malloc((size_t)length + 1);
            memcpy(r->seq, line2, (size_t)length); r->seq[length] = 0;
            r->len = (uint32_t)length;
        }
        /* records parsed before a format error are still searched and
         * written, as in the reference (it dies at the bad record). */
        #pragma omp parallel num_threads(threads)
        {
            OrcStats ls; memset(&ls, 0, sizeof(ls));
            #pragma omp for schedule(dynamic, 16)
            for (long r = 0; r < (long)nrec; ++r) {
                Rec *R = &recs[r];
                uint64_t cap = (uint64_t)R->len * 2 + 2;
                uint32_t *hits = (uint32_t *)malloc(cap * 4);
                uint64_t nf = orc_slide(db, R->seq, R->len, do_rc, hits, cap, NULL, 0, NULL, &ls);
                OrcVote v; orc_vote(db, hits, nf, &v);
                free(hits);
                outs[r] = NULL; outn[r] = 0;
                if (v.kind != ORC_NONE) {
                    ls.good_finds++;
                    outs[r] = (char *)malloc(strlen(R->name) + strlen(db->labels[v.label]) + 64);
                    outn[r] = orc_format(db, R->name, &v, outs[r]);
                }
            }
            #pragma omp critical
            {
                tot.good_finds += ls.good_finds; tot.lookups += ls.lookups; tot.hits += ls.hits;
                tot.probes += ls.probes; tot.sect_idx += ls.sect_idx; tot.sect_bkt += ls.sect_bkt;
            }
        }
        for (size_t r = 0; r < nrec; ++r) {
            if (outs[r]) { fwrite(outs[r], 1, outn[r], fo); tot.out_bytes += outn[r]; free(outs[r]); }
            free(recs[r].name); free(recs[r].seq);
        }
    }
    tot.reads = li;
    if (stats) *stats = tot;
    free(recs); free(outs); free(outn); free(line); free(line2);
    fclose(fp); fclose(fo);
    return rc;
}

/* ---- the non-GG search binary (-D SEARCH): itree.c:860-901 reader, :903-933 slide with the SPARSITY
 * skip (:948-951), shallow vote :969-1007.  Single-threaded in the reference (no omp parallel around it),
 * and order matters: AllTheKingsHorses is allocated once (:970) and `if (!kingsMen++)` (:982) both
 * never fires (kingsMen == foundUniq >= 1 there) and bumps the count, so the tally loops (:984-997)
 * run over foundUniq + 1 entries -- the last one is whatever an EARLIER read left at that index of
 * the array (0 from the fresh allocation if none did). */
int orc_search_file_shallow(const OrcDB *db, const char *fasta, const char *out, int do_rc,
                            OrcStats *stats, char *err, size_t errlen) {
    init_tables();
    FILE *fp = fopen(fasta, "rb");
    if (!fp) { seterr(err, errlen, "Invalid input files"); return 1; }  /* :835 */
    FILE *fo = fopen(out, "wb");
    if (!fo) { fclose(fp); seterr(err, errlen, "cannot open output"); return 1; }
    char *line = (char *)malloc((size_t)LINELEN + 2), *line2 = (char *)malloc((size_t)LINELEN + 2);
    uint32_t *horses = (uint32_t *)calloc((size_t)LINELEN * 2, sizeof(uint32_t));    /* :970 (fresh pages read as 0) */
    uint32_t *tally = (uint32_t *)calloc(db->max_ix ? db->max_ix : 1, sizeof(uint32_t));   /* :971 Hashes */
    OrcStats tot; memset(&tot, 0, sizeof(tot));
    uint64_t li = 0; int rc = 0;
    for (;;) {
        if (!fgets(line, LINELEN, fp)) break;                           /* :869 */
        if (!fgets(line2, LINELEN, fp)) {                                /* :871-872 */
            char m[96]; snprintf(m, sizeof m, "ERROR: can't read sequence L %u", (unsigned)li);
            seterr(err, errlen, m); rc = 2; break;
        }
        ++li;
        if (line[0] != '>') { seterr(err, errlen, "ERROR: no header '>'"); rc = 2; break; }
        char *src = line;
        while (*++src && *src != ' ' && *src != '\n');
        *src = 0;
        if (line2[0] == '>') { seterr(err, errlen, "ERROR: sequence begins '>'"); rc = 2; break; }
        int length = (int)strlen(line2);
        if (!length) { seterr(err, errlen, "ERROR: empty query line"); rc = 2; break; }
        if (line2[length - 1] == '\n') --length;
        if (length > 0 && line2[length - 1] == '\r') --length;
        const uint64_t n = orc_slide_shallow(db, line2, (uint32_t)length, do_rc, horses, (uint64_t)LINELEN * 2, &tot);   /* :976-978 */
        if (!n) continue;                                               /* :979 */
        ++tot.good_finds;                                               /* :980 */
        const uint64_t men = n + 1;                                     /* :982: kingsMen++ */
        for (uint64_t i = 0; i < men; ++i) ++tally[horses[i]];          /* :984-985 */
        uint32_t most = 0, second = 0, most_ix = 0;
        for (uint64_t i = 0; i < men; ++i) {                            /* :988-997 */
            const uint32_t h = tally[horses[i]];
            if (h > most) { second = most; most_ix = horses[i]; most = h; }
            else if (h > second) second = h;
            tally[horses[i]] = 0;
        }
        if (most < 2 || most < 2 * second) --tot.good_finds;            /* :1000: TOLERANCE_THRESHOLD 2, SLACK 2 */
        else tot.out_bytes += (uint64_t)fprintf(fo, "%s\t%s\t%f\t%d\n", line + 1, db->labels[most_ix],
                                                (double)1 - (double)second / most, (int)most);   /* :1002 */
    }
    tot.reads = li;
    if (stats) *stats = tot;
    free(line); free(line2); free(horses); free(tally);
    fclose(fp); fclose(fo);
    return rc;
}
